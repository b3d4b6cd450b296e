"""CUDA-event timings of the dominant kernels at BASELINE L0/L1 shapes (quick iteration aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vista_b200 import lib, ops
from vista_b200.weights import permute_geglu
lib.load()
dev = torch.device("cuda:0")


def timeit(name, fn, flops=0.0, nbytes=0.0, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:46s} {ms:8.3f} ms  {flops / ms / 1e9:8.0f} TFLOP/s  {nbytes / ms / 1e6:8.0f} GB/s", flush=True)


only = sys.argv[1] if len(sys.argv) > 1 else ""
_timeit = timeit
def timeit(name, fn, *a, **k):
    if only and only not in name:
        return
    _timeit(name, fn, *a, **k)


for (M, C, geom) in ((460800, 320, (128, 72, 50)), (115200, 640, (64, 36, 50)), (28800, 1280, (32, 18, 50))):
    x = (torch.randn(M, C, device=dev) * 0.5).half()
    w = torch.randn(8 * C, C, device=dev) * C ** -0.5
    b = torch.randn(8 * C, device=dev) * 0.05
    wp, bp = permute_geglu(w, b, 256)
    wp = wp.half().contiguous()
    out4 = torch.empty(M, 4 * C, dtype=torch.float16, device=dev)
    timeit(f"gemm GEGLU M={M} N={8*C} K={C}", lambda: ops.gemm(x, wp, out4, bias=bp, act=2, tile_n=256), 2.0 * M * 8 * C * C, 2.0 * M * 5 * C)
    w2 = (torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5).half()
    bb = torch.randn(C, device=dev) * 0.05
    res = torch.randn(M, C, device=dev).half()
    out = torch.empty(M, C, dtype=torch.float16, device=dev)
    timeit(f"gemm FF-down M={M} N={C} K={4*C} +res", lambda: ops.gemm(out4, w2, out, bias=bb, res1=res), 2.0 * M * C * 4 * C, 2.0 * M * 6 * C)
    w1 = (torch.randn(C, C, device=dev) * C ** -0.5).half()
    timeit(f"gemm C->C M={M} N={C} K={C} +res", lambda: ops.gemm(x, w1, out, bias=bb, res1=res), 2.0 * M * C * C, 2.0 * M * 3 * C)
    w3 = (torch.randn(3 * C, C, device=dev) * C ** -0.5).half()
    out3 = torch.empty(M, 3 * C, dtype=torch.float16, device=dev)
    timeit(f"gemm qkv M={M} N={3*C} K={C}", lambda: ops.gemm(x, w3, out3), 2.0 * M * 3 * C * C, 2.0 * M * 4 * C)
    w9 = (torch.randn(C, 9 * C, device=dev) * (9 * C) ** -0.5).half()
    timeit(f"gemm conv3x3 M={M} N={C} K={9*C}", lambda: ops.gemm(x, w9, out, bias=bb, taps=ops.TAPS_3X3, geom=geom), 2.0 * M * C * 9 * C, 2.0 * M * 2 * C)
    heads, seq = C // 64, geom[0] * geom[1]
    o = torch.empty(M, C, dtype=torch.float16, device=dev)
    qkv = torch.randn(M, 3 * C, device=dev).half()
    impls = tuple(int(v) for v in os.environ["BENCH_ATTN_IMPLS"].split(",")) if os.environ.get("BENCH_ATTN_IMPLS") else (3, 7)
    for impl in impls:
        timeit(f"attention spatial v{impl} seq={seq} heads={heads}",
               lambda: ops.attention_spatial(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, 50, seq, heads, impl=impl),
               4.0 * 64 * heads * 50 * seq * seq, 2.0 * 4 * M * C, reps=3)
    g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    y = torch.empty_like(x)
    timeit(f"groupnorm per-frame M={M} C={C}", lambda: ops.groupnorm(x, y, 50, seq, g, bt, 1e-5, True), 0, 2.0 * 3 * M * C)
    timeit(f"layernorm M={M} C={C}", lambda: ops.layernorm(x, y, g, bt), 0, 2.0 * 2 * M * C)
    del x, w, wp, out4, out, out3, qkv, o, y
