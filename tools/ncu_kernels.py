"""Launches the dominant kernels once each at BASELINE shapes (for `ncu --set full`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vista_b200 import lib, ops
from vista_b200.weights import permute_geglu
lib.load()
dev = torch.device("cuda:0")
which = set(sys.argv[1:]) or {"geglu", "cc", "conv", "attn", "gn", "tattn", "ln"}
M, C = 460800, 320
x = (torch.randn(M, C, device=dev) * 0.5).half()
torch.cuda.synchronize()
if "geglu" in which:
    w = (torch.randn(8 * C, C, device=dev) * C ** -0.5)
    b = torch.randn(8 * C, device=dev) * 0.05
    wp, bp = permute_geglu(w, b, 256)
    out = torch.empty(M, 4 * C, dtype=torch.float16, device=dev)
    for _ in range(2):
        ops.gemm(x, wp.half().contiguous(), out, bias=bp, act=2, tile_n=256)
if "cc" in which:
    w = (torch.randn(C, C, device=dev) * C ** -0.5).half()
    b = torch.randn(C, device=dev) * 0.05
    res = torch.randn(M, C, device=dev).half()
    out = torch.empty(M, C, dtype=torch.float16, device=dev)
    for _ in range(2):
        ops.gemm(x, w, out, bias=b, res1=res)
if "conv" in which:
    w = (torch.randn(C, 9 * C, device=dev) * (9 * C) ** -0.5).half()
    b = torch.randn(C, device=dev) * 0.05
    out = torch.empty(M, C, dtype=torch.float16, device=dev)
    for _ in range(2):
        ops.gemm(x, w, out, bias=b, taps=ops.TAPS_3X3, geom=(128, 72, 50))
if "conv1280" in which:
    M2, C2 = 28800, 1280
    x2 = (torch.randn(M2, C2, device=dev) * 0.5).half()
    w = (torch.randn(C2, 9 * C2, device=dev) * (9 * C2) ** -0.5).half()
    b = torch.randn(C2, device=dev) * 0.05
    out = torch.empty(M2, C2, dtype=torch.float16, device=dev)
    for _ in range(2):
        ops.gemm(x2, w, out, bias=b, taps=ops.TAPS_3X3, geom=(32, 18, 50))
if "attn" in which:
    qkv = torch.randn(M, 3 * C, device=dev).half()
    o = torch.empty(M, C, dtype=torch.float16, device=dev)
    for impl in (7,):
        for _ in range(2):
            ops.attention_spatial(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, 50, 9216, 5, impl=impl)
if "tattn" in which:
    qkv = torch.randn(M, 3 * C, device=dev).half()
    o = torch.empty(M, C, dtype=torch.float16, device=dev)
    for _ in range(2):
        ops.attention_temporal(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, 2, 25, 9216, 5)
if "gn" in which:
    g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    y = torch.empty_like(x)
    st = torch.zeros(50, 32, 2, device=dev)
    st[..., 1] = 1.0
    for fps in (1, 25):
        for _ in range(2):
            ops.groupnorm(x, y, 50, 9216, g, bt, 1e-5, True, frames_per_stat=fps)
            ops.groupnorm_apply(x, y, 50, 9216, g, bt, True, st[: 50 // fps], frames_per_stat=fps)
if "ln" in which:
    g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    y = torch.empty_like(x)
    for _ in range(2):
        ops.layernorm(x, y, g, bt)
torch.cuda.synchronize()
print("done")
