"""How does the time of one 128 x N x 16 tcgen05.mma depend on N?  Times a large-K implicit-GEMM convolution whose
output width equals one n-tile (N = tile_n) for several tile widths: time per (tile, k-step) ~ cost of one MMA.
Usage: python tools/umma_n_sweep.py          (VB_GEMM_PAIR=1 for the cta_group::2 kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vista_b200 import lib, ops
lib.load()
dev = torch.device("cuda:0")
M, C = 460800, 320
geom = (128, 72, 50)
x = (torch.randn(M, C, device=dev) * 0.5).half()
print(f"pair={os.environ.get('VB_GEMM_PAIR', '0')}  conv3x3 M={M} K={9 * C}: N = tile_n")
for tn in (32, 64, 96, 128, 160, 192, 224, 256):
    w = (torch.randn(tn, 9 * C, device=dev) * (9 * C) ** -0.5).half()
    out = torch.empty(M, tn, dtype=torch.float16, device=dev)
    fn = lambda: ops.gemm(x, w, out, taps=ops.TAPS_3X3, geom=geom, tile_n=tn)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tiles_per_sm = (M / 128) / 148
    mma = tiles_per_sm * (9 * C / 16)
    print(f"tile_n={tn:4d}  {ms:7.3f} ms  {2.0 * M * tn * 9 * C / ms / 1e9:7.0f} TFLOP/s   {ms * 1e6 / mma:7.1f} ns per MMA", flush=True)
