"""SM cycles per tcgen05.mma (kind::f16, K=16) by shape and accumulator count (b200v_debug_mma_probe).
Usage: python tools/mma_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes, subprocess
import torch
from vista_b200 import lib
# the probe is a diagnostic, not product code: it is built here into its own library (linked against the product's host
# helpers through libvista_b200.so) instead of shipping inside it
lib.load()
so = os.path.join(ROOT, "tools", "libvista_b200_probe.so")
src = os.path.join(ROOT, "tools", "csrc", "mma_probe.cu")
if not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.run(["nvcc", *lib.NVCC_FLAGS, "-shared", "-o", so, src, lib.LIB_PATH, "-Xlinker", "-rpath," + lib.PKG_DIR], check=True)
l = ctypes.CDLL(so)
l.b200v_debug_mma_probe.restype = ctypes.c_int
l.b200v_debug_mma_probe.argtypes = [ctypes.c_int32] * 6 + [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
dev = torch.device("cuda:0")
n_ctas = torch.cuda.get_device_properties(dev).multi_processor_count
out = torch.zeros(n_ctas, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
print(f"{n_ctas} CTAs (one per SM), 2000 x 4 MMAs each; cycles per MMA (median over CTAs)")
print("| M | N | accumulators | A operand | cycles / MMA | nominal (max(M,128)*N/256) |")
print("|---|---|---|---|---|---|")
for M in (128, 64):
    for N in (64, 80, 128, 192, 256):
        if N % (16 if M == 128 else 8):
            continue
        for n_acc in (1, 2):
            for mode, (a_mn, a_tm) in (("smem K-major", (0, 0)), ("smem MN-major", (1, 0)), ("TMEM", (0, 1))):
                if n_acc * N > (480 if a_tm else 512):
                    continue
                lib.check(l.b200v_debug_mma_probe(M, N, 2000, n_acc, a_mn, a_tm, out.data_ptr(), n_ctas, stream), "mma_probe")
                torch.cuda.synchronize()
                print(f"| {M} | {N} | {n_acc} | {mode} | {out.median().item():.1f} | {max(M, 128) * N / 256:.0f} |", flush=True)
