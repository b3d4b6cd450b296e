"""Per-kernel-family breakdown of one full-size EDM step, measured with CUDA events around every launch
(eager launches, so the inter-kernel gaps are excluded).  Writes a markdown table.
Usage: python tools/profile_step.py [out.md] [--config full|small]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from vista_b200 import lib, ops, spec
from vista_b200.unet import UNetRuntime

out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(ROOT, "gpurun_out", "step_breakdown.md")
config = "small" if "--config=small" in sys.argv else "full"
lib.load()
dev = torch.device("cuda:0")
ucfg, dcfg, h, w, rand_sd = bench.make_problem(config, dev)
T, B = 25, 50
rt = UNetRuntime(ucfg, rand_sd(spec.unet_param_specs(ucfg)), dev, T)
ctx = torch.randn(B, 1, 3456, device=dev)
y = torch.randn(B, ucfg.adm_in_channels, device=dev)
rt.set_conditioning(ctx, y)
from vista_b200.unet import padded_input_rows
tok = padded_input_rows(B * h * w, dev)          # the production layout (fused sampler): input conv on the tensor cores
tok.copy_(torch.randn(B * h * w, 8, device=dev).half())
cn = torch.full((B,), 0.5, device=dev)
mask = torch.zeros(B, device=dev); mask[0] = mask[T] = 1
for _ in range(2):
    rt.forward(tok, cn, mask, h, w)
torch.cuda.synchronize()
if "--ncu" in sys.argv:       # one step between profiler start/stop for `ncu --profile-from-start off`
    torch.cuda.cudart().cudaProfilerStart()
    rt.forward(tok, cn, mask, h, w)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); rt.forward(tok, cn, mask, h, w); e1.record(); torch.cuda.synchronize()
eager_ms = e0.elapsed_time(e1)
ops.PROFILE = []
rt.forward(tok, cn, mask, h, w)
torch.cuda.synchronize()
rec, ops.PROFILE = ops.PROFILE, None
fam, det = ops.profile_summary(rec)
tot = sum(r["ms"] for r in fam.values())
peaks = bench.load_peaks()
lines = [f"# One EDM step ({config}): kernel time by family (CUDA events per launch, eager)", "",
         f"eager step wall (events) {eager_ms:.1f} ms; sum of kernel times {tot:.1f} ms; peaks: {peaks}", "",
         "| family | launches | ms | share | TFLOP/s | frac of peak | GB/s (algorithmic) |", "|---|---|---|---|---|---|---|"]
for k, r in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    lines.append(f"| {k} | {r['launches']} | {r['ms']:.2f} | {100 * r['ms'] / tot:.1f}% | {r['tflops']:.0f} | {r['tflops'] / peaks['tflops']:.2f} | {r['gbs']:.0f} |")
lines += ["", "## by shape (top 40)", "", "| family | detail | launches | ms | TFLOP/s | GB/s |", "|---|---|---|---|---|---|"]
for (f, d), r in sorted(det.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    lines.append(f"| {f} | {d} | {r['launches']} | {r['ms']:.2f} | {r['tflops']:.0f} | {r['gbs']:.0f} |")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:14 + len(fam)]))
