"""Per-kernel-family breakdown of the chunked temporal-VAE decode of one 25-frame clip at 576x1024, measured with
CUDA events around every launch.  Writes a markdown table.
Usage: python tools/profile_decode.py [out.md]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from vista_b200 import lib, ops, spec
from vista_b200.vae import DecoderRuntime, decode_first_stage

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "decode_breakdown.md")
lib.load()
dev = torch.device("cuda:0")
ucfg, dcfg, h, w, rand_sd = bench.make_problem("full", dev)
T = 25
rt = DecoderRuntime(dcfg, rand_sd(spec.decoder_param_specs(dcfg)), dev)
z = torch.randn(T, dcfg.z_channels, h, w, device=dev) * 0.18215
decode_first_stage(rt, z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); decode_first_stage(rt, z); e1.record(); torch.cuda.synchronize()
wall_ms = e0.elapsed_time(e1)
ops.PROFILE = []
decode_first_stage(rt, z)
torch.cuda.synchronize()
rec, ops.PROFILE = ops.PROFILE, None
fam, det = ops.profile_summary(rec)
tot = sum(r["ms"] for r in fam.values())
peaks = bench.load_peaks()
lines = ["# Chunked decode of one clip (25 x 576 x 1024): kernel time by family (CUDA events per launch, eager)", "",
         f"decode wall (events) {wall_ms:.1f} ms; sum of profiled kernel times {tot:.1f} ms; peaks: {peaks}", "",
         "| family | launches | ms | share | TFLOP/s | frac of peak | GB/s (algorithmic) |", "|---|---|---|---|---|---|---|"]
for k, r in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    lines.append(f"| {k} | {r['launches']} | {r['ms']:.2f} | {100 * r['ms'] / tot:.1f}% | {r['tflops']:.0f} | {r['tflops'] / peaks['tflops']:.2f} | {r['gbs']:.0f} |")
lines += ["", "## by shape (top 40)", "", "| family | detail | launches | ms | TFLOP/s | GB/s |", "|---|---|---|---|---|---|"]
for (f, d), r in sorted(det.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    lines.append(f"| {f} | {d} | {r['launches']} | {r['ms']:.2f} | {r['tflops']:.0f} | {r['gbs']:.0f} |")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:14 + len(fam)]))
