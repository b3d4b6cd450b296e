"""ncu launch list (CSV with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of ONE step
-> per-kernel-family table (markdown on stdout) + profiles/r02_traffic.json (DRAM bytes of the tap-GEMM launches).
Usage: python tools/ncu_traffic.py gpurun_out/r02_ncu_launches.csv [out.json]"""
import csv, json, os, re, sys
from collections import defaultdict

path = sys.argv[1]
out_json = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
launch = defaultdict(dict)
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    key = int(r[col["ID"]])
    launch[key]["name"] = r[col["Kernel Name"]]
    val = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    m = r[col["Metric Name"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1,
             "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)
    launch[key][m] = val * scale


def family(name):
    if "tapgemm" in name: return "tapgemm (tcgen05)"
    if "attn7" in name or "attn5" in name or "attn3" in name or "attn2" in name or "attn_spatial" in name: return "spatial attention (tcgen05)"
    if "attn_temporal" in name: return "temporal attention"
    if name.startswith("void vb::gn_") or "gn_" in name: return "groupnorm"
    if "layernorm" in name: return "layernorm"
    return "other"


fam = defaultdict(lambda: dict(n=0, t=0.0, rd=0.0, wr=0.0))
for k, v in launch.items():
    f = fam[family(v["name"])]
    f["n"] += 1
    f["t"] += v.get("gpu__time_duration.sum", 0.0)
    f["rd"] += v.get("dram__bytes_read.sum", 0.0)
    f["wr"] += v.get("dram__bytes_write.sum", 0.0)
tot = sum(f["t"] for f in fam.values())
print(f"| family | launches | ms (ncu, cold / serialised) | share | DRAM read GB | DRAM write GB |\n|---|---|---|---|---|---|")
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["t"]):
    print(f"| {k} | {f['n']} | {f['t'] * 1e3:.2f} | {f['t'] / tot:.1%} | {f['rd'] / 1e9:.2f} | {f['wr'] / 1e9:.2f} |")
print(f"\n{len(launch)} launches, {tot * 1e3:.1f} ms")
g = fam["tapgemm (tcgen05)"]
json.dump({"gemm_dram_bytes_per_step": g["rd"] + g["wr"], "gemm_launches": g["n"], "gemm_ms_ncu": g["t"] * 1e3,
           "step_launches": len(launch),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the tap-GEMM launches of one EDM step "
                   "(ncu --clock-control none, tools/one_step.py); compare with roofline.algorithmic_bytes"},
          open(out_json, "w"), indent=1)
