"""One eager EDM step of the full BASELINE configuration between cudaProfilerStart / Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/r02_ncu_launches.csv python tools/one_step.py
(tools/ncu_traffic.py turns the CSV into profiles/r02_traffic.json and the launch-list summary)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VISTA_B200_GRAPH", "0")
import torch
import bench
from vista_b200 import lib
from vista_b200.diffusion import B200Denoiser

lib.load()
dev = torch.device("cuda:0")
config = sys.argv[1] if len(sys.argv) > 1 else "full"
eng, ucfg, dcfg, h, w = bench.build_engine(config, dev)
T = 25
c_h, uc_h, noise_h, z_h, mask_h = bench.host_inputs(ucfg, T, h, w)
td = lambda d: {k: v.to(dev) for k, v in d.items()}
bden = B200Denoiser(eng.denoiser, eng.model)
x = noise_h.to(dev)
eng.sampler(bden, x, td(c_h), uc=td(uc_h), cond_frame=z_h.to(dev), cond_mask=mask_h.to(dev), num_steps=3)   # warm-up, eager
rt = eng.model._rt_get(eng.model.diffusion_model, T, dev)
st = rt._loop_states[(T, h, w)]
torch.cuda.synchronize()
torch.cuda.profiler.start()
st.one_step(rt, 50)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one step done")
