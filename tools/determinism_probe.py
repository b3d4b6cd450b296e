"""Bring-up aid: runs each kernel twice on identical inputs at realistic sizes and reports the first
op whose output is not bit-identical; then runs the UNet executor twice with per-op checksums."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from vista_b200 import lib, ops, spec, synth
from vista_b200.unet import UNetRuntime
lib.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0, dt=torch.float16: (torch.randn(*s, generator=g) * scale).to(dt).to(dev)


def twice(name, fn):
    a = fn().clone(); torch.cuda.synchronize()
    b = fn().clone(); torch.cuda.synchronize()
    same = torch.equal(a, b)
    d = (a.float() - b.float()).abs()
    print(f"{name:40s} identical={same} max-diff={float(d.max()):.3e} n-diff={int((d > 0).sum())}/{d.numel()}")


M, C = 50 * 36 * 64, 640
x = rnd(M, C) + 0.5
w = rnd(C, C, scale=C ** -0.5)
out = torch.empty(M, C, dtype=torch.float16, device=dev)
twice("gemm linear", lambda: ops.gemm(x, w, out))
w9 = rnd(C, 9 * C, scale=(9 * C) ** -0.5)
twice("gemm conv3x3", lambda: ops.gemm(x, w9, out, taps=ops.TAPS_3X3, geom=(64, 36, 50)))
gamma, beta = rnd(C, dt=torch.float32) * 0.1 + 1, rnd(C, dt=torch.float32) * 0.1
y = torch.empty_like(x)
def gn(fps):
    return ops.groupnorm(x, y, 50, 36 * 64, gamma, beta, 1e-5, True, frames_per_stat=fps)
twice("groupnorm per-frame", lambda: gn(1))
twice("groupnorm temporal", lambda: gn(25))
twice("layernorm", lambda: ops.layernorm(x, y, gamma, beta))
qkv = rnd(M, 3 * C)
o = torch.empty(M, C, dtype=torch.float16, device=dev)
twice("attention spatial", lambda: ops.attention_spatial(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, 50, 36 * 64, 10))
twice("attention temporal", lambda: ops.attention_temporal(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, 2, 25, 36 * 64, 10))

from helpers import unet_weights, unet_inputs, to_t
for preset, h, w_ in (("tiny", 8, 16), ("small", 16, 32)):
    cfg, sd = unet_weights(preset)
    rt = UNetRuntime(cfg, to_t(sd), dev, 25)
    xx, cc, mask2 = unet_inputs(7, cfg, h, w_, 25)
    B = 50
    xin = torch.from_numpy(np.concatenate([xx * np.float32(0.2), cc["concat"]], 1)).to(dev)
    tok = torch.zeros(B * h * w_, 8, dtype=torch.float16, device=dev)
    ops.nchw_to_tokens(xin.contiguous(), tok, B, 8, h, w_)
    rt.set_conditioning(torch.from_numpy(cc["crossattn"]).to(dev), torch.from_numpy(cc["vector"]).to(dev))
    cn = torch.full((B,), 0.4, device=dev)
    m = torch.from_numpy(mask2).to(dev)
    traces = []
    outs = []
    for r in range(2):
        ops.TRACE = []
        outs.append(rt.forward(tok, cn, m, h, w_).clone())
        traces.append(ops.TRACE)
        ops.TRACE = None
    print(preset, "unet identical:", torch.equal(outs[0], outs[1]), "rel diff", float((outs[0] - outs[1]).norm() / outs[0].norm()))
    nd = 0
    for i, (a, b) in enumerate(zip(*traces)):
        if a != b:
            nd += 1
            if nd <= 8:
                print("  diverge at op", i, a, b)
    print("  ops traced", len(traces[0]), "diverging", nd)
    # untraced (fully async) runs
    o1 = rt.forward(tok, cn, m, h, w_).clone(); o2 = rt.forward(tok, cn, m, h, w_).clone()
    torch.cuda.synchronize()
    print("  async runs identical:", torch.equal(o1, o2), "rel diff", float((o1 - o2).norm() / o1.norm()), "vs traced", float((o1 - outs[0]).norm() / o1.norm()))
