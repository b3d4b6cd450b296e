"""SURVEY.md 8f rows 2-4 on the GPU: the glue kernels (csrc/glue.cu) against their CPU restatements in tests/fake_ops.py
(bit-exact: byte / copy / fixed-order work), then engine.rollout (uint8 output included) and engine.sample_ensemble
against the same loops over the CPU oracle (fp16 tolerance of the sampler / decoder)."""
import os

import pytest
import torch
import yaml

import fake_ops
from helpers import decoder_weights, rel_l2, to_t, unet_weights
from vista_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from vista_b200 import lib, ops as o
    lib.load()
    return o


def test_time_mix_small_u8_matches_reference_output_path(ops):
    torch.manual_seed(0)
    T, H, W, C, f0 = 5, 16, 32, 3, 2
    HW = H * W
    x = torch.randn(T * HW, 8, device=DEV) * 0.8
    w = torch.randn(C, C, 3, device=DEV) * 0.5
    b = torch.randn(C, device=DEV) * 0.1
    blend = torch.tensor([1, 1, 0, 0, 0], dtype=torch.int32, device=DEV)
    prev = torch.randn(f0 + T, C, H, W, device=DEV)
    out, out_ref = prev.clone(), prev.clone()
    ops.time_mix_small(x, w, b, out_ref, blend, T, HW, C, f0, 0)                       # the fp32 kernel of round 1
    out8 = torch.zeros(f0 + T, H, W, C, dtype=torch.uint8, device=DEV)
    ops.time_mix_small_u8(x, w, b, out, out8, blend, T, HW, C, f0, 0, keep_f32_from=3)
    torch.cuda.synchronize()
    # bytes == numpy's (255 * clamp((x + 1) / 2, 0, 1)).astype(uint8) in "t h w c" of the fp32 frames (sample_utils.py:96-126,374)
    want = (255.0 * torch.clamp((out_ref[f0:] + 1.0) / 2.0, 0.0, 1.0)).to(torch.uint8).permute(0, 2, 3, 1)
    assert torch.equal(out8[f0:], want)
    assert torch.equal(out8[:f0], torch.zeros_like(out8[:f0]))
    # fp32 kept only from frame 3 of the call on; earlier frames untouched
    assert torch.equal(out[f0 + 3:], out_ref[f0 + 3:]) and torch.equal(out[:f0 + 3], prev[:f0 + 3])
    # and the CPU restatement used by the executor tests agrees bit for bit
    o_cpu, o8_cpu = prev.cpu().clone(), torch.zeros(f0 + T, H, W, C, dtype=torch.uint8)
    fake_ops.time_mix_small_u8(x.cpu(), w.cpu(), b.cpu(), o_cpu, o8_cpu, blend.cpu(), T, HW, C, f0, 0, keep_f32_from=3)
    assert (o8_cpu[f0:].int() - out8[f0:].cpu().int()).abs().max() <= 1        # fma vs separate rounding in the 3x3x3 mix


def test_rollout_advance_and_ensemble_reward(ops):
    torch.manual_seed(1)
    T, shape, rounds = 25, (4, 8, 16), 3
    z = torch.randn(T, *shape, device=DEV)
    samples_z = torch.zeros(rounds * (T - 3) + 3, *shape, device=DEV)
    filled = torch.full((T, *shape), 7.0, device=DEV)
    ref_z, ref_f = samples_z.cpu().clone(), filled.cpu().clone()
    for n in range(rounds):
        s = torch.randn(T, *shape, device=DEV)
        s_cpu = s.cpu().clone()
        last = n + 1 == rounds
        ops.rollout_advance(s, z if n == 0 else None, samples_z, None if last else filled, n * (T - 3), 0 if n == 0 else 3, 3)
        fake_ops.rollout_advance(s_cpu, z.cpu() if n == 0 else None, ref_z, None if last else ref_f, n * (T - 3), 0 if n == 0 else 3, 3)
        torch.cuda.synchronize()
        assert torch.equal(s.cpu(), s_cpu)            # sample[0] = z[0] on the first round only
        assert torch.equal(filled.cpu(), ref_f)
    assert torch.equal(samples_z.cpu(), ref_z)
    K = 5
    members = [torch.randn(T, *shape, device=DEV) * (1 + 0.1 * k) for k in range(K)]
    out = ops.ensemble_reward(members)
    out2 = ops.ensemble_reward(members)
    torch.cuda.synchronize()
    ref = fake_ops.ensemble_reward([m.cpu() for m in members])
    assert torch.equal(out, out2)                     # fixed-order reduction: run-to-run identical
    assert abs(float(out[0]) - float(ref[0])) < 1e-5 * float(ref[0]) and abs(float(out[1]) - float(ref[1])) < 1e-6


def _engine(steps, guider=None):
    from vista_b200.diffusion import instantiate_from_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "configs", "inference", "vista_b200.yaml")))["model"]
    p = cfg["params"]
    p["network_config"]["params"].update(model_channels=64, channel_mult=[1, 2], num_res_blocks=1, attention_resolutions=[1, 2])
    p["first_stage_config"]["params"]["decoder_config"]["params"].update(ch=64, ch_mult=[1, 2], num_res_blocks=1)
    p["sampler_config"]["params"]["num_steps"] = steps
    if guider is not None:
        p["sampler_config"]["params"]["guider_config"] = guider
    p["replace_cond_frames"], p["fixed_cond_frames"] = True, [0]
    p["en_and_decode_n_samples_a_time"] = 14
    ucfg, usd = unet_weights("tiny")
    dcfg, dsd = decoder_weights("tiny")
    with torch.device(DEV):
        eng = instantiate_from_config(cfg)
    sd = {"model.diffusion_model." + k: torch.from_numpy(v) for k, v in usd.items()}
    sd.update({"first_stage_model.decoder." + k: torch.from_numpy(v) for k, v in dsd.items()})
    missing, unexpected = eng.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    return eng, (ucfg, usd), (dcfg, dsd)


def test_engine_rollout_u8_and_ensemble_vs_oracle():
    from oracle import vista_oracle as vo
    from test_executor_cpu import _reference_rollout
    T, h, w, steps, rounds = 25, 8, 16, 3, 3
    guider = {"target": "vista_b200.diffusion.TrianglePredictionGuider", "params": {"max_scale": 2.5, "num_frames": T}}
    eng, (ucfg, usd), (dcfg, dsd) = _engine(steps, guider)
    c, uc = synth.synth_conditioning(7, T, h, w, trajectory=True, context_dim=ucfg.context_dim, adm=ucfg.adm_in_channels)
    _, z, _ = synth.synth_latents(7, T, h, w)
    zt = torch.from_numpy(z)
    noises = [torch.from_numpy(synth.normal(40 + i, "rollout.noise", (T, 4, h, w), std=1.0)) for i in range(rounds)]
    frames, samples_z = eng.rollout(to_t(c, DEV), to_t(uc, DEV), zt.to(DEV), rounds, noises=noises)
    frames8, samples_z8 = eng.rollout(to_t(c, DEV), to_t(uc, DEV), zt.to(DEV), rounds, noises=noises, u8=True)
    torch.cuda.synchronize()
    assert torch.equal(samples_z, samples_z8)                      # same launches, deterministic
    want8 = (255.0 * frames).to(torch.uint8).permute(0, 2, 3, 1)   # reference output path on our fp32 frames
    assert torch.equal(frames8, want8)
    sdt = to_t(usd)
    with torch.no_grad():
        ref_z = _reference_rollout(lambda nz, cc, cf, m: vo.euler_edm_sample(sdt, ucfg, nz, cc, to_t(uc), cf, m, steps, T,
                                                                           guider="TrianglePredictionGuider", scale=2.5),
                                   to_t(c), zt, noises, T, eng.scale_factor)
        ref_frames = torch.clamp((vo.decode_first_stage(to_t(dsd), dcfg, ref_z) + 1.0) / 2.0, 0.0, 1.0)
    r1, r2 = rel_l2(samples_z.cpu(), ref_z), rel_l2(frames.cpu(), ref_frames)
    ref8 = (255.0 * ref_frames).to(torch.uint8).permute(0, 2, 3, 1)
    d8 = (frames8.cpu().int() - ref8.int()).abs()
    print(f"rollout {rounds} rounds: latents rel-L2 {r1:.3e}, frames {r2:.3e}; uint8 max diff {int(d8.max())}, differing bytes {float((d8 > 0).float().mean()):.3%}")
    assert r1 < 5e-3 and r2 < 1e-2 and int(d8.max()) <= 3

    eng2, _, _ = _engine(steps)
    K = 3
    en = [torch.from_numpy(synth.normal(60 + i, "ens.noise", (T, 4, h, w), std=1.0)) for i in range(K)]
    reward, members = eng2.sample_ensemble(to_t(c, DEV), to_t(uc, DEV), zt.to(DEV), K, noises=en)
    torch.cuda.synchronize()
    mask = torch.zeros(T)
    mask[0] = 1
    with torch.no_grad():
        refs = []
        for i in range(K):
            s = vo.euler_edm_sample(sdt, ucfg, en[i].clone(), to_t(c), to_t(uc), zt, mask, steps, T)
            s[0] = zt[0]
            refs.append(s)
        ref_reward = fake_ops.ensemble_reward(refs)[1]
    print(f"ensemble reward {float(reward):.6f} vs oracle {float(ref_reward):.6f}")
    assert all(rel_l2(a.cpu(), b) < 5e-3 for a, b in zip(members, refs))
    assert abs(float(reward) - float(ref_reward)) < 2e-3
