"""The drop-in seam, proven with the reference's OWN caller: the REAL `vwm.models.diffusion.DiffusionEngine`
is built from the reference's vista.yaml parameters with only two strings changed —

    network_wrapper:                          vista_b200.modules.B200Wrapper
    first_stage_config.decoder_config.target: vista_b200.vae.VideoDecoder

— the same checkpoint-style state_dict is loaded into it (identical key names), and the UNMODIFIED
`sample_utils.do_sample` (sample_utils.py:286-375: conditioning, encode_first_stage, two autoregressive rounds with the
decode -> re-condition step in between, final chunked decode) is run on it.  The control arm is the all-reference engine
(OpenAIWrapper + the reference VideoDecoder, CPU fp32) on the same seed; both arms share the real Encoder, Denoiser,
EulerEDMSampler and TrianglePredictionGuider.

Build-container only (needs /root/reference); no GPU here, so the B200 executors run on the emulated C-ABI operators of
tests/fake_ops.py (same rounding points as the kernels) — this test is about the seam and the host logic, the kernels'
numerics are the GPU tests' job.  What had to be patched for a GPU-less host, and nothing else: `load_model` /
`unload_model` (= `.cuda()` / `.cpu()`), the `autocast(device)` scope (CPU autocast would run the control arm in bf16), the
sampler's default `device="cuda"` for its sigma table, and the "CUDA only" guards of the two B200 modules.  The conditioner
is a stand-in (tests/seam_fakes.py): the real one needs the CLIP ViT-H weights, which are not available offline."""
import contextlib
import copy
import io
import sys
import types

import pytest
import torch

from oracle import ref_loader
from vista_b200 import spec, synth

from helpers import rel_l2

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="needs the reference checkout (build container)")

T, H, W = 25, 32, 64


def _reference_sample_utils():
    ref_loader.load_reference()
    if "train" not in sys.modules:           # sample_utils imports one video-writer helper from the training script
        m = types.ModuleType("train")
        m.save_img_seq_to_video = lambda *a, **k: None
        sys.modules["train"] = m
    import sample_utils
    return sample_utils


def _engine_config(native: bool):
    ucfg, dcfg, ecfg = spec.unet_preset("tiny"), spec.decoder_preset("tiny"), spec.encoder_preset("tiny")
    p = copy.deepcopy(ref_loader.vista_yaml()["model"]["params"])
    p["network_config"]["params"].update(model_channels=ucfg.model_channels, attention_resolutions=list(ucfg.attention_resolutions),
                                         num_res_blocks=ucfg.num_res_blocks, channel_mult=list(ucfg.channel_mult))
    p["conditioner_config"] = {"target": "seam_fakes.FakeConditioner", "params": {"down": 2 ** (len(ecfg.ch_mult) - 1)}}
    f = p["first_stage_config"]["params"]
    f["encoder_config"]["params"].update(ch=ecfg.ch, ch_mult=list(ecfg.ch_mult), num_res_blocks=ecfg.num_res_blocks)
    f["decoder_config"]["params"].update(ch=dcfg.ch, ch_mult=list(dcfg.ch_mult), num_res_blocks=dcfg.num_res_blocks)
    if native:                               # the whole integration: two strings
        p["network_wrapper"] = "vista_b200.modules.B200Wrapper"
        f["decoder_config"]["target"] = "vista_b200.vae.VideoDecoder"
    return p, (ucfg, dcfg, ecfg)


def _checkpoint(cfgs):
    ucfg, dcfg, ecfg = cfgs
    sd = {}
    for prefix, specs, seed in (("model.diffusion_model.", spec.unet_param_specs(ucfg), 1),
                                ("first_stage_model.decoder.", spec.decoder_param_specs(dcfg), 2),
                                ("first_stage_model.encoder.", spec.encoder_param_specs(ecfg), 3)):
        for k, v in synth.synth_state_dict(specs, seed=seed).items():
            sd[prefix + k] = torch.from_numpy(v)
    return sd


def _run_do_sample(su, native: bool, rounds: int, steps: int):
    from vwm.models.diffusion import DiffusionEngine
    from fake_ops import patched_ops
    p, cfgs = _engine_config(native)
    with contextlib.redirect_stdout(io.StringIO()):
        eng = DiffusionEngine(**p).eval()
    missing, unexpected = eng.load_state_dict(_checkpoint(cfgs), strict=False)
    assert not unexpected and all(m.startswith("conditioner.") for m in missing), (missing[:3], unexpected[:3])
    ops_ctx = contextlib.nullcontext()
    if native:
        from vista_b200.modules import B200Wrapper
        from vista_b200.vae import DecoderRuntime, VideoDecoder
        assert isinstance(eng.model, B200Wrapper) and isinstance(eng.first_stage_model.decoder, VideoDecoder)
        eng.model._require_cuda = lambda device: None             # GPU-less host: executors on the emulated operators
        dec = eng.first_stage_model.decoder
        with patched_ops():
            rt_dec = DecoderRuntime(dec.b200_config, dec.state_dict(), "cpu")
        dec.runtime = lambda device: rt_dec
        ops_ctx = patched_ops()
    sampler = su.init_sampling(guider="TrianglePredictionGuider", steps=steps, cfg_scale=2.5, num_frames=T)
    sampler.device = "cpu"                                         # its default "cuda" only places the sigma table
    images = torch.from_numpy(synth.normal(21, "seam.img", (T, 3, H, W), std=0.5))
    value_dict = {"cond_frames_without_noise": images[[0]],
                  "cond_frames": images[[0]] + 0.02 * torch.from_numpy(synth.normal(22, "seam.aug", (1, 3, H, W), std=1.0))}
    torch.manual_seed(1234)                                        # do_sample draws its noise from the global RNG
    with ops_ctx, contextlib.redirect_stderr(io.StringIO()):
        samples, samples_z, _ = su.do_sample(images, eng, sampler, value_dict, num_rounds=rounds, num_frames=T,
                                             initial_cond_indices=[0], device="cpu")
    return samples, samples_z


_CONTROL = {}


def _control_arm(su, rounds, steps):
    """The all-reference engine through the real do_sample, once per (rounds, steps): (frames, samples_z, noise drawn per round)."""
    key = (rounds, steps)
    if key not in _CONTROL:
        drawn, real = [], torch.randn_like

        def recording(t, *a, **k):
            out = real(t, *a, **k)
            if t.dim() == 4 and t.shape[0] == T and t.shape[1] == 4:
                drawn.append(out.clone())
            return out
        torch.randn_like = recording
        try:
            x, z = _run_do_sample(su, False, rounds, steps)
        finally:
            torch.randn_like = real
        assert len(drawn) == rounds
        _CONTROL[key] = (x, z, drawn)
    return _CONTROL[key]


def test_unmodified_do_sample_runs_on_the_b200_seams(monkeypatch):
    su = _reference_sample_utils()
    from vista_b200 import fused as fused_mod
    monkeypatch.setattr(fused_mod, "USE_GRAPH", False)
    monkeypatch.setattr(su, "load_model", lambda m: None)
    monkeypatch.setattr(su, "unload_model", lambda m: None)
    monkeypatch.setattr(su, "autocast", lambda device: contextlib.nullcontext())
    rounds, steps = 2, 2
    ref_x, ref_z, _ = _control_arm(su, rounds, steps)
    our_x, our_z = _run_do_sample(su, True, rounds, steps)
    n = rounds * (T - 3) + 3
    assert our_z.shape == ref_z.shape == (n, 4, H // 2, W // 2) and our_x.shape == ref_x.shape == (n, 3, H, W)
    rz, rx = rel_l2(our_z, ref_z), rel_l2(our_x, ref_x)
    print(f"do_sample through the B200 seams vs the all-reference engine: latents rel-L2 {rz:.3e}, frames rel-L2 {rx:.3e}")
    assert rz < 5e-3 and rx < 5e-3, (rz, rx)
    assert torch.equal(our_z[0], ref_z[0])          # sample[0] = z[0] (sample_utils.py:336): the encoder path is shared


def test_engine_rollout_equals_the_real_do_sample(monkeypatch):
    """SURVEY 8f row 2 pinned to the reference's OWN loop: `vista_b200.engine.DiffusionEngine.rollout` (device-side latent
    bookkeeping, `recondition` hook) must reproduce what the unmodified `sample_utils.do_sample` computes on the all-reference
    engine — same encoded clip, same noise draws (recorded from the control run), the re-conditioning between rounds done by
    the reference's own `get_condition` over the same stand-in conditioner (decode -> frame [-3] -> new c / uc)."""
    import yaml, os
    su = _reference_sample_utils()
    from fake_ops import patched_ops
    from vista_b200 import fused as fused_mod
    from vista_b200 import vae as vae_mod
    from vista_b200.diffusion import instantiate_from_config
    monkeypatch.setattr(fused_mod, "USE_GRAPH", False)
    monkeypatch.setattr(su, "load_model", lambda m: None)
    monkeypatch.setattr(su, "unload_model", lambda m: None)
    monkeypatch.setattr(su, "autocast", lambda device: contextlib.nullcontext())
    rounds, steps = 2, 2
    # ---- control arm: the real do_sample on the all-reference engine (shared with the seam test above), with the noise it drew
    ref_x, ref_z, drawn = _control_arm(su, rounds, steps)
    # ---- our engine from configs/inference/vista_b200.yaml (tiny sizes), the same checkpoint, the same stand-in conditioner
    ucfg, dcfg, ecfg = spec.unet_preset("tiny"), spec.decoder_preset("tiny"), spec.encoder_preset("tiny")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "configs", "inference", "vista_b200.yaml")))["model"]
    p = cfg["params"]
    p["network_config"]["params"].update(model_channels=ucfg.model_channels, channel_mult=list(ucfg.channel_mult),
                                         num_res_blocks=ucfg.num_res_blocks, attention_resolutions=list(ucfg.attention_resolutions))
    p["first_stage_config"]["params"]["decoder_config"]["params"].update(ch=dcfg.ch, ch_mult=list(dcfg.ch_mult), num_res_blocks=dcfg.num_res_blocks)
    p["conditioner_config"] = {"target": "seam_fakes.FakeConditioner", "params": {"down": 2 ** (len(ecfg.ch_mult) - 1)}}
    p["sampler_config"]["params"].update(num_steps=steps, device="cpu",
                                         guider_config={"target": "vista_b200.diffusion.TrianglePredictionGuider",
                                                        "params": {"max_scale": 2.5, "num_frames": T}})
    p["en_and_decode_n_samples_a_time"] = 14
    eng = instantiate_from_config(cfg)
    ck = {k: v for k, v in _checkpoint((ucfg, dcfg, ecfg)).items() if not k.startswith("first_stage_model.encoder.")}
    missing, unexpected = eng.load_state_dict(ck, strict=False)
    assert not unexpected and all(m.startswith("_conditioner.") for m in missing), (missing[:3], unexpected[:3])
    eng.model._require_cuda = eng.model.diffusion_model._require_cuda = lambda device: None
    monkeypatch.setattr(vae_mod.VideoDecoder, "runtime", lambda self, device: self.__dict__.setdefault(
        "_rt_cpu", vae_mod.DecoderRuntime(self.b200_config, self.state_dict(), "cpu")))
    images = torch.from_numpy(synth.normal(21, "seam.img", (T, 3, H, W), std=0.5))
    value_dict = {"cond_frames_without_noise": images[[0]],
                  "cond_frames": images[[0]] + 0.02 * torch.from_numpy(synth.normal(22, "seam.aug", (1, 3, H, W), std=1.0))}
    z = ref_z[:T].clone()                       # the encoded clip is shared (real Encoder in both arms): sample[0] = z[0] ...
    # ... but do_sample needs all T latents of the conditioning clip: re-encode with the reference's own encoder path
    from vwm.models.diffusion import DiffusionEngine as RefEngine
    pr, cfgs = _engine_config(False)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_eng = RefEngine(**pr).eval()
    ref_eng.load_state_dict(_checkpoint(cfgs), strict=False)
    torch.manual_seed(1234)                     # the posterior sample of encode_first_stage is the first draw of do_sample
    with torch.no_grad():
        z = ref_eng.encode_first_stage(images)
    assert torch.equal(z[0], ref_z[0])

    def recondition(round_idx, sample, decode_tail):                       # sample_utils.py:340-348, the reference's own code path
        vd = dict(value_dict)
        vd["cond_frames_without_noise"] = decode_tail()[[-3]]
        vd["cond_frames"] = sample[[-3]] / eng.scale_factor
        for e in eng.conditioner.embedders:
            if hasattr(e, "skip_encode"):
                e.skip_encode = True
        cc, ucc = su.get_condition(eng, vd, T, [], "cpu")
        for e in eng.conditioner.embedders:
            if hasattr(e, "skip_encode"):
                e.skip_encode = False
        return cc, ucc
    with patched_ops(), torch.no_grad():
        c, uc = su.get_condition(eng, value_dict, T, [], "cpu")
        frames, samples_z = eng.rollout(c, uc, z, rounds, noises=drawn, recondition=recondition)
    rz, rx = rel_l2(samples_z, ref_z), rel_l2(frames, ref_x)
    print(f"engine.rollout vs the real do_sample ({rounds} rounds): latents rel-L2 {rz:.3e}, frames rel-L2 {rx:.3e}")
    assert samples_z.shape == ref_z.shape and frames.shape == ref_x.shape
    assert rz < 5e-3 and rx < 5e-3, (rz, rx)


def test_engine_sample_ensemble_equals_the_real_reward_do_sample(monkeypatch):
    """SURVEY 8f row 3 pinned to the reference's OWN loop: `engine.sample_ensemble` against the unmodified
    `reward_utils.do_sample` (reward_utils.py:285-340) on the all-reference engine: same encoded clip, the member noises
    recorded from the control run; the reward exp(-mean unbiased variance) must agree."""
    import yaml, os
    _reference_sample_utils()                       # installs the `train` stub reward_utils imports too
    import reward_utils as ru
    from fake_ops import patched_ops
    from vista_b200 import fused as fused_mod
    from vista_b200.diffusion import instantiate_from_config
    from vwm.models.diffusion import DiffusionEngine as RefEngine
    monkeypatch.setattr(fused_mod, "USE_GRAPH", False)
    monkeypatch.setattr(ru, "load_model", lambda m: None)
    monkeypatch.setattr(ru, "unload_model", lambda m: None)
    monkeypatch.setattr(ru, "autocast", lambda device: contextlib.nullcontext())
    K, steps = 3, 2
    pr, cfgs = _engine_config(False)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_eng = RefEngine(**pr).eval()
    ref_eng.load_state_dict(_checkpoint(cfgs), strict=False)
    sampler = ru.init_sampling(guider="VanillaCFG", steps=steps, cfg_scale=2.5, num_frames=T)
    sampler.device = "cpu"
    images = torch.from_numpy(synth.normal(21, "seam.img", (T, 3, H, W), std=0.5))
    value_dict = {"cond_frames_without_noise": images[[0]],
                  "cond_frames": images[[0]] + 0.02 * torch.from_numpy(synth.normal(22, "seam.aug", (1, 3, H, W), std=1.0))}
    drawn, real = [], torch.randn_like

    def recording(t, *a, **k):
        out = real(t, *a, **k)
        if t.dim() == 4 and t.shape[0] == T and t.shape[1] == 4:
            drawn.append(out.clone())
        return out
    zs = []
    real_encode = ref_eng.encode_first_stage
    ref_eng.encode_first_stage = lambda x: (zs.append(real_encode(x)), zs[-1])[1]
    monkeypatch.setattr(torch, "randn_like", recording)
    torch.manual_seed(4321)
    with contextlib.redirect_stderr(io.StringIO()):
        _, ref_reward = ru.do_sample(images, ref_eng, sampler, value_dict, num_frames=T, ensemble_size=K, initial_cond_indices=[0], device="cpu")
    monkeypatch.setattr(torch, "randn_like", real)
    assert len(drawn) == K and len(zs) == 1
    # ---- our engine (vista_b200.yaml at tiny sizes) on the emulated operators, same checkpoint and stand-in conditioner
    ucfg, dcfg, ecfg = cfgs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "configs", "inference", "vista_b200.yaml")))["model"]
    p = cfg["params"]
    p["network_config"]["params"].update(model_channels=ucfg.model_channels, channel_mult=list(ucfg.channel_mult),
                                         num_res_blocks=ucfg.num_res_blocks, attention_resolutions=list(ucfg.attention_resolutions))
    p["first_stage_config"]["params"]["decoder_config"]["params"].update(ch=dcfg.ch, ch_mult=list(dcfg.ch_mult), num_res_blocks=dcfg.num_res_blocks)
    p["conditioner_config"] = {"target": "seam_fakes.FakeConditioner", "params": {"down": 2 ** (len(ecfg.ch_mult) - 1)}}
    p["sampler_config"]["params"].update(num_steps=steps, device="cpu")
    eng = instantiate_from_config(cfg)
    ck = {k: v for k, v in _checkpoint(cfgs).items() if not k.startswith("first_stage_model.encoder.")}
    eng.load_state_dict(ck, strict=False)
    eng.model._require_cuda = eng.model.diffusion_model._require_cuda = lambda device: None
    with patched_ops(), torch.no_grad():
        c, uc = ru.get_condition(eng, value_dict, T, [], "cpu")
        reward, members = eng.sample_ensemble(c, uc, zs[0], K, noises=drawn)
    print(f"ensemble reward: engine {float(reward):.6f} vs the real reward_utils.do_sample {float(ref_reward):.6f}")
    assert abs(float(reward) - float(ref_reward)) < 2e-3 * max(1.0, abs(float(ref_reward))), (float(reward), float(ref_reward))
