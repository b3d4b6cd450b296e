"""CPU tests (no GPU): C-ABI surface, host logic, reference-surface mirrors."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import vista_oracle as vo
from vista_b200 import spec, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vista_b200 import lib
    lib.build()
    l = lib.load()
    declared = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            declared |= set(re.findall(r"\b(b200v_\w+)\s*\(", open(os.path.join(ROOT, "include", fn)).read()))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(l, name), f"{name} declared in include/ but not exported"
    assert declared == set(lib.exported_symbols())
    assert l.b200v_version() >= 100
    assert isinstance(l.b200v_last_error(), bytes)


def test_gemm_argument_validation_without_gpu():
    """Bad descriptors are rejected by the host code before any CUDA call (error convention of §8b)."""
    import ctypes as C
    from vista_b200 import lib
    l = lib.load()
    d = lib.GemmDesc()
    assert l.b200v_gemm(C.byref(d), None) != 0
    assert b"null pointer" in l.b200v_last_error()
    d.a = d.b = d.out = 16
    d.cin, d.ntaps, d.N, d.tile_n = 60, 1, 64, 64
    assert l.b200v_gemm(C.byref(d), None) != 0
    assert b"cin" in l.b200v_last_error()


def test_no_cpu_fallback():
    from vista_b200.modules import B200Wrapper, VideoUNet
    cfg = spec.unet_preset("tiny")
    unet = VideoUNet(in_channels=8, model_channels=64, out_channels=4, num_res_blocks=1, attention_resolutions=[1, 2],
                     channel_mult=[1, 2], num_head_channels=64, num_classes="sequential", context_dim=1024,
                     adm_in_channels=768, extra_ff_mix_layer=True, use_spatial_context=True, video_kernel_size=[3, 1, 1],
                     use_linear_in_transformer=True, action_control=True)
    assert set(unet.state_dict()) == set(spec.unet_param_specs(cfg))
    assert {k: tuple(v.shape) for k, v in unet.state_dict().items()} == {k: tuple(v[0]) for k, v in spec.unet_param_specs(cfg).items()}
    net = B200Wrapper(unet)
    assert set(net.state_dict()) == {"diffusion_model." + k for k in spec.unet_param_specs(cfg)}
    x = torch.zeros(50, 4, 8, 16)
    c = {"concat": torch.zeros(50, 4, 8, 16), "crossattn": torch.zeros(50, 1, 3456), "vector": torch.zeros(50, 768)}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, torch.zeros(50), c, torch.zeros(50), 25)


def test_unsupported_options_raise():
    from vista_b200.modules import VideoUNet
    kw = dict(in_channels=8, model_channels=64, out_channels=4, num_res_blocks=1, attention_resolutions=[1], channel_mult=[1],
              num_head_channels=64, num_classes="sequential", context_dim=1024, adm_in_channels=768, extra_ff_mix_layer=True,
              use_spatial_context=True, video_kernel_size=[3, 1, 1], use_linear_in_transformer=True)
    VideoUNet(**kw)
    for bad in (dict(use_scale_shift_norm=True), dict(transformer_depth=2), dict(add_lora=True), dict(num_classes=None)):
        with pytest.raises(NotImplementedError):
            VideoUNet(**{**kw, **bad})


def test_plan_matches_survey_appendix_a():
    plan = spec.build_unet_plan(spec.unet_preset("vista"))
    assert len(plan.input_blocks) == 12 and len(plan.output_blocks) == 12
    assert len(plan.res_blocks()) == 22 and len(plan.transformers()) == 16
    assert sum(1 for r in plan.res_blocks() if r.has_skip) == 14
    assert [t.heads for t in plan.transformers()][:6] == [5, 5, 10, 10, 20, 20]
    assert plan.skip_channels == [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
    assert plan.output_blocks[0].layers[0].cin == 2560 and plan.output_blocks[-1].layers[0].cin == 640


def test_tile_pickers():
    from vista_b200 import ops
    for geom in [(128, 72, 50), (64, 36, 50), (32, 18, 50), (16, 9, 50), (9216, 25, 2), (144, 25, 2), (2, 25, 2), (1024, 576, 14)]:
        bw, bh, bb = ops.pick_box(*geom)
        assert bw * bh * bb == 128
    assert ops.pick_box(128, 72, 50)[1] * ops.pick_box(128, 72, 50)[0] == 128
    assert ops.pick_box(32, 18, 50) == (32, 2, 2)      # exact cover: 225 tiles
    # tile widths follow the measured cost of one MMA, t(n) = 61 + 0.22 max(n, 128) ns: wide tiles win even with a
    # partly empty last tile (profiles/r01_umma_n_sweep.md)
    assert ops.pick_tile_n(320) == 160 and ops.pick_tile_n(640) == 224 and ops.pick_tile_n(960) == 256
    assert ops.pick_tile_n(1920) == 256 and ops.pick_tile_n(1280) == 256 and ops.pick_tile_n(2560, True) == 256
    for n in (64, 96, 320, 640, 1280, 3840, 2560, 5120, 10240):
        tn = ops.pick_tile_n(n)
        assert 32 <= tn <= 256 and tn % 32 == 0


def test_weight_repacking():
    from vista_b200.weights import conv_weight_to_taps, permute_geglu
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    t = conv_weight_to_taps(w)
    assert t.shape == (2, 27) and float(t[1, (1 * 3 + 2) * 3 + 1]) == float(w[1, 1, 1, 2])   # K = (kh*3+kw)*I + i
    w5 = torch.arange(2 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 1, 1)
    assert float(conv_weight_to_taps(w5)[1, 2 * 3 + 1]) == float(w5[1, 1, 2, 0, 0])
    wg = torch.arange(16 * 2, dtype=torch.float32).reshape(16, 2)          # inner = 8
    bg = torch.arange(16, dtype=torch.float32)
    wp, bp = permute_geglu(wg, bg, tile_n=8)                                # h = 4
    assert bp.tolist() == [0, 1, 2, 3, 8, 9, 10, 11, 4, 5, 6, 7, 12, 13, 14, 15]
    assert torch.equal(wp[:, 0], wg[:, 0][bp.long()])


def test_diffusion_mirrors_match_oracle_on_cpu():
    """EulerEDMSampler / guiders / Denoiser mirrors with a cheap stand-in network == the oracle algebra."""
    from vista_b200.diffusion import (Denoiser, EDMDiscretization, EulerEDMSampler, LinearPredictionGuider,
                                      TrianglePredictionGuider, VanillaCFG)
    assert torch.equal(EDMDiscretization(0.002, 700.0, 7.0)(50), vo.edm_sigmas(50))
    assert torch.allclose(TrianglePredictionGuider(25, 2.5, 1.0).scale_vector(25), vo.triangle_scales(25))
    assert torch.allclose(LinearPredictionGuider(25, 2.5, 1.0).scale_vector(25), vo.guider_scales("LinearPredictionGuider", 25, 2.5))
    assert torch.equal(VanillaCFG(2.5).scale_vector(3), torch.full((3,), 2.5))
    T, h, w, steps = 25, 4, 6, 5
    g = torch.Generator().manual_seed(0)
    Wm = torch.randn(4, 8, generator=g) * 0.3

    def network(x, c_noise, c, cond_mask, num_frames):       # (B,4,h,w) -> (B,4,h,w), depends on everything
        xin = torch.cat([x, c["concat"]], 1)
        return torch.einsum("oc,bchw->bohw", Wm, xin) * (1 + c_noise[:, None, None, None]) + cond_mask[:, None, None, None] \
            + c["vector"][:, :1, None, None]

    den = Denoiser({"target": "vista_b200.diffusion.VScalingWithEDMcNoise"}, num_frames=T)
    c = {"concat": torch.randn(T, 4, h, w, generator=g), "crossattn": torch.randn(T, 1, 8, generator=g), "vector": torch.randn(T, 3, generator=g)}
    uc = {"concat": torch.zeros(T, 4, h, w), "crossattn": torch.zeros(T, 1, 8), "vector": c["vector"].clone()}
    noise, z = torch.randn(T, 4, h, w, generator=g), torch.randn(T, 4, h, w, generator=g)
    mask = torch.zeros(T)
    mask[[0, 1]] = 1
    for guider, gcfg in (("VanillaCFG", {"target": "vista_b200.diffusion.VanillaCFG", "params": {"scale": 2.5}}),
                         ("TrianglePredictionGuider", {"target": "vista_b200.diffusion.TrianglePredictionGuider", "params": {"max_scale": 2.5}})):
        smp = EulerEDMSampler(num_steps=steps, device="cpu", s_churn=0.0, s_tmin=0.0, s_tmax=999.0, s_noise=1.0,
                              discretization_config={"target": "vista_b200.diffusion.EDMDiscretization",
                                                     "params": {"sigma_min": 0.002, "sigma_max": 700.0, "rho": 7.0}},
                              guider_config=gcfg)
        x_in = noise.clone()
        out = smp(lambda x, s, cc, m: den(network, x, s, cc, m), x_in, c, uc=uc, cond_frame=z, cond_mask=mask)
        assert torch.allclose(x_in, noise * torch.sqrt(1 + vo.edm_sigmas(steps)[0] ** 2))    # in-place scaling (sampling.py:36)
        # oracle algebra with the same stand-in network
        sig = vo.edm_sigmas(steps)
        x = noise * torch.sqrt(1 + sig[0] ** 2)
        sc = vo.guider_scales(guider, T, 2.5)[:, None, None, None]
        m4 = mask[:, None, None, None]
        for i in range(steps):
            x = x * (1 - m4) + z * m4
            s = torch.full((2 * T,), float(sig[i]))
            cs, co, ci, cn = vo.vscaling_edm_cnoise(s[:, None, None, None])
            x2 = torch.cat([x, x])
            cc = {k: torch.cat([uc[k], c[k]]) for k in c}
            d = network(x2 * ci, cn.reshape(-1), cc, torch.cat([mask, mask]), T) * co + x2 * cs
            du, dc = d.chunk(2)
            d = du + sc * (dc - du)
            x = x + (x - d) / sig[i] * (sig[i + 1] - sig[i])
        x = x * (1 - m4) + z * m4
        assert torch.allclose(out, x, rtol=1e-5, atol=1e-5), guider


def test_synth_is_platform_stable():
    # pinned values: numpy Philox streams are specified to be reproducible across platforms
    a = synth.normal(1, "x", (4,))
    assert np.allclose(a, synth.normal(1, "x", (4,)))
    assert synth.checksum([a]) == synth.checksum([synth.normal(1, "x", (4,))])
    sd = synth.synth_state_dict(spec.unet_param_specs(spec.unet_preset("tiny")), seed=1)
    assert synth.state_dict_checksum(sd) == str(np.load(os.path.join(ROOT, "tests", "golden", "unet_tiny.npz"))["weight_checksum"])


def test_kernel_register_budgets_fit_their_launch_shape():
    """A kernel whose registers x threads (at the 512-register-per-warp allocation granularity) exceed the
    64 K register file fails at launch with 'too many resources' — check it at build time, without a GPU."""
    import subprocess
    from vista_b200 import lib
    lib.build()
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", lib.LIB_PATH], capture_output=True, text=True).stdout
    blocks = {"tapgemm_kernel": 320,   # instantiations with 16 epilogue warps (last template argument 4) launch 576
              "attn2_spatial_kernel": 320, "attn3_spatial_kernel": 768, "attn_spatial_kernel": 192, "attn_temporal_kernel": 128,
              "gn_stats_kernel": 256, "gn_apply_kernel": 256, "layernorm_kernel": 256}
    found = 0
    lines = out.splitlines()
    for i, line in enumerate(lines):
        for name, threads in blocks.items():
            if "Function" in line and name in line:
                m = re.search(r"REG:(\d+)", lines[i + 1])
                assert m, lines[i + 1]
                regs = int(m.group(1))
                if name == "tapgemm_kernel" and re.search(r"Li4EEEv", line):
                    threads = 576
                per_warp = -(-regs * 32 // 256) * 256            # register allocation unit: 256 per warp
                warps = -(-(threads // 32) // 4) * 4          # warps are allocated in groups of 4 (one per SM sub-partition)
                assert per_warp * warps <= 65536, (name, regs, threads)
                found += 1
    assert found >= 7


def test_engine_from_yaml_config_exposes_reference_state_dict_keys():
    """configs/inference/vista_b200.yaml instantiates through the reference's own `target:` mechanism and its
    state_dict carries the checkpoint key names of SURVEY.md Appendix D (model.diffusion_model.*, first_stage_model.decoder.*)."""
    import yaml
    from vista_b200.diffusion import instantiate_from_config
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "inference", "vista_b200.yaml")))["model"]
    p = cfg["params"]
    p["network_config"]["params"].update(model_channels=64, channel_mult=[1, 2], num_res_blocks=1, attention_resolutions=[1, 2])
    p["first_stage_config"]["params"]["decoder_config"]["params"].update(ch=64, ch_mult=[1, 2], num_res_blocks=1)
    eng = instantiate_from_config(cfg)
    keys = set(eng.state_dict())
    unet_keys = {"model.diffusion_model." + k for k in spec.unet_param_specs(spec.unet_preset("tiny"))}
    dec_keys = {"first_stage_model.decoder." + k for k in spec.decoder_param_specs(spec.decoder_preset("tiny"))}
    assert keys == unet_keys | dec_keys
    assert eng.scale_factor == 0.18215 and eng.en_and_decode_n_samples_a_time == 14
    with eng.ema_scope("x"):
        pass
    with pytest.raises(NotImplementedError):
        eng.encode_first_stage(torch.zeros(1, 3, 8, 8))
    from vista_b200.diffusion import EulerEDMSampler, VanillaCFG
    assert isinstance(eng.sampler, EulerEDMSampler) and isinstance(eng.sampler.guider, VanillaCFG)


def test_launch_tape_records_and_replays_in_order():
    """vista_b200.lib launch tape: C-ABI calls issued through the recording proxy and host-side operations
    (collectives) are replayed in their original order; a failing replayed call raises."""
    from vista_b200 import lib
    log = []
    lib.begin_tape()
    assert lib.taping()
    l = lib.load()                                   # recording proxy while a tape is open
    assert l.b200v_groupnorm_chunk() >= 8            # queries are not recorded
    lib.tape_host(lambda: log.append("h1"))
    rc = l.b200v_layernorm(None, 8, None, 8, 1, 8, None, None, 1e-5, None, 0, 1, 1, None)    # null pointers: rc != 0, no launch
    assert rc != 0
    lib.tape_host(lambda: log.append("h2"))
    tape = lib.end_tape()
    assert not lib.taping() and len(tape) == 3 and log == ["h1", "h2"]
    with pytest.raises(RuntimeError):                # the recorded failing call fails again on replay, after h1
        lib.replay(tape)
    assert log == ["h1", "h2", "h1"]
    lib.replay([tape[0], tape[2]])
    assert log == ["h1", "h2", "h1", "h1", "h2"]


def test_decode_chunk_plan_matches_reference_loop():
    """vae._decode_chunks restates the chunk walk of DiffusionEngine.decode_first_stage (vwm/models/diffusion.py:150-180):
    context = previous `overlap` frames + the next n_samples - overlap frames, written at pos - overlap, the first
    `overlap` frames of every chunk but the first averaged with what is there."""
    from vista_b200.vae import _decode_chunks

    def reference_walk(F, n_samples, overlap):
        frames = list(range(F))
        out = []
        if overlap < n_samples:
            prev, pos, first = frames[:overlap], overlap, True
            rest = frames[overlap:]
            step = n_samples - overlap
            for i in range(0, len(rest), step):
                cur = rest[i:i + step]
                ctx = prev + cur
                prev = cur[-overlap:]
                out.append((ctx[0], len(ctx), pos - overlap, 0 if first else overlap))
                pos += len(cur)
                first = False
        else:
            for i in range(0, F, n_samples):
                cur = frames[i:i + n_samples]
                out.append((cur[0], len(cur), i, 0))
        return out

    for F, n, ov in [(25, 14, 3), (25, 25, 3), (8, 4, 1), (14, 14, 3), (30, 14, 3), (5, 2, 3)]:
        assert _decode_chunks(F, n, ov) == reference_walk(F, n, ov), (F, n, ov)
    assert _decode_chunks(25, 14, 3) == [(0, 14, 0, 0), (11, 14, 11, 3)]


def test_sampler_sees_through_the_reference_closure():
    """sample_utils.do_sample wraps the engine in a local closure (sample_utils.py:314-315); our sampler recovers the
    (Denoiser, B200Wrapper) pair from it so that the unmodified caller reaches the fused loop; foreign closures pass."""
    from vista_b200 import spec
    from vista_b200.diffusion import B200Denoiser, Denoiser, _unwrap_reference_closure
    from vista_b200.modules import B200Wrapper, VideoUNet
    cfg = spec.unet_preset("tiny")
    unet = VideoUNet(in_channels=cfg.in_channels, model_channels=cfg.model_channels, out_channels=cfg.out_channels,
                     num_res_blocks=cfg.num_res_blocks, attention_resolutions=list(cfg.attention_resolutions),
                     channel_mult=list(cfg.channel_mult), num_head_channels=64, num_classes="sequential",
                     context_dim=cfg.context_dim, adm_in_channels=cfg.adm_in_channels, extra_ff_mix_layer=True,
                     use_spatial_context=True, merge_strategy="learned_with_images", video_kernel_size=[3, 1, 1],
                     use_linear_in_transformer=True, action_control=True)

    class Engine:                      # the attributes do_sample touches
        pass
    model = Engine()
    model.model = B200Wrapper(unet)
    model.denoiser = Denoiser({"target": "vista_b200.diffusion.VScalingWithEDMcNoise"}, num_frames=25)

    def denoiser(x, sigma, cond, cond_mask):          # verbatim shape of the reference's closure
        return model.denoiser(model.model, x, sigma, cond, cond_mask)

    got = _unwrap_reference_closure(denoiser)
    assert isinstance(got, B200Denoiser) and got.network is model.model and got.denoiser is model.denoiser
    other = lambda x, sigma, cond, cond_mask: x
    assert _unwrap_reference_closure(other) is other
    k = 3
    assert _unwrap_reference_closure(lambda x: x * k)(2) == 6
