"""Pins oracle/vista_oracle.py (the CPU restatement) against outputs of the REAL reference
modules (tests/golden/*.npz, produced by oracle/make_golden.py in the build container) and
against the closed-form anchors of SURVEY.md §8c.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import vista_oracle as vo
from vista_b200 import spec, synth

from helpers import (decoder_weights, golden, has_golden, rel_l2, to_t, unet_inputs, unet_weights)

UNET_CASES = {"unet_tiny": ("tiny", 8, 16, 25), "unet_small": ("small", 16, 32, 25)}


def test_anchors():
    g = golden("anchors")
    for n in (1, 3, 10, 50):
        assert torch.allclose(vo.edm_sigmas(n), torch.from_numpy(g[f"sigmas_{n}"]), rtol=1e-6, atol=0)
    s3 = vo.edm_sigmas(3)
    assert abs(float(s3[0]) - 700.0) < 1e-3 and abs(float(s3[1]) - 15.590) < 1e-3 and float(s3[3]) == 0.0
    s = torch.tensor([700.0, 15.59, 1.0, 0.002])
    assert torch.allclose(torch.stack(vo.vscaling_edm_cnoise(s)), torch.from_numpy(g["vscaling"]), rtol=1e-6)
    assert torch.allclose(vo.triangle_scales(25)[None], torch.from_numpy(g["triangle_25"]), rtol=1e-6)
    assert torch.allclose(vo.guider_scales("LinearPredictionGuider", 25, 2.5)[None],
                          torch.from_numpy(g["linear_25"]), rtol=1e-6)
    t = torch.tensor([0.25 * np.log(700.0), -1.5, 0.0], dtype=torch.float32)
    assert torch.allclose(vo.timestep_embedding(t, 320), torch.from_numpy(g["temb_320"]), atol=1e-6)
    assert torch.allclose(vo.timestep_embedding(torch.arange(25), 64), torch.from_numpy(g["temb_frames_64"]), atol=1e-6)


def test_param_inventory_counts():
    # 1 648 079 082 UNet params in 1496 tensors; decoder 63 579 183
    s = spec.unet_param_specs(spec.unet_preset("vista"))
    assert len(s) == 1496
    assert sum(int(np.prod(v[0])) for v in s.values()) == 1648079082
    d = spec.decoder_param_specs(spec.decoder_preset("vista"))
    assert sum(int(np.prod(v[0])) for v in d.values()) == 63579183


@pytest.mark.parametrize("name", list(UNET_CASES))
def test_unet_forward_matches_reference(name):
    if not has_golden(name):
        pytest.skip("fixture not generated")
    preset, h, w, T = UNET_CASES[name]
    g = golden(name)
    cfg, sd = unet_weights(preset)
    assert synth.state_dict_checksum(sd) == str(g["weight_checksum"])
    x, cc, mask2 = unet_inputs(7, cfg, h, w, T)
    assert synth.checksum([x, mask2] + [cc[k] for k in sorted(cc)]) == str(g["input_checksum"])
    sdt = to_t(sd)
    sigma = torch.full((2 * T,), 5.0)
    with torch.no_grad():
        out = vo.denoise(sdt, cfg, torch.from_numpy(x), sigma, to_t(cc), torch.from_numpy(mask2), T)
    ref = torch.from_numpy(g["denoised"])
    assert rel_l2(out, ref) < 2e-5, rel_l2(out, ref)
    assert float((out - ref).abs().max()) < 2e-4


@pytest.mark.parametrize("name,steps,guider,n_cond", [("sampler_tiny_cfg", 4, "VanillaCFG", 1),
                                                     ("sampler_tiny_triangle", 3, "TrianglePredictionGuider", 3)])
def test_sampler_matches_reference(name, steps, guider, n_cond):
    g = golden(name)
    cfg, sd = unet_weights("tiny")
    assert synth.state_dict_checksum(sd) == str(g["weight_checksum"])
    T, h, w = 25, 8, 16
    c, uc = synth.synth_conditioning(7, T, h, w, trajectory=True, context_dim=cfg.context_dim, adm=cfg.adm_in_channels)
    noise, z, mask = synth.synth_latents(7, T, h, w)
    mask[:n_cond] = 1.0
    with torch.no_grad():
        out = vo.euler_edm_sample(to_t(sd), cfg, torch.from_numpy(noise), to_t(c), to_t(uc), torch.from_numpy(z),
                                  torch.from_numpy(mask), steps, T, guider=guider, scale=2.5)
    ref = torch.from_numpy(g["sample"])
    assert rel_l2(out, ref) < 5e-5, rel_l2(out, ref)


@pytest.mark.parametrize("name,preset", [("decoder_tiny", "tiny"), ("decoder_small", "small")])
def test_decoder_matches_reference(name, preset):
    if not has_golden(name):
        pytest.skip("fixture not generated")
    g = golden(name)
    cfg, sd = decoder_weights(preset)
    assert synth.state_dict_checksum(sd) == str(g["weight_checksum"])
    z = synth.normal(9, "dec.z", (14, cfg.z_channels, 8, 16), std=1.0)
    with torch.no_grad():
        out = vo.decoder_forward(to_t(sd), cfg, torch.from_numpy(z), 14)
    ref = torch.from_numpy(g["out"])
    assert rel_l2(out, ref) < 2e-5, rel_l2(out, ref)


def test_decode_first_stage_matches_reference():
    g = golden("decode_first_stage_tiny")
    cfg, sd = decoder_weights("tiny")
    z = synth.normal(9, "decfs.z", (25, cfg.z_channels, 8, 16), std=0.18215)
    with torch.no_grad():
        out = vo.decode_first_stage(to_t(sd), cfg, torch.from_numpy(z))
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape == (25, 3, 16, 32)
    assert rel_l2(out, ref) < 2e-5, rel_l2(out, ref)


# ---- VAE encoder: SURVEY.md §8f rank 1 (the next row) — oracle pinned before any kernel work ----
@pytest.mark.parametrize("name,preset,h,w,n", [("encoder_tiny", "tiny", 32, 64, 5), ("encoder_small", "small", 64, 128, 3)])
def test_encoder_matches_reference(name, preset, h, w, n):
    if not has_golden(name):
        pytest.skip("fixture not generated")
    g = golden(name)
    cfg = spec.encoder_preset(preset)
    sd = synth.synth_state_dict(spec.encoder_param_specs(cfg), seed=3)
    assert synth.state_dict_checksum(sd) == str(g["weight_checksum"])
    x = torch.from_numpy(synth.normal(11, "enc.x", (n, cfg.in_channels, h, w), std=0.5))
    with torch.no_grad():
        mom = vo.encoder_forward(to_t(sd), cfg, x)
    ref = torch.from_numpy(g["moments"])
    assert mom.shape == ref.shape == (n, 2 * cfg.z_channels, h // 2 ** (len(cfg.ch_mult) - 1), w // 2 ** (len(cfg.ch_mult) - 1))
    assert rel_l2(mom, ref) < 2e-5, rel_l2(mom, ref)
    # encode_first_stage: chunked, sampled with the recorded noise (the reference draws it from the device RNG), scaled
    noise = torch.from_numpy(synth.normal(12, "enc.noise", (n, cfg.z_channels, mom.shape[2], mom.shape[3]), std=1.0))
    with torch.no_grad():
        z = vo.encode_first_stage(to_t(sd), cfg, x, n_samples=int(g["n_chunk"]), noise=noise)
        z_mode = vo.encode_first_stage(to_t(sd), cfg, x)
    assert rel_l2(z, torch.from_numpy(g["z"])) < 2e-5
    assert torch.allclose(z_mode, ref[:, :cfg.z_channels] * 0.18215, rtol=1e-4, atol=1e-6)


def test_encoder_param_inventory():
    # 34.2 M parameters
    e = spec.encoder_param_specs(spec.encoder_preset("vista"))
    assert len(e) == 106 and sum(int(np.prod(v[0])) for v in e.values()) == 34163592
