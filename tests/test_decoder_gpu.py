"""VAE decoder parity on the GPU against the REAL reference's fp32 outputs (tests/golden).
The reference decodes in fp32; our convolutions use fp16 operands with fp32 accumulation, so the
stated tolerance is the fp16 one: rel-L2 <= 5e-3 on the decoded frames (values in about [-1, 1])."""
import os

import pytest
import torch

from helpers import decoder_weights, golden, has_golden, rel_l2, to_t
from vista_b200 import spec, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make(preset):
    from vista_b200.vae import VideoDecoder
    cfg, sd = decoder_weights(preset)
    with torch.device(DEV):
        dec = VideoDecoder(ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                           attn_resolutions=[], z_channels=cfg.z_channels, in_channels=3, resolution=256,
                           video_kernel_size=[3, 1, 1], attn_type="vanilla", double_z=True)
    dec.load_state_dict(to_t(sd), strict=True)
    return cfg, dec


@pytest.mark.parametrize("name,preset", [("decoder_tiny", "tiny"), ("decoder_small", "small")])
def test_decoder_forward_vs_reference(name, preset):
    if not has_golden(name):
        pytest.skip("fixture not generated")
    cfg, dec = make(preset)
    z = torch.from_numpy(synth.normal(9, "dec.z", (14, cfg.z_channels, 8, 16), std=1.0)).to(DEV)
    out = dec(z, timesteps=14)
    torch.cuda.synchronize()
    ref = torch.from_numpy(golden(name)["out"])
    r = rel_l2(out.cpu(), ref)
    print(f"{name}: rel-L2 {r:.3e} max-abs {float((out.cpu() - ref).abs().max()):.3e} ref absmean {float(ref.abs().mean()):.3f}")
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert r < 5e-3, r


# The decoder the bench times: vista.yaml architecture (ch = 128), latents of the real magnitude (std 1/0.18215 = 5.5 after
# decode_first_stage's division), fp16 tensor-core operands against the reference's fp32 decode.  The reference outputs
# are committed as strided samples + 8 x 8 block means (oracle/make_golden.py:gen_decoder_big).  Stated tolerance of the
# fp16-operand decode: rel-L2 <= 5e-3 on both; the measured figures are printed.
@pytest.mark.parametrize("name,h,w,n", [("decoder_vista_16x32", 16, 32, 14), ("decoder_vista_72x128_t5", 72, 128, 5)])
def test_decoder_vista_arch_vs_reference(name, h, w, n):
    if not has_golden(name):
        pytest.skip("fixture not generated")
    g = golden(name)
    cfg, dec = make("vista")
    z = torch.from_numpy(synth.normal(9, "decbig.z", (n, cfg.z_channels, h, w), std=1.0 / 0.18215)).to(DEV)
    out = dec(z, timesteps=n)
    torch.cuda.synchronize()
    assert out.shape == (n, 3, 8 * h, 8 * w) and torch.isfinite(out).all()
    stride = int(g["stride"])
    smp = out[:, :, ::stride, ::stride].cpu()
    bm = out.double().reshape(n, 3, h, 8, w, 8).mean(dim=(3, 5)).float().cpu()
    rs, rb = rel_l2(smp, torch.from_numpy(g["samples"])), rel_l2(bm, torch.from_numpy(g["block_means"]))
    print(f"{name}: samples rel-L2 {rs:.3e} max-abs {float((smp - torch.from_numpy(g['samples'])).abs().max()):.3e}; "
          f"block means rel-L2 {rb:.3e}; reference rms {float(g['rms']):.3f}")
    assert rs < 5e-3 and rb < 5e-3, (rs, rb)


def test_decode_first_stage_chunks_and_overlap():
    from vista_b200.vae import decode_first_stage
    cfg, dec = make("tiny")
    z = torch.from_numpy(synth.normal(9, "decfs.z", (25, cfg.z_channels, 8, 16), std=0.18215)).to(DEV)
    out = decode_first_stage(dec.runtime(DEV), z)
    torch.cuda.synchronize()
    ref = torch.from_numpy(golden("decode_first_stage_tiny")["out"])
    r = rel_l2(out.cpu(), ref)
    print(f"decode_first_stage: rel-L2 {r:.3e}")
    assert out.shape == ref.shape == (25, 3, 16, 32)
    assert r < 5e-3, r
    # frames 11..13 are the average of two chunks: they must differ from a single-chunk decode of the same frames
    single = dec(z[:14] / 0.18215, timesteps=14)
    torch.cuda.synchronize()
    assert rel_l2(single[:11].cpu(), ref[:11]) < 5e-3


# ---- VAE encoder ----
@pytest.mark.parametrize("name,preset,h,w,n", [("encoder_tiny", "tiny", 32, 64, 5), ("encoder_small", "small", 64, 128, 3)])
def test_encoder_matches_reference(name, preset, h, w, n):
    from vista_b200.vae import EncoderRuntime, encode_first_stage
    g = golden(name)
    cfg = spec.encoder_preset(preset)
    sd = synth.synth_state_dict(spec.encoder_param_specs(cfg), seed=3)
    rt = EncoderRuntime(cfg, to_t(sd), DEV)
    x = torch.from_numpy(synth.normal(11, "enc.x", (n, cfg.in_channels, h, w), std=0.5)).to(DEV)
    noise = torch.from_numpy(synth.normal(12, "enc.noise", tuple(g["z"].shape), std=1.0)).to(DEV)
    z = encode_first_stage(rt, x, n_samples=int(g["n_chunk"]), noise=noise)
    torch.cuda.synchronize()
    r = rel_l2(z.cpu(), torch.from_numpy(g["z"]))
    print(f"{name}: encode_first_stage rel-L2 {r:.3e}")
    assert r < 5e-3, r


def test_cond_frames_embedder_matches_reference():
    """conditioner.VideoPredictionEmbedderWithEncoder on the GPU (quant_conv folded into the encoder's conv_out) against the
    fixture from the REAL embedder + AutoencoderKLModeOnly (encoders/modules.py:428-502, autoencoder.py:519-528)."""
    from test_executor_cpu import _cond_embedder
    emb, x = _cond_embedder(DEV)
    out = emb(x)
    torch.cuda.synchronize()
    ref = torch.from_numpy(golden("cond_embedder_tiny")["out"])
    r = rel_l2(out.cpu(), ref)
    print(f"cond_frames embedder: rel-L2 {r:.3e}")
    assert out.shape == ref.shape and r < 5e-3, r
