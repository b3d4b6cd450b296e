"""Host-side executors (vae.DecoderRuntime / EncoderRuntime) run on CPU tensors against an emulation of the C-ABI
operators (tests/fake_ops.py): checks the orchestration — buffer reuse, weight repacking, tap and gather conventions,
residual / blend wiring — against the REAL reference's fixtures without a GPU.  The decoder is the control (its GPU
parity is established): if the emulated decoder matches, the emulation is faithful, and the encoder executor, which has
not run on hardware yet, is checked by the same means."""
import pytest
import torch

from fake_ops import patched_ops
from helpers import decoder_weights, golden, rel_l2, to_t
from vista_b200 import spec, synth


def test_decoder_executor_on_emulated_ops_matches_reference():
    from vista_b200.vae import DecoderRuntime, decode_first_stage
    cfg, sd = decoder_weights("tiny")
    z = torch.from_numpy(synth.normal(9, "decfs.z", (25, cfg.z_channels, 8, 16), std=0.18215))
    with patched_ops(), torch.no_grad():
        rt = DecoderRuntime(cfg, to_t(sd), "cpu")
        out = decode_first_stage(rt, z)
    r = rel_l2(out, torch.from_numpy(golden("decode_first_stage_tiny")["out"]))
    assert r < 5e-3, r


@pytest.mark.parametrize("name,preset,h,w,n", [("encoder_tiny", "tiny", 32, 64, 5), ("encoder_small", "small", 64, 128, 3)])
def test_encoder_executor_on_emulated_ops_matches_reference(name, preset, h, w, n):
    from vista_b200.vae import EncoderRuntime, encode_first_stage
    g = golden(name)
    cfg = spec.encoder_preset(preset)
    sd = synth.synth_state_dict(spec.encoder_param_specs(cfg), seed=3)
    x = torch.from_numpy(synth.normal(11, "enc.x", (n, cfg.in_channels, h, w), std=0.5))
    noise = torch.from_numpy(synth.normal(12, "enc.noise", tuple(g["z"].shape), std=1.0))
    with patched_ops(), torch.no_grad():
        rt = EncoderRuntime(cfg, to_t(sd), "cpu")
        z = encode_first_stage(rt, x, n_samples=int(g["n_chunk"]), noise=noise)
    r = rel_l2(z, torch.from_numpy(g["z"]))
    assert r < 5e-3, r
