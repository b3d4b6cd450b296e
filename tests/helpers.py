"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from vista_b200 import spec, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    return np.load(path, allow_pickle=False)


def has_golden(name):
    return os.path.isfile(os.path.join(GOLDEN, name + ".npz"))


def to_t(d, device="cpu"):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in d.items()}


def unet_weights(preset, seed=1):
    cfg = spec.unet_preset(preset)
    sd = synth.synth_state_dict(spec.unet_param_specs(cfg), seed=seed)
    return cfg, sd


def decoder_weights(preset, seed=2):
    cfg = spec.decoder_preset(preset)
    sd = synth.synth_state_dict(spec.decoder_param_specs(cfg), seed=seed)
    return cfg, sd


def unet_inputs(seed, cfg, h, w, T, sigma=5.0, n_cond=1):
    """Same construction as oracle/make_golden.py:unet_inputs."""
    c, uc = synth.synth_conditioning(seed, T, h, w, trajectory=True, context_dim=cfg.context_dim,
                                     adm=cfg.adm_in_channels)
    noise, z, mask = synth.synth_latents(seed, T, h, w)
    mask[:n_cond] = 1.0
    x = np.concatenate([noise, noise], 0) * np.float32(sigma)
    cc = {k: np.concatenate([uc[k], c[k]], 0) for k in c}
    mask2 = np.concatenate([mask, mask], 0)
    return x, cc, mask2


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
