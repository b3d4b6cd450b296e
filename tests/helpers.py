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


def rollout(sample_fn, z, noises, T):
    """The multi-round loop of sample_utils.do_sample (sample_utils.py:318-365) with a fixed synthetic conditioning
    (the reference re-encodes the condition frame with CLIP / the VAE between rounds: outside the path).
    sample_fn(noise, cond_frame, cond_mask) -> (T,4,h,w).  Returns samples_z ((rounds*(T-3)+3), 4, h, w)."""
    rounds = len(noises)
    init_mask, pred_mask = torch.zeros(T, device=z.device), torch.zeros(T, device=z.device)
    init_mask[0] = 1
    pred_mask[[0, 1, 2]] = 1
    samples_z = torch.zeros((rounds * (T - 3) + 3,) + tuple(z.shape[1:]), device=z.device)
    sample = sample_fn(noises[0].clone(), z, init_mask)
    sample[0] = z[0]
    samples_z[:T] = sample
    for n in range(rounds - 1):
        filled = torch.zeros_like(z)
        filled[[0, 1, 2]] = sample[-3:]
        sample = sample_fn(noises[n + 1].clone(), filled, pred_mask)
        samples_z[(n + 1) * (T - 3) + 3:(n + 1) * (T - 3) + T] = sample[3:]
    return samples_z
