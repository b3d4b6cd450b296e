"""Deterministic synthetic weights and inputs (no checkpoint / dataset is reachable offline).

Everything is drawn from numpy's Philox bit generator keyed by (seed, crc32(name)), which is
specified to be identical across platforms, so the build container (where the real reference
produced the golden fixtures) and the GPU box regenerate bit-identical tensors.  A checksum
helper lets tests assert that before comparing anything.

Shapes / conditioning layout follow SURVEY.md §8(d) "Synthetic inputs":
  c["crossattn"] (T,1,3456) = 1024 CLIP dims + 2432 action dims (configs/inference/vista.yaml:106-144),
  c["vector"] (T,768), c["concat"] (T,4,h,w) un-scaled VAE mode, uc = zeros for crossattn/concat.
"""
from __future__ import annotations

import os
import zlib
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from .spec import ParamSpec


def _rng(seed: int, name: str) -> np.random.Generator:
    key = zlib.crc32(name.encode("utf-8")) & 0xFFFFFFFF
    return np.random.Generator(np.random.Philox(key=[int(seed) & 0xFFFFFFFF, key]))


def normal(seed: int, name: str, shape: Tuple[int, ...], std: float = 1.0, mean: float = 0.0) -> np.ndarray:
    g = _rng(seed, name)
    return (g.standard_normal(size=shape, dtype=np.float32) * np.float32(std) + np.float32(mean)).astype(np.float32)


def synth_param(seed: int, name: str, spec: ParamSpec) -> np.ndarray:
    shape, kind = spec
    if kind == "w":
        fan_in = int(np.prod(shape[1:]))
        return normal(seed, name, shape, std=0.8 / np.sqrt(fan_in))
    if kind == "wz":      # reference zero-inits these (zero_module / nn.init.zeros_); see SURVEY App. G.20
        fan_in = int(np.prod(shape[1:]))
        return normal(seed, name, shape, std=0.5 / np.sqrt(fan_in))
    if kind == "b":
        return normal(seed, name, shape, std=0.05)
    if kind == "g":
        return normal(seed, name, shape, std=0.1, mean=1.0)
    if kind == "mix":     # UNet AlphaBlender logit, reference init 0.5 (video_model.py:105)
        return normal(seed, name, shape, std=0.3, mean=0.5)
    if kind == "mix0":    # VAE decoder blend logit, reference init 0.0 (temporal_ae.py:112)
        return normal(seed, name, shape, std=0.3, mean=0.0)
    raise KeyError(kind)


def synth_state_dict(specs: Dict[str, ParamSpec], seed: int = 1, prefix: str = "", threads: int = 0) -> Dict[str, np.ndarray]:
    """Every tensor has its own (seed, name) stream, so generation order / threading cannot change the values."""
    n = sum(int(np.prod(v[0])) for v in specs.values())
    if threads <= 0:
        threads = min(16, os.cpu_count() or 1) if n > 50_000_000 else 1
    if threads == 1:
        return {prefix + k: synth_param(seed, k, v) for k, v in specs.items()}
    from concurrent.futures import ThreadPoolExecutor
    keys = list(specs)
    with ThreadPoolExecutor(threads) as ex:
        vals = list(ex.map(lambda k: synth_param(seed, k, specs[k]), keys))
    return {prefix + k: v for k, v in zip(keys, vals)}


def checksum(arrays: Iterable[np.ndarray]) -> str:
    """Order-dependent crc32 over raw bytes; used to assert that two boxes built the same weights."""
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1), c)
    return f"{c & 0xFFFFFFFF:08x}"


def state_dict_checksum(sd: Dict[str, np.ndarray]) -> str:
    return checksum(sd[k] for k in sorted(sd))


def sinusoid(values: np.ndarray, dim: int, max_period: float = 10000.0) -> np.ndarray:
    """cos||sin embedding of scalars, as ConcatTimestepEmbedderND does
    (vwm/modules/encoders/modules.py:414-425 via util.timestep_embedding)."""
    half = dim // 2
    freqs = np.exp(-np.log(max_period) * np.arange(half, dtype=np.float32) / half).astype(np.float32)
    args = values.astype(np.float32)[:, None] * freqs[None]
    return np.concatenate([np.cos(args), np.sin(args)], axis=-1).astype(np.float32)


def synth_conditioning(seed: int, num_frames: int, h: int, w: int, *, trajectory: bool = False,
                       context_dim: int = 1024, adm: int = 768, zc: int = 4):
    """Returns (c, uc) dicts of numpy arrays with the layouts ``do_sample`` builds
    (sample_utils.py:255-276)."""
    clip = normal(seed, "cond.clip", (1, 1, context_dim))
    action = np.zeros((1, 1, 128 * 19), np.float32)
    if trajectory:  # config 3: 8 floats in the trajectory slot [1152:2176] -> [128:1152] of the action part
        g = _rng(seed, "cond.traj")
        traj = g.uniform(-5.0, 30.0, size=(8,)).astype(np.float32)
        action[0, 0, 128:128 + 8 * 128] = sinusoid(traj, 128).reshape(-1)
    cross = np.concatenate([clip, action], axis=-1)
    vec = normal(seed, "cond.vector", (1, adm))
    concat = normal(seed, "cond.concat", (1, zc, h, w), std=5.0)
    c = {"crossattn": np.repeat(cross, num_frames, 0), "vector": np.repeat(vec, num_frames, 0),
         "concat": np.repeat(concat, num_frames, 0)}
    uc = {"crossattn": np.zeros_like(c["crossattn"]), "vector": c["vector"].copy(),
          "concat": np.zeros_like(c["concat"])}
    return c, uc


def synth_latents(seed: int, num_frames: int, h: int, w: int, zc: int = 4):
    """(noise, cond_frame latents, cond_mask) — noise is injected, never device RNG."""
    noise = normal(seed, "latent.noise", (num_frames, zc, h, w))
    z = normal(seed, "latent.cond_frame", (num_frames, zc, h, w), std=0.9)
    mask = np.zeros((num_frames,), np.float32)
    mask[0] = 1.0
    return noise, z, mask
