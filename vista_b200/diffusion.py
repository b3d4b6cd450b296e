"""Diffusion mechanics with the reference's public surface (names, constructor arguments, call
signatures, error behaviour) — so that ``sample_utils.do_sample`` / ``DiffusionEngine.sample`` style
callers work unchanged — plus the fused B200 loop.

Mirrors (paths relative to the reference root):
  EDMDiscretization ........ vwm/modules/diffusionmodules/discretizer.py:15-37
  VScalingWithEDMcNoise & co vwm/modules/diffusionmodules/denoiser_scaling.py
  Denoiser ................. vwm/modules/diffusionmodules/denoiser.py:10-35
  VanillaCFG / Identity / Linear / TrianglePredictionGuider ... guiders.py
  EulerEDMSampler .......... vwm/modules/diffusionmodules/sampling.py:15-124
  instantiate_from_config .. vwm/util.py:154-173

The generic path keeps the reference's step algebra in torch (a dozen tiny fp32 ops per step on a
(25,4,h,w) latent); the network call is where the time goes and that is the B200 executor.  When the
network is a ``B200Wrapper`` the sampler switches to the fused loop (``fused_sample``): two small CUDA
kernels per step around the UNet, no host synchronisation, no per-step tensor allocation.
"""
from __future__ import annotations

import importlib
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------------------------
# config plumbing
# ----------------------------------------------------------------------------------------------
def get_obj_from_str(string: str):
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    d = target_dims - x.ndim
    if d < 0:
        raise ValueError(f"Input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * d]


# ----------------------------------------------------------------------------------------------
# discretisation / scalings
# ----------------------------------------------------------------------------------------------
class EDMDiscretization:
    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 80.0, rho: float = 7.0):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho

    def get_sigmas(self, n: int, device="cpu") -> torch.Tensor:
        ramp = torch.linspace(0, 1, n, device=device)
        min_inv_rho = self.sigma_min ** (1 / self.rho)
        max_inv_rho = self.sigma_max ** (1 / self.rho)
        return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** self.rho

    def __call__(self, n: int, do_append_zero: bool = True, device="cpu", flip: bool = False) -> torch.Tensor:
        sigmas = self.get_sigmas(n, device=device)
        if do_append_zero:
            sigmas = torch.cat((sigmas, sigmas.new_zeros([1])))
        return sigmas if not flip else torch.flip(sigmas, (0,))


class EDMScaling:
    def __init__(self, sigma_data: float = 0.5):
        self.sigma_data = sigma_data

    def __call__(self, sigma):
        c_skip = self.sigma_data ** 2 / (sigma ** 2 + self.sigma_data ** 2)
        c_out = sigma * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out, c_in, 0.25 * sigma.log()


class EpsScaling:
    def __call__(self, sigma):
        return torch.ones_like(sigma), -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()


class VScaling:
    def __call__(self, sigma):
        return 1.0 / (sigma ** 2 + 1.0), -sigma / (sigma ** 2 + 1.0) ** 0.5, 1.0 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()


class VScalingWithEDMcNoise:
    def __call__(self, sigma):
        c_skip = 1.0 / (sigma ** 2 + 1.0)
        c_out = -sigma / (sigma ** 2 + 1.0) ** 0.5
        c_in = 1.0 / (sigma ** 2 + 1.0) ** 0.5
        c_noise = 0.25 * sigma.log()
        return c_skip, c_out, c_in, c_noise


class Denoiser(nn.Module):
    def __init__(self, scaling_config: Dict, num_frames: int = 25):
        super().__init__()
        self.scaling = instantiate_from_config(scaling_config)
        self.num_frames = num_frames

    def possibly_quantize_sigma(self, sigma):
        return sigma

    def possibly_quantize_c_noise(self, c_noise):
        return c_noise

    def forward(self, network: nn.Module, noised_input: torch.Tensor, sigma: torch.Tensor, cond: Dict,
                cond_mask: torch.Tensor):
        sigma = self.possibly_quantize_sigma(sigma)
        sigma_shape = sigma.shape
        sigma = append_dims(sigma, noised_input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(sigma_shape))
        return network(noised_input * c_in, c_noise, cond, cond_mask, self.num_frames) * c_out + noised_input * c_skip


# ----------------------------------------------------------------------------------------------
# guiders
# ----------------------------------------------------------------------------------------------
class Guider:
    additional_cond_keys: List[str] = []

    def scale_vector(self, num_frames: int) -> torch.Tensor:
        """Per-frame guidance scale (fused path)."""
        raise NotImplementedError

    def _merge(self, c, uc):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"] + list(self.additional_cond_keys):
                c_out[k] = torch.cat((uc[k], c[k]), 0)
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return c_out


class VanillaCFG(Guider):
    def __init__(self, scale: float):
        self.scale = scale

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        return x_u + self.scale * (x_c - x_u)

    def prepare_inputs(self, x, s, c, cond_mask, uc):
        return torch.cat([x] * 2), torch.cat([s] * 2), self._merge(c, uc), torch.cat([cond_mask] * 2)

    def scale_vector(self, num_frames):
        return torch.full((num_frames,), float(self.scale))


class IdentityGuider(Guider):
    def __call__(self, x, sigma):
        return x

    def prepare_inputs(self, x, s, c, cond_mask, uc):
        return x, s, {k: c[k] for k in c}, cond_mask


class LinearPredictionGuider(Guider):
    def __init__(self, num_frames: int = 25, max_scale: float = 2.5, min_scale: float = 1.0,
                 additional_cond_keys: Optional[Union[List[str], str]] = None):
        self.min_scale, self.max_scale, self.num_frames = min_scale, max_scale, num_frames
        self.scale = torch.linspace(min_scale, max_scale, num_frames).unsqueeze(0)
        keys = additional_cond_keys or list()
        self.additional_cond_keys = [keys] if isinstance(keys, str) else list(keys)

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        T = self.num_frames
        shp = x_u.shape
        x_u = x_u.reshape(shp[0] // T, T, *shp[1:])
        x_c = x_c.reshape(shp[0] // T, T, *shp[1:])
        scale = append_dims(self.scale.expand(x_u.shape[0], T), x_u.ndim).to(x_u.device)
        return (x_u + scale * (x_c - x_u)).reshape(shp)

    def prepare_inputs(self, x, s, c, cond_mask, uc):
        return torch.cat([x] * 2), torch.cat([s] * 2), self._merge(c, uc), torch.cat([cond_mask] * 2)

    def scale_vector(self, num_frames):
        assert num_frames == self.num_frames
        return self.scale[0].clone()


class TrianglePredictionGuider(LinearPredictionGuider):
    def __init__(self, num_frames: int = 25, max_scale: float = 2.5, min_scale: float = 1.0, period=1.0,
                 period_fusing: str = "max", additional_cond_keys=None):
        super().__init__(num_frames, max_scale, min_scale, additional_cond_keys)
        values = torch.linspace(0, 1, num_frames)
        periods = [period] if isinstance(period, float) else list(period)
        scales = [self.triangle_wave(values, p) for p in periods]
        if period_fusing == "mean":
            scale = sum(scales) / len(periods)
        elif period_fusing == "multiply":
            scale = torch.prod(torch.stack(scales), dim=0)
        elif period_fusing == "max":
            scale = torch.max(torch.stack(scales), dim=0).values
        else:
            raise NotImplementedError
        self.scale = (scale * (max_scale - min_scale) + min_scale).unsqueeze(0)

    @staticmethod
    def triangle_wave(values, period):
        return 2 * (values / period - torch.floor(values / period + 0.5)).abs()


# ----------------------------------------------------------------------------------------------
# sampler
# ----------------------------------------------------------------------------------------------
class B200Denoiser:
    """Callable handed to the sampler by our engine: same ``(x, sigma, c, cond_mask)`` signature as the
    reference's lambda (sample_utils.py:314-315), but it also exposes the pieces so that the sampler can
    run the fused loop."""

    def __init__(self, denoiser: Denoiser, network: nn.Module):
        self.denoiser, self.network = denoiser, network

    def __call__(self, x, sigma, c, cond_mask):
        return self.denoiser(self.network, x, sigma, c, cond_mask)


def _unwrap_reference_closure(fn):
    """The reference's callers hand the sampler a local closure, ``def denoiser(x, sigma, cond, cond_mask): return
    model.denoiser(model.model, x, sigma, cond, cond_mask)`` (sample_utils.py:314-315, diffusion.py:324-325), which
    hides the engine.  When that engine's network is a B200Wrapper and its denoiser is ours, recover the pair so that an
    UNMODIFIED ``do_sample`` / ``DiffusionEngine.sample`` runs the fused loop; anything else is returned untouched."""
    if isinstance(fn, B200Denoiser) or not callable(fn):
        return fn
    cells = getattr(fn, "__closure__", None) or ()
    from .modules import B200Wrapper
    for cell in cells:
        try:
            obj = cell.cell_contents
        except ValueError:          # empty cell
            continue
        net, den = getattr(obj, "model", None), getattr(obj, "denoiser", None)
        if isinstance(net, B200Wrapper) and isinstance(den, Denoiser):
            return B200Denoiser(den, net)
    return fn


class BaseDiffusionSampler:
    def __init__(self, discretization_config, num_steps: Optional[int] = None, guider_config=None,
                 verbose: bool = False, device: str = "cuda"):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(guider_config)
        self.verbose = verbose
        self.device = device

    def prepare_sampling_loop(self, x, cond, uc=None, num_steps=None):
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device=self.device)
        uc = cond if uc is None else uc
        x *= torch.sqrt(1.0 + sigmas[0] ** 2)        # in place on the caller's tensor, as the reference does
        num_sigmas = len(sigmas)
        s_in = x.new_ones([x.shape[0]])
        return x, s_in, sigmas, num_sigmas, cond, uc

    def denoise(self, x, denoiser, sigma, cond, cond_mask, uc):
        denoised = denoiser(*self.guider.prepare_inputs(x, sigma, cond, cond_mask, uc))
        return self.guider(denoised, sigma)

    def get_sigma_gen(self, num_sigmas):
        gen = range(num_sigmas - 1)
        if self.verbose:
            from tqdm import tqdm
            gen = tqdm(gen, total=num_sigmas, desc=f"Sampling with {self.__class__.__name__} for {num_sigmas} steps")
        return gen


class EulerEDMSampler(BaseDiffusionSampler):
    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise

    def euler_step(self, x, d, dt):
        return x + dt * d

    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, cond_mask=None, uc=None, gamma=0.0):
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            eps = torch.randn_like(x) * self.s_noise
            x = x + eps * append_dims(sigma_hat ** 2 - sigma ** 2, x.ndim) ** 0.5
        denoised = self.denoise(x, denoiser, sigma_hat, cond, cond_mask, uc)
        d = (x - denoised) / append_dims(sigma_hat, x.ndim)
        dt = append_dims(next_sigma - sigma_hat, x.ndim)
        return self.euler_step(x, d, dt)

    def __call__(self, denoiser, x, cond, uc=None, cond_frame=None, cond_mask=None, num_steps=None):
        denoiser = _unwrap_reference_closure(denoiser)
        if isinstance(denoiser, B200Denoiser) and self.s_churn == 0.0 and self._fusable(denoiser, cond, uc):
            from .fused import fused_sample
            return fused_sample(self, denoiser, x, cond, uc, cond_frame, cond_mask, num_steps)
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        replace_cond_frames = cond_mask is not None and bool(cond_mask.any())
        for i in self.get_sigma_gen(num_sigmas):
            if replace_cond_frames:
                x = x * append_dims(1 - cond_mask, x.ndim) + cond_frame * append_dims(cond_mask, cond_frame.ndim)
            gamma = (min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1)
                     if self.s_tmin <= sigmas[i] <= self.s_tmax else 0.0)
            x = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], denoiser, x, cond, cond_mask, uc, gamma)
        if replace_cond_frames:
            x = x * append_dims(1 - cond_mask, x.ndim) + cond_frame * append_dims(cond_mask, cond_frame.ndim)
        return x

    def _fusable(self, denoiser: "B200Denoiser", cond, uc) -> bool:
        from .modules import B200Wrapper
        return (isinstance(denoiser.network, B200Wrapper) and isinstance(denoiser.denoiser.scaling, VScalingWithEDMcNoise)
                and isinstance(self.guider, (VanillaCFG, LinearPredictionGuider))
                and all(k in cond for k in ("crossattn", "vector", "concat")))
