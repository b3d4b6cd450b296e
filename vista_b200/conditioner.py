"""The conditioner's VAE branch on the B200 encoder.

``VideoPredictionEmbedderWithEncoder`` mirrors vwm/modules/encoders/modules.py:428-502 (same constructor keywords, same
``forward`` / ``skip_encode`` behaviour) so that the ``cond_frames`` entry of ``conditioner_config`` (vista.yaml:68-96) can
name it; ``AutoencoderKLModeOnly`` mirrors vwm/models/autoencoder.py:432-528 for the one thing that embedder calls,
``encode(x)`` = mode of the posterior after ``quant_conv``.  Checkpoint keys are the reference's
(``...encoder.encoder.*``, ``...encoder.quant_conv.*``; the reference also carries an unused decoder + post_quant_conv
there, which ``load_state_dict(strict=False)`` — what sample_utils.py:72 uses — skips).  The encoder runs on
``vista_b200.vae.EncoderRuntime`` with ``quant_conv`` (1x1, 8 -> 8) folded into the 3x3 ``conv_out`` weights at packing
time: no extra pass, no torch compute on the path.  CLIP (the other image branch of the conditioner) stays out of scope."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .diffusion import instantiate_from_config
from .modules import register_param_tree
from .vae import Encoder, EncoderRuntime


class AutoencoderKLModeOnly(nn.Module):
    def __init__(self, embed_dim: int, ddconfig: Dict, **reference_only):
        super().__init__()
        known = {"monitor", "loss_config", "ckpt_path", "ckpt_engine", "max_batch_size", "lr_g_factor", "input_key",
                 "optimizer_config", "regularizer_config", "trainable_ae_params", "ae_optimizer_args", "ema_decay"}
        unknown = sorted(set(reference_only) - known)
        if unknown:
            raise TypeError(f"vista_b200.conditioner.AutoencoderKLModeOnly: unexpected keyword(s) {unknown}")
        dd = dict(ddconfig)
        if dd.get("attn_type") == "vanilla-xformers":      # same arithmetic (model.py:199-227 vs :158-170), our kernel either way
            dd["attn_type"] = "vanilla"
        if not dd.get("double_z", True):
            raise NotImplementedError("AutoencoderKLModeOnly needs double_z (mean | logvar moments)")
        self.encoder = Encoder(**dd)
        zc = self.encoder.b200_config.z_channels
        self.embed_dim = embed_dim
        register_param_tree(self, {"quant_conv.weight": ((2 * embed_dim, 2 * zc, 1, 1), "w"), "quant_conv.bias": ((2 * embed_dim,), "b")})
        self._runtime = None
        self.register_load_state_dict_post_hook(lambda module, keys: setattr(module, "_runtime", None))

    def _apply(self, fn, *args, **kwargs):
        before = tuple((p.device, p.dtype) for p in self.parameters())
        out = super()._apply(fn, *args, **kwargs)
        if tuple((p.device, p.dtype) for p in self.parameters()) != before:
            self._runtime = None
        return out

    def runtime(self, device) -> EncoderRuntime:
        if torch.device(device).type != "cuda":
            raise RuntimeError("vista_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self._runtime is None:
            qw = self.get_parameter("quant_conv.weight").detach().float().flatten(1)       # [2e, 2z]
            qb = self.get_parameter("quant_conv.bias").detach().float()
            self._runtime = EncoderRuntime(self.encoder.b200_config, self.encoder.state_dict(), device, post=(qw, qb))
        return self._runtime

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_reg_log: bool = False):
        """autoencoder.py:467-488 with the DiagonalGaussianRegularizer in mode (sample=False): (n,3,H,W) -> (n,embed_dim,H/8,W/8)."""
        rt = self.runtime(x.device)
        n, cin, H, W = x.shape
        down = 2 ** (len(rt.cfg.ch_mult) - 1)
        tok = rt.buf("e.x", n * H * W, 8)
        tok.zero_()
        ops.nchw_to_tokens(x.float().contiguous(), tok, n, cin, H, W)
        mom_tok = rt.forward(tok, n, H, W)
        mom = torch.empty(n, 8, H // down, W // down, dtype=torch.float32, device=x.device)
        ops.tokens_to_nchw(mom_tok, mom, n, 8, H // down, W // down)
        z = mom[:, : self.embed_dim]
        return (z, {}) if return_reg_log else z


class VideoPredictionEmbedderWithEncoder(nn.Module):
    """encoders/modules.py:428-502.  ``forward(vid)``: latents pass through when ``skip_encode`` is set (the rollout's
    re-conditioning, sample_utils.py:345-350); otherwise optional noise augmentation, the encoder in chunks of
    ``en_and_decode_n_samples_a_time``, ``* scale_factor``, "(b t) c h w -> b () (t c) h w" and ``n_copies`` repeats."""

    def __init__(self, n_cond_frames: int, n_copies: int, encoder_config: dict, sigma_sampler_config: Optional[dict] = None,
                 sigma_cond_config: Optional[dict] = None, is_ae: bool = False, scale_factor: float = 1.0,
                 disable_encoder_autocast: bool = False, en_and_decode_n_samples_a_time: Optional[int] = None):
        super().__init__()
        self.n_cond_frames, self.n_copies = n_cond_frames, n_copies
        self.encoder = instantiate_from_config(encoder_config)
        self.sigma_sampler = instantiate_from_config(sigma_sampler_config) if sigma_sampler_config is not None else None
        self.sigma_cond = instantiate_from_config(sigma_cond_config) if sigma_cond_config is not None else None
        self.is_ae, self.scale_factor = is_ae, scale_factor
        self.disable_encoder_autocast = disable_encoder_autocast
        self.en_and_decode_n_samples_a_time = en_and_decode_n_samples_a_time
        self.skip_encode = False
        # AbstractEmbModel attributes the GeneralConditioner sets / reads (encoders/modules.py:31-74)
        self.is_trainable, self.ucg_rate, self.input_key = False, 0.0, None

    def forward(self, vid: torch.Tensor, noise: Optional[torch.Tensor] = None):
        if self.skip_encode:
            return vid
        sigma_cond = None
        if self.sigma_sampler is not None:
            bs = vid.shape[0] // self.n_cond_frames
            sigmas = self.sigma_sampler(bs).to(vid.device)
            if self.sigma_cond is not None:
                sigma_cond = self.sigma_cond(sigmas).repeat_interleave(self.n_copies, dim=0)
            sigmas = sigmas.repeat_interleave(self.n_cond_frames, dim=0)
            noise = torch.randn_like(vid) if noise is None else noise
            vid = vid + noise * sigmas.reshape(-1, *([1] * (vid.ndim - 1)))
        n_samples = self.en_and_decode_n_samples_a_time or vid.shape[0]
        outs = []
        for i in range(math.ceil(vid.shape[0] / n_samples)):
            chunk = vid[i * n_samples:(i + 1) * n_samples]
            outs.append(self.encoder.encode(chunk) if self.is_ae else self.encoder(chunk))
        out = torch.cat(outs, dim=0) * self.scale_factor
        bt, c, h, w = out.shape
        out = out.reshape(bt // self.n_cond_frames, self.n_cond_frames * c, h, w)          # "(b t) c h w -> b () (t c) h w"
        out = out.repeat_interleave(self.n_copies, dim=0)                                 # "b 1 c h w -> (b t) c h w"
        return (out, sigma_cond) if sigma_cond is not None else out
