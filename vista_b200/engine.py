"""``DiffusionEngine`` surface (vwm/models/diffusion.py:20-131,150-180,306-329) for the hot path: the attributes
and methods ``sample_utils`` uses (.model, .denoiser, .first_stage_model, .scale_factor, .decode_first_stage,
.sample, .ema_scope) on top of the B200 executors, plus the SURVEY §8f rows as engine calls (``encode_first_stage`` with
``encoder_config: vista_b200.vae.Encoder``, ``rollout``, ``sample_ensemble``, ``decode_first_stage_u8``).  Training is out of
scope and raises; a conditioner is hosted when a ``conditioner_config`` is given (its ``cond_frames`` embedder can be
``vista_b200.conditioner.VideoPredictionEmbedderWithEncoder``), not built otherwise."""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from .diffusion import B200Denoiser, get_obj_from_str, instantiate_from_config
from .modules import B200Wrapper
from .vae import VideoDecoder, decode_first_stage as _decode_first_stage


class FirstStage(nn.Module):
    """Holds the decoder under the reference's key prefix ``first_stage_model.decoder.*``
    (AutoencodingEngine.decode, vwm/models/autoencoder.py:206-208)."""

    def __init__(self, decoder_config: Dict, encoder_config: Optional[Dict] = None, **reference_only):
        super().__init__()
        # AutoencodingEngine keywords (vwm/models/autoencoder.py:106-125) without a meaning at inference are accepted
        known = {"loss_config", "regularizer_config", "optimizer_config", "lr_g_factor", "trainable_ae_params",
                 "ae_optimizer_args", "trainable_disc_params", "disc_optimizer_args", "disc_start_iter", "diff_boost_factor",
                 "ckpt_engine", "ckpt_path", "additional_decode_keys", "ema_decay", "monitor", "input_key"}
        unknown = sorted(set(reference_only) - known)
        if unknown:
            raise TypeError(f"vista_b200.engine.FirstStage: unexpected keyword(s) {unknown}")
        self.decoder = instantiate_from_config(decoder_config)
        # optional: the B200 encoder (vista_b200.vae.Encoder), keys
        # ``first_stage_model.encoder.*`` as in the reference checkpoint
        if encoder_config is not None:
            self.encoder = instantiate_from_config(encoder_config)

    def decode(self, z: torch.Tensor, **kwargs) -> torch.Tensor:
        return self.decoder(z, **kwargs)

    def encode(self, x, **kwargs):
        if not hasattr(self, "encoder"):
            raise NotImplementedError("no encoder_config given: the VAE encoder is the next row after the hot path "
                                      "(SURVEY.md §8f); pass latents, or add encoder_config: vista_b200.vae.Encoder")
        return self.encoder(x)                 # moments (mean | logvar); sampling lives in encode_first_stage


class DiffusionEngine(nn.Module):
    def __init__(self, network_config: Dict, denoiser_config: Dict, first_stage_config: Optional[Dict] = None,
                 conditioner_config=None, sampler_config: Optional[Dict] = None, scale_factor: float = 1.0,
                 disable_first_stage_autocast: bool = False, en_and_decode_n_samples_a_time: Optional[int] = None,
                 num_frames: int = 25, network_wrapper: Optional[str] = None, replace_cond_frames: bool = False,
                 fixed_cond_frames: Optional[List[int]] = None, input_key: str = "img_seq", **reference_only):
        super().__init__()
        # keywords of the reference constructor (vwm/models/diffusion.py:20-46) that only matter for training / logging
        # are accepted and ignored; anything else is a mistyped YAML key and must not be dropped silently
        known = {"optimizer_config", "scheduler_config", "loss_fn_config", "ckpt_path", "use_ema", "ema_decay_rate",
                 "log_keys", "no_cond_log", "compile_model", "slow_spatial_layers", "train_peft_adapters"}
        unknown = sorted(set(reference_only) - known)
        if unknown:
            raise TypeError(f"vista_b200.engine.DiffusionEngine: unexpected keyword(s) {unknown}")
        if reference_only.get("use_ema") or reference_only.get("ckpt_path"):
            raise NotImplementedError("use_ema / ckpt_path are training-side options; load weights with load_state_dict")
        model = instantiate_from_config(network_config)
        wrapper = get_obj_from_str(network_wrapper) if network_wrapper else B200Wrapper
        self.model = wrapper(model, compile_model=False)
        self.denoiser = instantiate_from_config(denoiser_config)
        self.sampler = instantiate_from_config(sampler_config) if sampler_config is not None else None
        # The conditioner (CLIP / VAE-encoder / sinusoids, encoders/modules.py) is outside the hot path: a user-supplied
        # one is hosted as is (do_sample calls model.conditioner.get_unconditional_conditioning), none is built here.
        self._conditioner = instantiate_from_config(conditioner_config) if conditioner_config is not None else None
        if first_stage_config is not None:
            params = first_stage_config.get("params", first_stage_config)
            self.first_stage_model = FirstStage(params["decoder_config"], params.get("encoder_config")
                                                if str(params.get("encoder_config", {}).get("target", "")).startswith("vista_b200.") else None)
        else:
            self.first_stage_model = None
        self.scale_factor = scale_factor
        self.disable_first_stage_autocast = disable_first_stage_autocast
        self.en_and_decode_n_samples_a_time = en_and_decode_n_samples_a_time
        self.num_frames = num_frames
        self.replace_cond_frames = replace_cond_frames
        self.fixed_cond_frames = fixed_cond_frames
        self.input_key = input_key
        self.use_ema = False

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def conditioner(self):
        if self._conditioner is None:
            raise NotImplementedError("no conditioner_config given: the conditioner (CLIP / VAE-encoder / sinusoids) is "
                                      "outside the hot path; pass c / uc dicts, or give a conditioner_config to host")
        return self._conditioner

    @contextlib.contextmanager
    def ema_scope(self, context=None):          # no-op at inference (diffusion.py:241-255, use_ema False)
        yield None

    @torch.no_grad()
    def decode_first_stage(self, z: torch.Tensor, overlap: int = 3) -> torch.Tensor:
        dec = self.first_stage_model.decoder
        if not isinstance(dec, VideoDecoder):
            raise NotImplementedError("decode_first_stage needs vista_b200.vae.VideoDecoder as decoder_config.target")
        if getattr(self.model, "frame_sharded", False):     # one clip on several ranks: spread the decode as well
            import os
            import torch.distributed as dist
            from .vae import _decode_chunks, decode_first_stage_parallel
            group = getattr(self.model, "world_group", None)
            n_chunks = len(_decode_chunks(z.shape[0], self.en_and_decode_n_samples_a_time or z.shape[0], overlap))
            mode = os.environ.get("VISTA_B200_SHARDED_DECODE", "auto")
            if mode == "grouped":
                # opt-in: chunks over sub-groups of <= 4 ranks that frame-shard them (8 ranks: 2 groups x 4 on the 2 chunks)
                from .sharded import ShardedDecoderRuntime, decode_first_stage_grouped
                cache = dec.__dict__.setdefault("_grouped_cache", {})
                return decode_first_stage_grouped(dec.b200_config, lambda g: ShardedDecoderRuntime(dec.b200_config, dec.state_dict(), z.device, group=g),
                                                  cache, z, self.scale_factor, self.en_and_decode_n_samples_a_time, overlap, world_group=group)
            # auto: frame-shard the chunks when there are more ranks than chunks — up to 4 ranks, where that path is validated
            # on hardware; an 8-rank frame chain over NCCL point-to-point timed out in its first hardware run
            # (profiles/r02_sharded_tests_n8_frames_timeout.log), so larger worlds deal whole chunks out instead
            if mode == "1" or (mode == "auto" and n_chunks < dist.get_world_size(group) <= 4):
                # more ranks than chunks: shard the FRAMES of every chunk (sharded.ShardedDecoderRuntime)
                from .sharded import ShardedDecoderRuntime, decode_first_stage_sharded
                srt = getattr(dec, "_sharded_rt", None)
                if srt is None or srt.dev != torch.device(z.device) or srt.group is not group:
                    srt = dec._sharded_rt = ShardedDecoderRuntime(dec.b200_config, dec.state_dict(), z.device, group=group)
                return decode_first_stage_sharded(srt, z, self.scale_factor, self.en_and_decode_n_samples_a_time, overlap)
            # as many chunks as ranks (or fewer ranks): deal whole chunks out, bit-identical to the serial decode
            return decode_first_stage_parallel(dec.runtime(z.device), z, self.scale_factor,
                                               self.en_and_decode_n_samples_a_time, overlap, group=group)
        return _decode_first_stage(dec.runtime(z.device), z, self.scale_factor, self.en_and_decode_n_samples_a_time, overlap)

    @torch.no_grad()
    def decode_first_stage_u8(self, z: torch.Tensor, overlap: int = 3) -> torch.Tensor:
        """Extension: the decoded frames as (F, H, W, 3) uint8 — bit for bit what the reference's
        output path (sample_utils.py:374 clamp, :96-126 scaling / truncation / "t h w c") makes of decode_first_stage's
        result, produced by the decoder's last kernel instead of three full-resolution fp32 passes on the host."""
        dec = self.first_stage_model.decoder
        if not isinstance(dec, VideoDecoder):
            raise NotImplementedError("decode_first_stage_u8 needs vista_b200.vae.VideoDecoder as decoder_config.target")
        return _decode_first_stage(dec.runtime(z.device), z, self.scale_factor, self.en_and_decode_n_samples_a_time, overlap,
                                   u8=True)

    @torch.no_grad()
    def encode_first_stage(self, x, noise: Optional[torch.Tensor] = None, sample: bool = True):
        """diffusion.py:183-195.  Needs ``encoder_config`` (the B200 encoder).  The reference samples the
        posterior with device RNG (DiagonalGaussianRegularizer, sample=True): pass ``noise`` for a reproducible draw,
        ``sample=False`` for the mode."""
        enc = getattr(self.first_stage_model, "encoder", None)
        if enc is None:
            raise NotImplementedError("the VAE encoder is outside the B200 hot path (SURVEY.md §8f); add encoder_config")
        from .vae import Encoder, encode_first_stage as _encode_first_stage
        if not isinstance(enc, Encoder):
            raise NotImplementedError("encode_first_stage needs vista_b200.vae.Encoder as encoder_config.target")
        rt = enc.runtime(x.device)
        if sample and noise is None:
            d = 2 ** (len(rt.cfg.ch_mult) - 1)
            noise = torch.randn(x.shape[0], rt.cfg.z_channels, x.shape[2] // d, x.shape[3] // d, device=x.device)
        return _encode_first_stage(rt, x, self.scale_factor, self.en_and_decode_n_samples_a_time, noise if sample else None)

    @torch.no_grad()
    def sample(self, cond: Dict, cond_frame=None, uc: Union[Dict, None] = None, N: int = 25,
               shape: Union[None, Tuple, List] = None, noise: Optional[torch.Tensor] = None, **kwargs):
        """diffusion.py:306-329; ``noise`` may be injected (device RNG is not reproducible across devices)."""
        randn = torch.randn(N, *shape).to(self.device) if noise is None else noise.to(self.device).clone()
        cond_mask = torch.zeros(N).to(self.device)
        if self.replace_cond_frames:
            assert self.fixed_cond_frames
            cond_mask = cond_mask.reshape(-1, self.num_frames)
            cond_mask[:, self.fixed_cond_frames] = 1
            cond_mask = cond_mask.reshape(-1)
        denoiser = B200Denoiser(self.denoiser, self.model)
        return self.sampler(denoiser, randn, cond, uc=uc, cond_frame=cond_frame, cond_mask=cond_mask)

    # ---- the callers' loops as engine operations (vista_b200/rollout.py) ----
    def rollout(self, cond: Dict, uc: Dict, z: torch.Tensor, num_rounds: int, **kwargs):
        """Long-horizon rollout, the body of sample_utils.do_sample (sample_utils.py:318-373) -> (frames, samples_z)."""
        from .rollout import rollout
        return rollout(self, cond, uc, z, num_rounds, **kwargs)

    def sample_ensemble(self, cond: Dict, uc: Dict, z: torch.Tensor, ensemble_size: int = 5, **kwargs):
        """The reward path (reward_utils.py:318-337) -> (reward, members)."""
        from .rollout import sample_ensemble
        return sample_ensemble(self, cond, uc, z, ensemble_size, **kwargs)
