"""TEST INFRASTRUCTURE — imports the *real* reference (OpenDriveLab/Vista) modules in the build
container so that golden fixtures can be generated from them.  Never imported by the product.

The reference cannot be imported as-is offline: ``xformers``, ``pytorch_lightning``,
``omegaconf``, ``open_clip``, ``kornia`` are missing.  This installs the minimal
``sys.modules`` stubs of SURVEY.md Appendix H and an SDPA shim for
``xformers.ops.memory_efficient_attention`` (mathematically softmax(QK^T/sqrt(d))V, the same
function xformers computes with attn_bias=None, p=0).  ``/root/reference`` does not exist on the
GPU box — callers must use :func:`reference_available` and skip.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F
import yaml

REFERENCE_ROOT = os.environ.get("VISTA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vwm", "modules", "attention.py"))


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.__version__ = "2.0.1"

        class LightningModule(nn.Module):
            global_step = 0

            @property
            def device(self):
                return next(self.parameters()).device

        pl.LightningModule = LightningModule
        pl.seed_everything = lambda s, **k: torch.manual_seed(s)
        sys.modules["pytorch_lightning"] = pl
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class ListConfig(list):
            pass

        class OmegaConf:
            load = staticmethod(lambda p: yaml.safe_load(open(p)))

        oc.ListConfig, oc.OmegaConf = ListConfig, OmegaConf
        sys.modules["omegaconf"] = oc
    for name in ("kornia", "open_clip"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "xformers" not in sys.modules:
        xf, xo = types.ModuleType("xformers"), types.ModuleType("xformers.ops")

        def memory_efficient_attention(q, k, v, attn_bias=None, op=None):  # (B*H, N, D)
            return F.scaled_dot_product_attention(q[None], k[None], v[None])[0]

        xo.memory_efficient_attention = memory_efficient_attention
        xo.LowerTriangularMask = type("LowerTriangularMask", (), {})
        xf.ops = xo
        sys.modules["xformers"] = xf
        sys.modules["xformers.ops"] = xo


def load_reference():
    """Returns a namespace with the reference classes on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        from vwm.modules.diffusionmodules.video_model import VideoUNet
        from vwm.modules.diffusionmodules.wrappers import OpenAIWrapper
        from vwm.modules.diffusionmodules.denoiser import Denoiser
        from vwm.modules.diffusionmodules.sampling import EulerEDMSampler
        from vwm.modules.diffusionmodules import guiders, discretizer, denoiser_scaling
        from vwm.modules.autoencoding.temporal_ae import VideoDecoder
    ns = types.SimpleNamespace(VideoUNet=VideoUNet, OpenAIWrapper=OpenAIWrapper, Denoiser=Denoiser,
                               EulerEDMSampler=EulerEDMSampler, guiders=guiders, discretizer=discretizer,
                               denoiser_scaling=denoiser_scaling, VideoDecoder=VideoDecoder)
    return ns


def vista_yaml():
    return yaml.safe_load(open(os.path.join(REFERENCE_ROOT, "configs", "inference", "vista.yaml")))


def build_ref_unet(cfg):
    """Instantiate the reference VideoUNet for a ``vista_b200.spec.UNetConfig`` (other ctor args
    as in configs/inference/vista.yaml:19-40)."""
    ref = load_reference()
    params = dict(vista_yaml()["model"]["params"]["network_config"]["params"])
    params.update(in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                  model_channels=cfg.model_channels,
                  attention_resolutions=list(cfg.attention_resolutions),
                  num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult),
                  num_head_channels=cfg.num_head_channels, context_dim=cfg.context_dim,
                  adm_in_channels=cfg.adm_in_channels, action_control=cfg.action_control)
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.VideoUNet(**params).eval()


def build_ref_decoder(cfg):
    ref = load_reference()
    params = dict(vista_yaml()["model"]["params"]["first_stage_config"]["params"]["decoder_config"]["params"])
    params.update(ch=cfg.ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                  z_channels=cfg.z_channels, out_ch=cfg.out_ch)
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.VideoDecoder(**params).eval()


def build_ref_encoder(cfg):
    """Reference ``Encoder`` (vwm/modules/diffusionmodules/model.py:445) for a ``vista_b200.spec.EncoderConfig``; other
    ctor args as in configs/inference/vista.yaml:155-168."""
    load_reference()
    from vwm.modules.diffusionmodules.model import Encoder
    params = dict(vista_yaml()["model"]["params"]["first_stage_config"]["params"]["encoder_config"]["params"])
    params.update(ch=cfg.ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, z_channels=cfg.z_channels,
                  in_channels=cfg.in_channels, double_z=cfg.double_z)
    with contextlib.redirect_stdout(io.StringIO()):
        return Encoder(**params).eval()


def build_ref_sampler(num_steps, guider="VanillaCFG", scale=2.5, num_frames=25):
    ref = load_reference()
    if guider == "VanillaCFG":
        gcfg = {"target": "vwm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": scale}}
    elif guider == "TrianglePredictionGuider":
        gcfg = {"target": "vwm.modules.diffusionmodules.guiders.TrianglePredictionGuider",
                "params": {"max_scale": scale, "num_frames": num_frames}}
    else:
        raise KeyError(guider)
    return ref.EulerEDMSampler(
        num_steps=num_steps, device="cpu", s_churn=0.0, s_tmin=0.0, s_tmax=999.0, s_noise=1.0, verbose=False,
        discretization_config={"target": "vwm.modules.diffusionmodules.discretizer.EDMDiscretization",
                               "params": {"sigma_min": 0.002, "sigma_max": 700.0, "rho": 7.0}},
        guider_config=gcfg)


def build_ref_denoiser(num_frames=25):
    ref = load_reference()
    return ref.Denoiser({"target": "vwm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise"},
                        num_frames=num_frames)
