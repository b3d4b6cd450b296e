"""Test doubles for the reference-caller seam test (tests/test_reference_seam_cpu.py): a conditioner stand-in that the
REAL vwm DiffusionEngine instantiates by dotted path (the real one needs the CLIP ViT-H tower, which is not available
offline).  Deterministic functions of the batch, so that the re-conditioning between rollout rounds is exercised."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _Embedder:
    def __init__(self, input_key, with_skip=False):
        self.input_key = input_key
        if with_skip:
            self.skip_encode = False


class FakeConditioner(nn.Module):
    """Same surface as GeneralConditioner for sample_utils.get_condition (sample_utils.py:255-277): `.embedders[i].input_key`
    (+ `.skip_encode` on the frame embedder) and `get_unconditional_conditioning(batch, batch_uc, force_uc_zero_embeddings)`
    -> (c, uc) with the keys OpenAIWrapper reads: crossattn (b,1,ctx), vector (b,adm), concat (b,4,h,w)."""

    def __init__(self, context_dim=1024, action_dim=2432, adm_in_channels=768, down=8, seed=5):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("ctx", torch.randn(1, 1, context_dim + action_dim, generator=g))
        self.register_buffer("vec", torch.randn(1, adm_in_channels, generator=g))
        self.down = down
        self.embedders = [_Embedder("cond_frames_without_noise"), _Embedder("cond_frames", with_skip=True)]

    def get_unconditional_conditioning(self, batch_c, batch_uc=None, force_uc_zero_embeddings=None):
        cf, cw = batch_c["cond_frames"].float(), batch_c["cond_frames_without_noise"].float()
        n = cf.shape[0]
        # frame embedder: latents pass through when skip_encode is set (encoders/modules.py:470-475), else "encode" = pool
        if getattr(self.embedders[1], "skip_encode", False) or cf.shape[1] == 4:
            concat = cf
        else:
            concat = F.avg_pool2d(cf, self.down)[:, [0, 1, 2, 0]] * 4.0
        shade = cw.mean(dim=(1, 2, 3)).reshape(n, 1, 1)                   # the "CLIP" embedding depends on the clean frame
        cross = (self.ctx + 0.1 * shade).expand(n, -1, -1).clone()
        vec = self.vec.expand(n, -1).clone()
        c = {"crossattn": cross, "vector": vec, "concat": concat.clone()}
        uc = {"crossattn": torch.zeros_like(cross), "vector": vec.clone(), "concat": torch.zeros_like(concat)}
        return c, uc
