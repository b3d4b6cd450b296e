"""CPU emulation of the C-ABI operators, for testing the HOST-side executors (buffer plumbing, weight packing,
tap / gather conventions) without a GPU.  Test infrastructure only: it patches vista_b200.ops inside a context
manager; nothing under vista_b200/ imports it.  Numerics: fp16 operands, fp32 accumulation and statistics, fp16 (or
fp32) outputs — the same rounding points as the kernels, so the executors can be compared with the golden fixtures
at the GPU tolerance."""
import contextlib

import torch
import torch.nn.functional as F


def _f(t):
    return t.float()


def gemm(a, w, out, *, taps=((0, 0),), geom=None, bias=None, rowvec=None, rv_div=1, rv_mod=1, res1=None, s_res1=1.0,
         res2=None, s_res2=1.0, s_acc=1.0, act=0, tile_n=None, cin=None):
    tokens = a.shape[0]
    N, K = w.shape
    ntaps = len(taps)
    cin = cin if cin is not None else K // ntaps
    x = _f(a[:, :cin])
    if geom is None:
        cols = x
    else:
        W, H, NB = geom
        assert W * H * NB == tokens
        img = x.reshape(NB, H, W, cin)
        parts = []
        for dh, dw in taps:                      # out[b,h,w] reads in[b, h+dh, w+dw], zero outside
            sh = torch.zeros_like(img)
            hs, he = max(0, -dh), min(H, H - dh)
            ws, we = max(0, -dw), min(W, W - dw)
            if hs < he and ws < we:
                sh[:, hs:he, ws:we] = img[:, hs + dh:he + dh, ws + dw:we + dw]
            parts.append(sh.reshape(tokens, cin))
        cols = torch.cat(parts, dim=1)
    acc = cols @ _f(w).t()
    if bias is not None:
        acc = acc + _f(bias)[:N]
    acc = acc * s_acc
    if rowvec is not None:
        rows = (torch.arange(tokens) // rv_div) % rv_mod
        acc = acc + _f(rowvec)[rows][:, :N]
    if act == 1:
        acc = F.silu(acc)
    elif act == 2:
        raise NotImplementedError("GEGLU is not needed by the VAE executors")
    if res1 is not None:
        acc = acc + s_res1 * _f(res1)
    if res2 is not None:
        acc = acc + s_res2 * _f(res2)
    n_out = out.shape[1]
    out.copy_(acc[:, :n_out].to(out.dtype))
    return out


class GNWorkspace:
    def __init__(self, device, max_stats=4096):
        self.device = device

    def reserve(self, n):
        pass


def groupnorm_scratch(frames, tokens_per_frame, groups=32):
    return 0


def groupnorm(x, y, frames, tokens_per_frame, gamma, beta, eps, silu, stats=None, frames_per_stat=1, groups=32, ws=None):
    C = gamma.numel()
    xs = _f(x[:, :C]).reshape(frames // frames_per_stat, frames_per_stat * tokens_per_frame, groups, C // groups)
    mean = xs.mean(dim=(1, 3), keepdim=True)
    var = xs.var(dim=(1, 3), unbiased=False, keepdim=True)
    o = ((xs - mean) * torch.rsqrt(var + eps)).reshape(-1, C) * _f(gamma) + _f(beta)
    if silu:
        o = F.silu(o)
    y.copy_(o.to(y.dtype))
    return y


def conv3x3_small_cin(x8, cin, w, bias, out, NB, H, W):
    img = _f(x8[:, :cin]).reshape(NB, H, W, cin).permute(0, 3, 1, 2)
    o = F.conv2d(img, _f(w), None if bias is None else _f(bias), padding=1)
    out.copy_(o.permute(0, 2, 3, 1).reshape(NB * H * W, -1).to(out.dtype))
    return out


def im2col_s2_asym(x, out, NB, H, W, Cc):
    Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    img = F.pad(x[:, :Cc].reshape(NB, H, W, Cc), (0, 0, 0, 2, 0, 2))          # zero beyond the right / bottom edge
    parts = [img[:, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2] for kh in range(3) for kw in range(3)]
    out.copy_(torch.cat(parts, dim=-1).reshape(NB * Ho * Wo, 9 * Cc))
    return out


def upsample2x(x, out, NB, H, W, Cc):
    img = x[:, :Cc].reshape(NB, H, W, Cc)
    out.copy_(img.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).reshape(NB * 4 * H * W, Cc))
    return out


def softmax_rows(x, y):
    y.copy_(torch.softmax(_f(x), dim=-1).to(y.dtype))
    return y


def nchw_to_tokens(x, out, NB, Cc, H, W):
    out[:, :Cc] = x.permute(0, 2, 3, 1).reshape(NB * H * W, Cc).to(out.dtype)
    return out


def tokens_to_nchw(x, out, NB, Cc, H, W):
    out.copy_(_f(x[:, :Cc]).reshape(NB, H, W, Cc).permute(0, 3, 1, 2))
    return out


def time_mix_small(x, w, bias, out, blend, T, HW, Cc, out_frame0=0, skip_frames=0):
    v = _f(x[:, :Cc]).reshape(T, HW, Cc)
    acc = torch.zeros(T, HW, Cc) + (0 if bias is None else _f(bias))
    for kt in range(3):
        for t in range(T):
            tt = t + kt - 1
            if 0 <= tt < T:
                acc[t] += v[tt] @ _f(w)[:, :, kt].t()
    hh = out.shape[2]
    img = acc.reshape(T, hh, HW // hh, Cc).permute(0, 3, 1, 2)
    for t in range(skip_frames, T):
        if blend is not None and int(blend[t]) != 0:
            out[out_frame0 + t] = 0.5 * (out[out_frame0 + t] + img[t])
        else:
            out[out_frame0 + t] = img[t]
    return out


_PATCHED = ["gemm", "GNWorkspace", "groupnorm_scratch", "groupnorm", "conv3x3_small_cin", "im2col_s2_asym", "upsample2x",
            "softmax_rows", "nchw_to_tokens", "tokens_to_nchw", "time_mix_small"]


@contextlib.contextmanager
def patched_ops():
    """Swap the emulations into vista_b200.ops for the duration of the block."""
    from vista_b200 import ops
    saved = {k: getattr(ops, k) for k in _PATCHED}
    try:
        for k in _PATCHED:
            setattr(ops, k, globals()[k])
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
