"""Frame-sharded execution of ONE clip over several GPUs (BASELINE config 5).

Rank r owns a contiguous range of the T frames of the clip — for both CFG halves — and runs every
spatial operator (2-D convolutions, per-frame GroupNorm, spatial attention, feed-forwards) locally.
The three temporal couplings of the UNet cross ranks:

  * temporal self-attention (video_attention.py:127): each query frame needs K, V of all T frames of its
    pixel -> ONE all-gather of the K|V projection per SpatialVideoTransformer (the collective named in
    BASELINE.json), consumed in place by ``b200v_attention_temporal_sharded`` through a frame table;
  * the temporal ResBlock's GroupNorm over (C/32, T, H, W) (video_model.py:67-72): fixed-order local sums,
    an all-reduce of [clips, 32, 2] fp64 values, finalisation with the global element count;
  * the (3,1,1) convolution (video_model.py:38-52): the local zero-padded conv plus two one-frame halo
    corrections  out[first frame] += W_tap0 x prev-rank's last frame,  out[last frame] += W_tap2 x next-rank's
    first frame  (linearity of the convolution), the halos travelling by point-to-point send/recv.

Weights are replicated.  The sampler state is frame-local too; the latent is all-gathered once at the end.
torch.distributed (NCCL) is the transport; the math stays in the C-ABI kernels.

With an even number of ranks the classifier-free-guidance batch is split first (modules.enable_frame_sharding):
ranks [0, W/2) run the unconditional clip, ranks [W/2, W) the conditional one, each half sharding the frames
over W/2 ranks with the collectives above inside its own sub-group; the only traffic between the halves is the
4-channel network output of a rank's frames, exchanged pairwise once per step for the guidance (fused.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import lib as _lib
from . import ops
from .parallel import frame_shards, halo_neighbours
from .spec import UNetConfig
from .unet import Lin, UNetRuntime


class HaloExchange:
    """One-frame halo exchange of a (3,1,1) convolution between frame-shard neighbours: the first local frame goes to
    the previous rank, the last one to the next rank, their boundary frames come back.  `start` / `wait` are plain
    host callables so that they can sit on the launch tape (vista_b200.lib) between the kernels and be replayed;
    `first` / `last` = (staging view, source view, send buffer, receive buffer).  Peers are GLOBAL ranks."""

    def __init__(self, group, prev: Optional[int], nxt: Optional[int], first, last):
        self.copies, self.p2p, self.pending = [], [], []
        if prev is not None:
            stage, src, send, recv = first
            self.copies.append((stage, src))
            self.p2p.append(dist.P2POp(dist.isend, send, prev, group))
            self.p2p.append(dist.P2POp(dist.irecv, recv, prev, group))
        if nxt is not None:
            stage, src, send, recv = last
            self.copies.append((stage, src))
            self.p2p.append(dist.P2POp(dist.isend, send, nxt, group))
            self.p2p.append(dist.P2POp(dist.irecv, recv, nxt, group))

    def start(self):
        for dst, src in self.copies:
            dst.copy_(src)
        self.pending = list(dist.batch_isend_irecv(self.p2p)) if self.p2p else []

    def wait(self):
        for r in self.pending:
            r.wait()
        self.pending = []


class ShardedUNetRuntime(UNetRuntime):
    has_collectives = True        # its step issues NCCL calls: never CUDA-graph-captured, replayed from the launch tape

    def __init__(self, cfg: UNetConfig, sd: Dict[str, torch.Tensor], device, num_frames: int = 25, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.T_full = num_frames
        self.shards = frame_shards(num_frames, self.world)
        self.t0, self.t1 = self.shards[self.rank]
        self.T_pad = max(b - a for a, b in self.shards)
        super().__init__(cfg, sd, device, num_frames=self.t1 - self.t0)     # self.T = local frames per clip
        prev, nxt = halo_neighbours(self.rank, self.world)
        to_global = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
        self.prev = None if prev is None else to_global(prev)       # P2POp peers are global ranks
        self.next = None if nxt is None else to_global(nxt)
        self._tap_w: Dict[int, tuple] = {}
        self._frame_tables: Dict[tuple, torch.Tensor] = {}
        self.comm_bytes = 0

    # ------------------------------------------------------------------ conditioning
    def set_conditioning(self, context: torch.Tensor, y: torch.Tensor):
        """context / y hold the rows of the WHOLE clip(s): (nb*T_full, ...).  The frame-index embedding and the
        temporal cross-attention constant follow the global frame numbering (video_attention.py:256,270-279)."""
        Tf, T = self.T_full, self.T
        nb = context.shape[0] // Tf
        ctx = context.reshape(nb, Tf, -1)
        yy = y.reshape(nb, Tf, -1)
        local_ctx = ctx[:, self.t0:self.t1].reshape(nb * T, 1, -1)
        local_y = yy[:, self.t0:self.t1].reshape(nb * T, -1)
        # time context = context of the first frame of every clip, whoever owns it
        first_ctx = ctx[:, :1].reshape(nb, 1, -1)
        super().set_conditioning(local_ctx, local_y)
        cond = self.cond
        t16 = self.buf("cond.tctx", nb, first_ctx.numel() // nb)
        t16.copy_(first_ctx.reshape(nb, -1))
        frames = torch.arange(Tf, dtype=torch.float32, device=self.dev)
        for t in self.plan.transformers():
            L = self.layers[t.prefix]
            cond["tm"][t.prefix] = self._attn2_const(L["tattn2"], t16, f"cond.tm.{t.prefix}")
            temb = self.buf("cond.temb_full", Tf, t.ch)
            ops.timestep_embedding(frames, temb, t.ch)
            full = self._mlp(temb, *L["pos"], f"cond.posfull.{t.prefix}")
            cond["pos"][t.prefix] = full[self.t0:self.t1]          # rows of the local frames, contiguous view

    # ------------------------------------------------------------------ temporal GroupNorm
    def _gn_temporal(self, x, y, B, hw, norm, eps, silu, idx, fps):
        nb = B // fps
        Cc = norm[0].numel()
        sums = self.buf(f"gn.sums{nb}", nb * self.cfg.num_groups, 2, torch.float64)
        ops.groupnorm_sums(x, B, hw, Cc, sums, fps, groups=self.cfg.num_groups, ws=self.gn_ws)
        _lib.tape_host(lambda: dist.all_reduce(sums, group=self.group))
        self.comm_bytes += sums.numel() * 8
        count = float(Cc // self.cfg.num_groups) * hw * self.T_full
        return ops.groupnorm_finalize_apply(x, y, B, hw, norm[0], norm[1], eps, silu, sums, count,
                                            self.gn_stats[idx, :nb], fps, groups=self.cfg.num_groups)

    # ------------------------------------------------------------------ temporal convolution with halos
    def _tap_weights(self, lin: Lin):
        key = lin.w.data_ptr()
        tw = self._tap_w.get(key)
        if tw is None:
            Cc = lin.w.shape[1] // 3
            tw = (Lin(lin.w[:, :Cc].contiguous(), None, lin.tile_n), Lin(lin.w[:, 2 * Cc:].contiguous(), None, lin.tile_n))
            self._tap_w[key] = tw
        return tw

    def _tconv(self, a, lin: Lin, out, hw: int, nb: int, **epi):
        T, Cc = self.T, a.shape[1]
        # boundary frames of the input go to the neighbours while the local convolution runs
        send_first = self.buf("halo.sf", nb * hw, Cc)
        send_last = self.buf("halo.sl", nb * hw, Cc)
        recv_prev = self.buf("halo.rp", nb * hw, Cc)
        recv_next = self.buf("halo.rn", nb * hw, Cc)
        a4 = a.reshape(nb, T, hw, Cc) if a.is_contiguous() else None
        av = a4 if a4 is not None else a.as_strided((nb, T, hw, Cc), (T * hw * a.stride(0), hw * a.stride(0), a.stride(0), 1))
        halo = HaloExchange(self.group, self.prev, self.next,
                            first=(send_first.view(nb, hw, Cc), av[:, 0], send_first, recv_prev),
                            last=(send_last.view(nb, hw, Cc), av[:, T - 1], send_last, recv_next))
        start_halo, wait_halo = halo.start, halo.wait
        _lib.tape_host(start_halo)
        self.comm_bytes += 2 * nb * hw * Cc * 2 * ((self.prev is not None) + (self.next is not None))
        self.gemm(a, lin, out, taps=ops.TAPS_T3, geom=(hw, T, nb), **epi)
        _lib.tape_host(wait_halo)
        s_acc = epi.get("s_acc", 1.0)
        w0, w2 = self._tap_weights(lin)
        ov = out.as_strided((nb, T, hw, out.shape[1]), (T * hw * out.stride(0), hw * out.stride(0), out.stride(0), 1))
        for b in range(nb):
            if self.prev is not None:      # frame t0-1 contributes through tap 0 to our first frame
                o = ov[b, 0]
                self.gemm(recv_prev[b * hw:(b + 1) * hw], w0, o, s_acc=s_acc, res1=o)
            if self.next is not None:      # frame t1 contributes through tap 2 to our last frame
                o = ov[b, T - 1]
                self.gemm(recv_next[b * hw:(b + 1) * hw], w2, o, s_acc=s_acc, res1=o)
        return out

    # ------------------------------------------------------------------ temporal attention with gathered K|V
    def _frame_table(self, nb: int, hw: int) -> torch.Tensor:
        key = (nb, hw)
        tab = self._frame_tables.get(key)
        if tab is None:
            rows = []
            for b in range(nb):
                for t in range(self.T_full):
                    r = next(i for i, (a, e) in enumerate(self.shards) if a <= t < e)
                    tl = t - self.shards[r][0]
                    rows.append(((r * nb + b) * self.T_pad + tl) * hw)
            tab = torch.tensor(rows, dtype=torch.int64, device=self.dev)
            self._frame_tables[key] = tab
        return tab

    def _attn_temporal(self, qkv, o, nb: int, hw: int, heads: int, Cc: int):
        T, Tp, W = self.T, self.T_pad, self.world
        send = self.buf("kv.send", nb * Tp * hw, 2 * Cc)
        recv = self.buf("kv.recv", W * nb * Tp * hw, 2 * Cc)
        # the K|V column block of the fused q|k|v projection as a strided (clip, frame, pixel, 2C) view
        kv = qkv.as_strided((nb, T, hw, 2 * Cc), (T * hw * qkv.stride(0), hw * qkv.stride(0), qkv.stride(0), 1),
                            qkv.storage_offset() + Cc)
        dst = send.view(nb, Tp, hw, 2 * Cc)[:, :T]
        _lib.tape_host(lambda: dst.copy_(kv))
        _lib.tape_host(lambda: dist.all_gather_into_tensor(recv, send, group=self.group))
        self.comm_bytes += recv.numel() * 2
        tab = self._frame_table(nb, hw)
        return ops.attention_temporal_sharded(qkv[:, :Cc], recv[:, :Cc], recv[:, Cc:], o, nb, T, self.T_full, hw, heads, tab)


def gather_latent(x_local: torch.Tensor, num_frames: int, group=None) -> torch.Tensor:
    """All-gather of the frame-sharded latent (T_loc,4,h,w) -> (T,4,h,w) on every rank."""
    world = dist.get_world_size(group)
    shards = frame_shards(num_frames, world)
    pad = max(b - a for a, b in shards)
    buf = x_local.new_zeros((pad,) + tuple(x_local.shape[1:]))
    buf[: x_local.shape[0]] = x_local
    out = x_local.new_empty((world * pad,) + tuple(x_local.shape[1:]))
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = out.reshape(world, pad, *x_local.shape[1:])
    return torch.cat([parts[r, : b - a] for r, (a, b) in enumerate(shards)], dim=0)
