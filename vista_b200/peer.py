"""Peer windows for the NVLink collectives of the frame-sharded step (csrc/peer.cu, include/vista_b200.h).

Every rank of a process group allocates one window through the C-ABI (cudaMalloc + CUDA IPC export), the handles travel once
over ``torch.distributed`` (plumbing), every rank maps its peers' windows.  Regions are carved with a bump allocator that all
ranks run with the same sizes, so an offset names the same region in every window and ``remote(r, off)`` is the address of
rank r's copy as mapped here.  torch tensors over window memory are zero-copy views (``__cuda_array_interface__``)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

from . import lib as _lib

_TYPESTR = {torch.float16: "<f2", torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1",
            torch.int64: "<i8"}


class _Raw:
    def __init__(self, ptr: int, shape: Tuple[int, ...], dtype: torch.dtype):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": _TYPESTR[dtype], "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerWindow:
    def __init__(self, group, nbytes: int, device):
        self.group, self.dev = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = int(nbytes)
        l = _lib.load()
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(self.dev):
            _lib.check(l.b200v_peer_alloc(self.nbytes, C.byref(ptr), handle), "b200v_peer_alloc")
        self.base = int(ptr.value)
        handles: List[bytes] = [b""] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.bases: List[int] = []
        self._opened: List[int] = []
        for r in range(self.world):
            if r == self.rank:
                self.bases.append(self.base)
                continue
            p = C.c_void_p()
            with torch.cuda.device(self.dev):
                _lib.check(l.b200v_peer_open(C.create_string_buffer(handles[r], 64), C.byref(p)), "b200v_peer_open")
            self.bases.append(int(p.value))
            self._opened.append(int(p.value))
        self.windows_dev = torch.tensor(self.bases, dtype=torch.int64, device=self.dev)
        self._off = 0
        self._named: Dict[str, Tuple[int, int]] = {}
        self._counters: Dict[str, torch.Tensor] = {}
        self._done = set()
        dist.barrier(group=group)          # every window exists and is zeroed before anyone stores into a peer

    # ------------------------------------------------------------------ layout
    def region(self, name: str, nbytes: int, align: int = 1024) -> int:
        """Offset of the named region (allocated on first use; every rank must ask for the same names / sizes in the same
        order — the executors do, they run the same plan)."""
        hit = self._named.get(name)
        if hit is not None:
            assert hit[1] >= nbytes, (name, hit, nbytes)
            return hit[0]
        off = (self._off + align - 1) // align * align
        if off + nbytes > self.nbytes:
            raise RuntimeError(f"peer window of {self.nbytes} bytes exhausted by region {name!r} ({nbytes} bytes at {off})")
        self._off = off + nbytes
        self._named[name] = (off, nbytes)
        return off

    def local(self, off: int) -> int:
        return self.base + off

    def remote(self, r: int, off: int) -> int:
        return self.bases[r] + off

    def tensor(self, off: int, shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        t = torch.as_tensor(_Raw(self.base + off, shape, dtype), device=self.dev)
        assert t.data_ptr() == self.base + off
        return t

    def ptr_array(self, ptrs: List[int]) -> torch.Tensor:
        return torch.tensor(ptrs, dtype=torch.int64, device=self.dev)

    def counter(self, name: str) -> torch.Tensor:
        """Device-resident sequence counter (or ticket) of a channel, kept with the WINDOW: the flags it is compared with live
        in the window too, so a runtime that is rebuilt (new weights, new loop state) must go on counting where the old one
        stopped instead of restarting at zero against flags that already hold larger values."""
        t = self._counters.get(name)
        if t is None:
            t = self._counters[name] = torch.zeros(1, dtype=torch.int32, device=self.dev)
        return t

    def once(self, name: str) -> bool:
        """True the first time it is asked for `name` (one-off protocol steps such as priming an acknowledgement flag)."""
        if name in self._done:
            return False
        self._done.add(name)
        return True

    def close(self):
        l = _lib.load()
        for p in self._opened:
            l.b200v_peer_close(p)
        self._opened = []
        if self.base:
            l.b200v_peer_free(self.base)
            self.base = 0
