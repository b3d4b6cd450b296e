"""CPU emulation of the NVLink peer windows (vista_b200/peer.py, csrc/peer.cu) for testing the HOST logic of the peer-memory
sharded step — window layout, remote addresses, flag / sequence protocol, interior shards with two halo neighbours —
without GPUs: every rank's window is a file in /dev/shm that all ranks map (the role CUDA IPC plays on the device), the
three kernels are restated over raw addresses with the same slot / flag arithmetic as the CUDA code.  Test infrastructure
only (memory-ordering questions of the real kernels are out of its reach; tests/test_sharded_gpu.py covers those)."""
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

AR_MAX, PEER_MAX_WORLD = 2048, 16          # kArMax, kPeerMaxWorld of csrc/peer.cu
_NP = {torch.float16: np.float16, torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.uint8: np.uint8}


class FakePeerWindow:
    def __init__(self, group, nbytes: int, tag: str):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = int(nbytes)
        self._paths = [f"/dev/shm/vbwin_{tag}_{r}.bin" for r in range(self.world)]
        mm = np.memmap(self._paths[self.rank], dtype=np.uint8, mode="w+", shape=(self.nbytes,))
        mm[:] = 0
        mm.flush()
        dist.barrier(group=group)
        self._maps = [mm if r == self.rank else np.memmap(self._paths[r], dtype=np.uint8, mode="r+", shape=(self.nbytes,))
                      for r in range(self.world)]
        self.bases = [int(m.ctypes.data) for m in self._maps]
        self.base = self.bases[self.rank]
        self.windows_dev = torch.tensor(self.bases, dtype=torch.int64)
        self._off, self._named = 0, {}
        self._counters, self._done = {}, set()
        dist.barrier(group=group)

    def region(self, name, nbytes, align=1024):
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

    def local(self, off):
        return self.base + off

    def remote(self, r, off):
        return self.bases[r] + off

    def tensor(self, off, shape, dtype):
        n = int(np.prod(shape))
        arr = np.ndarray(shape, dtype=_NP[dtype], buffer=self._maps[self.rank], offset=off)
        t = torch.from_numpy(arr)
        assert t.data_ptr() == self.base + off and t.numel() == n
        return t

    def ptr_array(self, ptrs):
        return torch.tensor(ptrs, dtype=torch.int64)

    def counter(self, name):
        t = self._counters.get(name)
        if t is None:
            t = self._counters[name] = torch.zeros(1, dtype=torch.int32)
        return t

    def once(self, name):
        if name in self._done:
            return False
        self._done.add(name)
        return True

    def close(self):
        self._maps = []
        try:
            os.unlink(self._paths[self.rank])
        except OSError:
            pass


def _u32(addr):
    return C.c_uint32.from_address(int(addr))


def _spin(addr, seq, what, timeout=120.0):
    t0 = time.time()
    while (_u32(addr).value - seq) & 0xFFFFFFFF >= 0x80000000:          # have < seq (wrap-safe, like seq_reached)
        time.sleep(0.0002)
        assert time.time() - t0 < timeout, f"fake peer wait timeout: {what} seq={seq} have={_u32(addr).value}"


def peer_put(src_ptr, src_pitch, rows, row_bytes, dsts_dev, dst_pitch, flags_dev, n_dst, counter, ticket, detail=""):
    seq = int(counter) + 1
    for d in range(n_dst):
        for r in range(rows):
            C.memmove(int(dsts_dev[d]) + r * dst_pitch, int(src_ptr) + r * src_pitch, row_bytes)
    for d in range(n_dst):
        _u32(int(flags_dev[d])).value = seq
    counter += 1


def peer_wait(flags_dev, n, counter, detail=""):
    seq = int(counter) + 1
    for i in range(n):
        _spin(int(flags_dev[i]), seq, f"wait {detail}[{i}]")
    counter += 1


def peer_allreduce_f64(data, windows_dev, slot_off, flag_off, rank, world, counter):
    seq = int(counter) + 1
    par, n = seq & 1, data.numel()
    src = data.contiguous().numpy()
    for r in range(world):
        C.memmove(int(windows_dev[r]) + slot_off + ((par * world + rank) * AR_MAX) * 8, src.ctypes.data, n * 8)
    for r in range(world):
        _u32(int(windows_dev[r]) + flag_off + (par * PEER_MAX_WORLD + rank) * 4).value = seq
    mine = int(windows_dev[rank])
    for r in range(world):
        _spin(mine + flag_off + (par * PEER_MAX_WORLD + r) * 4, seq, f"allreduce flag of rank {r}")
    total = np.zeros(n, dtype=np.float64)
    for r in range(world):                  # rank order, like the kernel
        total += np.ctypeslib.as_array((C.c_double * n).from_address(mine + slot_off + ((par * world + r) * AR_MAX) * 8))
    data.copy_(torch.from_numpy(total).reshape(data.shape))
    counter += 1
    return data
