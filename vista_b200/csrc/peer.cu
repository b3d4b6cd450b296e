// Collectives of the frame-sharded step over NVLink peer memory, without NCCL on the data path.
// Every rank owns one "window" (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by its peers): the peers STORE into
// it (halo frames, K|V slabs, GroupNorm partial sums), then raise a flag in it with a system-scope release; the owner's
// next kernel spins on its local flags with acquire loads.  All of it is plain stream-ordered kernels — no host
// synchronisation, no communicator call — so a sharded sampler step is a fixed launch sequence again and is replayed from a
// CUDA graph like the single-GPU one.  Sequence numbers live in device memory (a counter per channel and side, bumped by the
// kernels themselves), which is what makes a captured graph valid for every replay.
//   peer_allreduce_f64 : [n] doubles, every rank writes its vector into slot[rank] of every window, waits for the W flags of
//                        its own window, sums the slots in RANK ORDER (bit-identical result on every rank) — the temporal
//                        GroupNorm statistic (video_model.py:67-72) — one launch instead of an NCCL all-reduce;
//   peer_put           : strided rows -> the same rows in up to 8 remote (or local) destinations + one flag per destination,
//                        raised by the last block (ticket): the (3,1,1) convolution's halo frames (openaimodel.py:190-193 across
//                        shards) and the K|V slabs of the temporal attention's all-gather (video_attention.py:127);
//   peer_wait          : one thread spins until n local flags reach the next sequence number.
// Slots are double-buffered by sequence parity where a fast peer could otherwise overwrite data its neighbour still reads;
// waits are bounded (trap) like every other wait of this library.
#include "../../include/vista_b200.h"
#include "host.cuh"

namespace vb {

#ifndef VB_PEER_TIMEOUT_CYCLES
#define VB_PEER_TIMEOUT_CYCLES (120000000000ll)   // ~60 s: a peer may still be capturing its step graph (cudaFree storms with peer access on are slow); the NCCL watchdog of the caller is the outer bound
#endif

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// sequence numbers only grow; the comparison survives a wrap of the 32-bit counter
__device__ __forceinline__ bool seq_reached(uint32_t have, uint32_t want) { return (int32_t)(have - want) >= 0; }

__device__ __forceinline__ void spin_until(const uint32_t* flag, uint32_t seq, int tag) {
  long long t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    if (seq_reached(ld_acquire_sys(flag), seq)) return;
    __nanosleep(64);
    if ((spins & 1023u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > VB_PEER_TIMEOUT_CYCLES) {
        printf("vista_b200: peer flag wait timeout tag=%d seq=%u have=%u\n", tag, seq, ld_acquire_sys(flag));
        __trap();
      }
    }
  }
}

constexpr int kArMax = 2048;     // doubles per all-reduce
constexpr int kPeerMaxWorld = 16;

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* __restrict__ data, int n, void* const* __restrict__ windows, long long slot_off,
                          long long flag_off, int rank, int world, uint32_t* __restrict__ counter) {
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) s_seq = *counter + 1;
  __syncthreads();
  const uint32_t seq = s_seq;
  const int par = seq & 1;
  for (int r = 0; r < world; ++r) {
    double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(windows[r]) + slot_off) + ((long long)(par * world + rank)) * kArMax;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = data[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world)
    st_release_sys(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(windows[threadIdx.x]) + flag_off) + par * kPeerMaxWorld + rank, seq);
  const char* mine = reinterpret_cast<const char*>(windows[rank]);
  if (threadIdx.x < world)
    spin_until(reinterpret_cast<const uint32_t*>(mine + flag_off) + par * kPeerMaxWorld + threadIdx.x, seq, 100 + threadIdx.x);
  __syncthreads();
  const double* slots = reinterpret_cast<const double*>(mine + slot_off) + (long long)par * world * kArMax;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_volatile_f64(slots + (long long)r * kArMax + i);   // rank order: same bits everywhere
    data[i] = s;
  }
  if (threadIdx.x == 0) *counter = seq;
}

// rows of row_bytes (multiple of 16) at src + r * src_pitch -> dst[d] + r * dst_pitch for d < n_dst
__global__ void __launch_bounds__(256)
peer_put_kernel(const char* __restrict__ src, long long src_pitch, long long rows, int row_vecs, char* const* __restrict__ dsts,
                long long dst_pitch, uint32_t* const* __restrict__ flags, int n_dst, uint32_t* __restrict__ counter,
                uint32_t* __restrict__ ticket) {
  const long long total = rows * row_vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row_vecs;
    const int v = (int)(i - r * row_vecs);
    const uint4 val = *reinterpret_cast<const uint4*>(src + r * src_pitch + (long long)v * 16);
    for (int d = 0; d < n_dst; ++d) *reinterpret_cast<uint4*>(dsts[d] + r * dst_pitch + (long long)v * 16) = val;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t seq = *counter + 1;
    for (int d = 0; d < n_dst; ++d) st_release_sys(flags[d], seq);
    *counter = seq;
    *ticket = 0;
  }
}

__global__ void peer_wait_kernel(const uint32_t* const* __restrict__ flags, int n, uint32_t* __restrict__ counter) {
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) s_seq = *counter + 1;
  __syncthreads();
  if ((int)threadIdx.x < n) spin_until(flags[threadIdx.x], s_seq, 200 + threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) *counter = s_seq;
}

}  // namespace vb

extern "C" int b200v_peer_alloc(int64_t bytes, void** ptr, void* handle64) {
  VB_REQUIRE(bytes > 0 && ptr && handle64, "peer_alloc: bad args");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  VB_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  VB_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  VB_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  VB_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

extern "C" int b200v_peer_open(const void* handle64, void** ptr) {
  VB_REQUIRE(handle64 && ptr, "peer_open: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  VB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return 0;
}

extern "C" int b200v_peer_close(void* ptr) {
  if (ptr) VB_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int b200v_peer_free(void* ptr) {
  if (ptr) VB_CHECK_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int b200v_peer_allreduce_max(void) { return vb::kArMax; }

extern "C" int b200v_peer_allreduce_f64(double* data, int32_t n, void* const* windows_dev, int64_t slot_off, int64_t flag_off,
                                        int32_t rank, int32_t world, uint32_t* counter, void* stream) {
  VB_REQUIRE(data && windows_dev && counter && n > 0 && n <= vb::kArMax && world >= 1 && world <= vb::kPeerMaxWorld && rank >= 0 &&
                 rank < world && slot_off % 8 == 0 && flag_off % 4 == 0,
             "peer_allreduce_f64: bad args (n <= %d, world <= %d)", vb::kArMax, vb::kPeerMaxWorld);
  vb::peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(data, n, windows_dev, slot_off, flag_off, rank, world, counter);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200v_peer_put(const void* src, int64_t src_pitch, int64_t rows, int64_t row_bytes, void* const* dsts_dev,
                              int64_t dst_pitch, uint32_t* const* flags_dev, int32_t n_dst, uint32_t* counter, uint32_t* ticket,
                              void* stream) {
  VB_REQUIRE(src && dsts_dev && flags_dev && counter && ticket && rows > 0 && row_bytes > 0 && row_bytes % 16 == 0 &&
                 src_pitch % 16 == 0 && dst_pitch % 16 == 0 && n_dst >= 1 && n_dst <= 8 &&
                 (reinterpret_cast<uintptr_t>(src) & 15) == 0,
             "peer_put: bad args (16-byte rows / pitches, 1..8 destinations)");
  const long long total = rows * (row_bytes / 16);
  long long blocks = (total + 255) / 256;
  const long long cap = 2ll * vb::device_sm_count();
  if (blocks > cap) blocks = cap;
  vb::peer_put_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const char*>(src), src_pitch, rows, (int)(row_bytes / 16), reinterpret_cast<char* const*>(dsts_dev), dst_pitch,
      flags_dev, n_dst, counter, ticket);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200v_peer_wait(const uint32_t* const* flags_dev, int32_t n, uint32_t* counter, void* stream) {
  VB_REQUIRE(flags_dev && counter && n >= 1 && n <= 32, "peer_wait: bad args");
  vb::peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_dev, n, counter);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
