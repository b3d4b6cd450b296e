// Diagnostic: how long does one tcgen05.mma (kind::f16, K = 16, operands in shared memory) occupy the tensor pipe
// as a function of its shape (M in {64, 128}, N), and do MMAs into DIFFERENT accumulators overlap?
// One CTA per SM; one thread issues `iters` groups of 4 MMAs (the 4 K-steps of a 64-wide SW128 operand tile),
// cycling over `n_acc` accumulator column ranges, then commits to an mbarrier; the elapsed SM clock over the
// issue + completion of the whole train is reported per CTA.  Operand contents are irrelevant (zeros).
// This measurement decides tile shapes (profiles/r01_umma_n_sweep.md); it is not on the
// product path.
#include <stdint.h>
#include "../../vista_b200/csrc/host.cuh"
#include "../../vista_b200/csrc/ptx.cuh"

namespace vb {

__global__ void __launch_bounds__(128, 1)
mma_probe_kernel(int M, int N, int iters, int n_acc, int a_mn_major, int a_in_tmem, float* __restrict__ cycles_per_mma) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* done = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* sA = smem + 1024;             // 128 rows x 128 B
  uint8_t* sB = sA + 128 * 128;          // 256 rows x 128 B
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_f16(M, N, 0, a_mn_major, 0);
    const uint64_t desc_hi = make_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lo = (uint32_t)(make_desc_sw128(smem_u32(sA), 16, 1024) & 0xFFFFFFFFull);
    const uint32_t b_lo = (uint32_t)(make_desc_sw128(smem_u32(sB), 16, 1024) & 0xFFFFFFFFull);
    const long long t0 = clock64();
    int acc = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + acc * N;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (a_in_tmem)     // A operand: 8 packed columns per K = 16 step, placed behind the accumulators
          umma_f16_ts(d, tmem_base + 480 + 8 * k, desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (it | k) != 0 ? 1u : 0u);
        else
          umma_f16(d, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (it | k) != 0 ? 1u : 0u);
      }
      if (++acc == n_acc) acc = 0;
    }
    umma_commit(done);
    mbar_wait(done, 0, 90);
    const long long t1 = clock64();
    cycles_per_mma[blockIdx.x] = (float)(t1 - t0) / (4.0f * iters);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vb

extern "C" int b200v_debug_mma_probe(int32_t M, int32_t N, int32_t iters, int32_t n_acc, int32_t a_mn_major,
                                     int32_t a_in_tmem, float* cycles_per_mma, int32_t n_ctas, void* stream) {
  using namespace vb;
  VB_REQUIRE(cycles_per_mma && n_ctas > 0, "mma_probe: null output");
  VB_REQUIRE(M == 64 || M == 128, "mma_probe: M=%d must be 64 or 128", M);
  VB_REQUIRE(N >= 8 && N <= 256 && N % (M == 128 ? 16 : 8) == 0, "mma_probe: N=%d invalid for M=%d", N, M);
  VB_REQUIRE(iters > 0 && n_acc >= 1 && n_acc * N <= (a_in_tmem ? 480 : 512), "mma_probe: iters / n_acc out of range");
  const int smem_bytes = 1024 + 1024 + (128 + 256) * 128;
  static bool attr_set[64] = {false};
  if (vb::first_use_on_device(attr_set)) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  mma_probe_kernel<<<n_ctas, 128, smem_bytes, (cudaStream_t)stream>>>(M, N, iters, n_acc, a_mn_major, a_in_tmem,
                                                                       cycles_per_mma);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
