// Spatial self-attention, head dim 64, on tcgen05 tensor cores (flash-attention style, non-causal).
// One CTA = 128 queries of one (frame, head); KV streamed in blocks of 128 keys by TMA.
//   warp 4  : TMA producer (Q once, then K/V blocks, 2-deep ring)
//   warp 5  : TMEM owner + tcgen05.mma issuer:  S = Q K^T  (M128 N128 K64)  and  O_j = P_j V_j (M128 N64 K128)
//   warps 0-3: softmax; thread i owns query row i (TMEM lane i): no shuffles for row max / sum.
// S is double-buffered in TMEM so S_{j+1} is computed while the softmax of block j runs; P goes back
// through shared memory (K-major, 128B swizzle) as the A operand of the second MMA; V is consumed
// straight from its TMA tile as an MN-major B operand (no transposed copy of V anywhere).
// O_j is produced non-accumulating and folded into fp32 registers with the running rescale.
#include "../../include/vista_b200.h"
#include "host.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace vb {

constexpr int kTile = 128;              // queries per CTA, keys per block
constexpr int kTileBytes = 128 * 128;   // 128 rows x 64 fp16 = 16 KB

struct AttnParams {
  int seq;        // tokens per frame
  int heads;
  int n_kv;       // ceil(seq / 128)
  long long ld_o;
  void* out;
  float scale_log2;  // (1/sqrt(64)) * log2(e)
  int v_lbo, v_sbo, v_kstep;  // V (MN-major B) descriptor strides in bytes
};

__global__ void __launch_bounds__(192, 1)
attn_spatial_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* kv_full = q_full + 1;    // [2]
  uint64_t* kv_empty = kv_full + 2;  // [2]
  uint64_t* s_full = kv_empty + 2;   // [2]
  uint64_t* p_full = s_full + 2;     // [2]
  uint64_t* o_full = p_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  uint8_t* sQ = smem + 1024;
  uint8_t* sK = sQ + kTileBytes;       // 2 stages
  uint8_t* sV = sK + 2 * kTileBytes;   // 2 stages
  uint8_t* sP = sV + 2 * kTileBytes;   // 2 buffers x 2 sub-tiles of 64 keys

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTile;
  const int head = blockIdx.y;
  const int frame = blockIdx.z;
  const int n_kv = p.n_kv;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;         // columns [0,128) and [128,256)
  const uint32_t tO = tmem_base + 256;   // columns [256,320) and [320,384)

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tmQ, q_full, head * 64, q0, frame);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1, 11);
        mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
        tma_load_3d(sK + st * kTileBytes, &tmK, &kv_full[st], head * 64, j * kTile, frame);
        tma_load_3d(sV + st * kTileBytes, &tmV, &kv_full[st], head * 64, j * kTile, frame);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0, 0);  // S = Q K^T : both K-major
      const uint32_t idesc_o = make_idesc_f16(128, 64, 0, 0, 1);   // O = P V   : B (=V) MN-major
      const uint32_t q_base = smem_u32(sQ), p_base = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1, 12);
        tc_fence_after();
        const uint32_t k_base = smem_u32(sK + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tS + st * 128, make_desc_sw128(q_base + k * 32, 16, 1024),
                   make_desc_sw128(k_base + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0, 13);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(&p_full[st], (j >> 1) & 1, 14);
        tc_fence_after();
        const uint32_t v_base = smem_u32(sV + st * kTileBytes);
        const uint32_t pj_base = p_base + st * 2 * kTileBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A = P: keys [16k, 16k+16) -> sub-tile k/4, 32 B steps inside the 128 B swizzle row
          const uint64_t ad = make_desc_sw128(pj_base + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024);
          // B = V (MN-major): 16 keys = 2 groups of 8 rows (SBO = 1024 B); N = 64 dims = one 128 B atom
          const uint64_t bd = make_desc_sw128(v_base + k * p.v_kstep, p.v_lbo, p.v_sbo);
          umma_f16(tO + st * 64, ad, bd, idesc_o, k != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&o_full[st]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps
    const int r = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    const uint32_t p_row = smem_u32(sP) + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1, 15);
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t* sc = s + c * 32;
        tmem_ld32(tS + lane_off + st * 128 + c * 32, *reinterpret_cast<uint32_t(*)[32]>(sc));
      }
      tmem_ld_wait();
      const int kv_left = p.seq - j * kTile;  // keys valid in this block
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        float x = __uint_as_float(s[i]);
        if (i >= kv_left) x = -INFINITY;
        s[i] = __float_as_uint(x);
        mx = fmaxf(mx, x);
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = ex2_f(m_run - m_new);  // first block: ex2(-inf) = 0
      float rs = 0.f;
      // P = exp2(s*scale - m) -> fp16 -> swizzled K-major smem (A operand of the PV MMA).
      // Buffer st was last read by PV_{j-2}, whose completion (o_full) was observed in iteration j-1.
      const uint32_t p_row_j = p_row + st * 2 * kTileBytes;
#pragma unroll
      for (int c = 0; c < 16; ++c) {  // 16 chunks of 8 keys (16 B)
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p0 = ex2_f(__uint_as_float(s[c * 8 + 2 * i]) * p.scale_log2 - m_new);
          const float p1 = ex2_f(__uint_as_float(s[c * 8 + 2 * i + 1]) * p.scale_log2 - m_new);
          rs += p0 + p1;
          __half2 h = __floats2half2_rn(p0, p1);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        const uint32_t addr = p_row_j + (c >> 3) * kTileBytes + (((c & 7) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(&p_full[st]);
      // fold the previous block's O (computed against the previous max) while PV_j runs, then rescale
      if (j > 0) {
        const int pst = (j - 1) & 1;
        mbar_wait(&o_full[pst], ((j - 1) >> 1) & 1, 16);
        tc_fence_after();
        uint32_t ov[64];
        tmem_ld32(tO + lane_off + pst * 64, *reinterpret_cast<uint32_t(*)[32]>(ov));
        tmem_ld32(tO + lane_off + pst * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(ov + 32));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; ++i) o_acc[i] = (o_acc[i] + __uint_as_float(ov[i])) * alpha;
      }
    }
    // last block's O
    {
      const int pst = (n_kv - 1) & 1;
      mbar_wait(&o_full[pst], ((n_kv - 1) >> 1) & 1, 17);
      tc_fence_after();
      uint32_t ov[64];
      tmem_ld32(tO + lane_off + pst * 64, *reinterpret_cast<uint32_t(*)[32]>(ov));
      tmem_ld32(tO + lane_off + pst * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(ov + 32));
      tmem_ld_wait();
      const float inv = 1.0f / l_run;
#pragma unroll
      for (int i = 0; i < 64; ++i) o_acc[i] = (o_acc[i] + __uint_as_float(ov[i])) * inv;
    }
    if (q0 + r < p.seq) {
      uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + ((long long)frame * p.seq + q0 + r) * p.ld_o + head * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half2 h = __floats2half2_rn(o_acc[c * 8 + 2 * i], o_acc[c * 8 + 2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(op + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vb

extern "C" int b200v_attention_spatial(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v,
                                       int64_t ld_v, void* out, int64_t ld_o, int32_t frames, int32_t seq,
                                       int32_t heads, void* stream_) {
  using namespace vb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VB_REQUIRE(q && k && v && out, "b200v_attention_spatial: null pointer");
  VB_REQUIRE(frames > 0 && seq > 0 && heads > 0, "b200v_attention_spatial: bad sizes");
  VB_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_o % 8 == 0,
             "b200v_attention_spatial: row strides must be multiples of 8 elements");
  CUtensorMap tm[3];
  const void* ptrs[3] = {q, k, v};
  const int64_t lds[3] = {ld_q, ld_k, ld_v};
  for (int i = 0; i < 3; ++i) {
    VB_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "b200v_attention_spatial: unaligned pointer");
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)seq, (uint64_t)frames};
    uint64_t strides[2] = {(uint64_t)lds[i] * 2, (uint64_t)lds[i] * 2 * seq};
    uint32_t box[3] = {64, 128, 1};
    uint32_t es[3] = {1, 1, 1};
    if (encode_tmap_16bit(&tm[i], ptrs[i], 3, dims, strides, box, es, 0)) return 3;
  }
  AttnParams p;
  p.seq = seq;
  p.heads = heads;
  p.n_kv = (seq + kTile - 1) / kTile;
  p.ld_o = ld_o;
  p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  // MN-major SW128: 8-key groups 1024 B apart (SBO); a 16-key MMA step spans 2048 B; LBO (stride between
  // 64-element groups along N) is unused for N = 64.
  p.v_lbo = kTileBytes; p.v_sbo = 1024; p.v_kstep = 2048;
  const int smem_bytes = 1024 + 1024 + 9 * kTileBytes;
  static bool attr_set[64] = {false};
  if (vb::first_use_on_device(attr_set)) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  dim3 grid((seq + kTile - 1) / kTile, heads, frames);
  attn_spatial_kernel<<<grid, 192, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
