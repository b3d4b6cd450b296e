// Spatial self-attention for SHORT sequences (head dim 64, non-causal): the third-generation kernel (one 128-query tile per
// CTA, two CTAs per SM, two softmax threads per row, P through swizzled shared memory, O and the row sum in tensor memory
// with lazy rescaling).  Long sequences run the persistent kernel of attn7_tc.cu.  Generations 1, 2, 4, 5, 6 are gone from
// the library; their measurements are in profiles/ (r01_step_breakdown_v*.md, r02_attn*_variant_matrix.txt).
#include <stdlib.h>

#include "../../include/vista_b200.h"
#include "host.cuh"
#include "ptx.cuh"

namespace vb {

constexpr int kT2 = 128;
constexpr int kT2Bytes = 128 * 128;  // 16 KB: 128 rows x 64 fp16

struct Attn2Params {
  int seq;
  int n_kv;
  long long ld_o;
  void* out;
  float scale_log2;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
// exp2 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (max relative error 7.7e-5, below the fp16 rounding of P), exponent patched in with one
// shift-add.  Used for a fraction of the elements so that the MUFU (16 ex2/clk/SM) stops being the bound.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#ifndef VB_ATTN_POLY_OF_4
#define VB_ATTN_POLY_OF_4 2   // how many of every 4 exponentials go to the polynomial path
#endif
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t p;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
  return p;
}

}  // namespace vb

// =================================================================================================
// v3: one 128-query tile per CTA, TWO CTAs per SM (98 KB smem, 256 TMEM columns each), eight softmax
// warps per CTA: warp w serves TMEM lane quadrant w % 4 and the key half w / 4 (64 of the 128 keys of a
// block), i.e. two threads per query row.  Four softmax warps per SM sub-partition hide each other's
// latencies; the two co-resident CTAs interleave their MMA / softmax phases without any coupling.
// Row maxima are exchanged between the two threads of a row through shared memory (one named barrier
// per block); O and the row sum stay in TMEM (ones-column) with lazy rescaling as in v2; S is consumed in
// 32-column chunks (max pass, then exp pass) to fit the 80-register budget of 2 x 12 warps per SM.
// =================================================================================================
namespace vb {

#ifndef VB_ATTN3_POLY_OF_8
#define VB_ATTN3_POLY_OF_8 0   // of every 8 exponentials, this many use exp2_poly (FMA pipe) instead of MUFU
#endif

// TS = true (the dropped v4; not instantiated): P does not go through shared memory.  The softmax threads write it as
// packed fp16 into the TMEM columns of their own (already consumed) half of S with tcgen05.st, and O += P V is issued
// with the A operand in TMEM.  The single-thread issue order keeps S(j+1) behind the PV(j) that reads the aliased
// columns.  Motivation: with shared-memory operands every MMA costs ~90 ns however small N is (profiles/
// r01_umma_n_sweep.md); the 8 PV MMAs (N = 80) are two thirds of the kernel's tensor time.
template <int POLY8, bool TS>
__global__ void __launch_bounds__(384, 2)
attn3_spatial_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = q_full + 2;
  uint64_t* v_full = q_full + 3;
  uint64_t* v_empty = q_full + 4;
  uint64_t* s_full = q_full + 5;
  uint64_t* p_full = q_full + 6;
  uint64_t* o_full = q_full + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 8);
  float* mx_ex = reinterpret_cast<float*>(smem + 128);   // [2 parities][2 halves][128 rows] = 2 KB
  uint8_t* sQ = smem + 3072;
  uint8_t* sK = sQ + kT2Bytes;
  uint8_t* sV = sK + kT2Bytes;
  uint8_t* sOnes = sV + kT2Bytes;
  uint8_t* sP = sOnes + kT2Bytes;   // 2 sub-tiles of 64 keys

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kT2;
  const int head = blockIdx.y;
  const int frame = blockIdx.z;
  const int n_kv = p.n_kv;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int row = i >> 3, slot = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (slot == (row & 7)) v.x = 0x00003C00u;
    *reinterpret_cast<uint4*>(sOnes + row * 128 + slot * 16) = v;
  }
  fence_proxy_async_smem();
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // S [0,128)  O [128,208)

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kT2Bytes);
      tma_load_3d(sQ, &tmQ, q_full, head * 64, q0, frame);
      for (int j = 0; j < n_kv; ++j) {
        if (j > 0) mbar_wait(k_empty, (j - 1) & 1, 31);
        mbar_expect_tx(k_full, kT2Bytes);
        tma_load_3d(sK, &tmK, k_full, head * 64, j * kT2, frame);
        if (j > 0) mbar_wait(v_empty, (j - 1) & 1, 32);
        mbar_expect_tx(v_full, kT2Bytes);
        tma_load_3d(sV, &tmV, v_full, head * 64, j * kT2, frame);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, 80, 0, 0, 1);
      const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV), p_base = smem_u32(sP);
      const uint32_t lbo = smem_u32(sOnes) - v_base;
      auto issue_s = [&](int j) {
        mbar_wait(k_full, j & 1, 33);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base, make_desc_sw128(q_base + k * 32, 16, 1024), make_desc_sw128(k_base + k * 32, 16, 1024),
                   idesc_s, k != 0 ? 1u : 0u);
        umma_commit(k_empty);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0, 34);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(p_full, j & 1, 35);
        mbar_wait(v_full, j & 1, 36);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (TS) {
            // keys 16 k .. 16 k + 15 of half k / 4: 8 packed columns inside that half's S range
            const uint64_t bd = make_desc_sw128(v_base + k * 2048, lbo, 1024);
            umma_f16_ts(tmem_base + 128, tmem_base + (k >> 2) * 64 + (k & 3) * 8, bd, idesc_o, (j | k) != 0 ? 1u : 0u);
          } else {
            const uint64_t ad = make_desc_sw128(p_base + (k >> 2) * kT2Bytes + (k & 3) * 32, 16, 1024);
            const uint64_t bd = make_desc_sw128(v_base + k * 2048, lbo, 1024);
            umma_f16(tmem_base + 128, ad, bd, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(v_empty);
        umma_commit(o_full);
        if (j + 1 < n_kv) issue_s(j + 1);
      }
    }
  } else if (warp < 8) {
    const int quad = warp & 3, hcol = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + hcol * 64;
    const uint32_t tO = tmem_base + 128 + lane_off;
    const uint32_t p_row = smem_u32(sP) + hcol * kT2Bytes + r * 128;
    const int sw = r & 7;
    float m_used = -INFINITY;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1, 37);   // implies PV(j-1) has completed (in-order commits)
      tc_fence_after();
      const int kv_left = p.seq - j * kT2 - hcol * 64;   // valid keys in this thread's 64-key half
      // ---- pass 1: maximum of this half (key masking only in the last, partial block)
      const bool tail = kv_left < 64;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[32];
        tmem_ld32(tS + c * 32, s);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= kv_left) s[i] = 0xFF800000u;
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) m4[u] = max3(m4[u], __uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1]));
        }
        mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      }
      float* ex = mx_ex + (j & 1) * 256;
      ex[hcol * 128 + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, ex[(hcol ^ 1) * 128 + r]);
      const float m_blk = mx * p.scale_log2;
      const bool need = m_blk > m_used + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? m_blk : m_used;
        if (j > 0) {
          const float alpha = need ? ex2_f(m_used - m_new) : 1.0f;
          // half 0 rescales accumulator columns [0,48), half 1 columns [48,80)
          const int c_lo = hcol ? 3 : 0, c_hi = hcol ? 5 : 3;
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t ov[16];
            tmem_ld16(tO + c * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st16(tO + c * 16, ov);
          }
          tmem_st_wait();
        }
        m_used = m_new;
      }
      // ---- pass 2: P = exp2(s*scale - m_used) -> fp16 -> swizzled smem (sub-tile hcol)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[32];
        tmem_ld32(tS + c * 32, s);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= kv_left) s[i] = 0xFF800000u;
        }
        uint32_t pw[16];                   // TS: the 32 keys of this chunk as 16 packed fp16 pairs
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {   // 8 keys = one 16-byte chunk
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e0 = q4 * 8 + 2 * i, e1 = e0 + 1;
            const float x0 = fmaf(__uint_as_float(s[e0]), p.scale_log2, -m_used);
            const float x1 = fmaf(__uint_as_float(s[e1]), p.scale_log2, -m_used);
            const float p0 = ((e0 & 7) < POLY8) ? exp2_poly(x0) : ex2_f(x0);
            const float p1 = ((e1 & 7) < POLY8) ? exp2_poly(x1) : ex2_f(x1);
            w[i] = pack_h2(p0, p1);
            pw[q4 * 4 + i] = w[i];
          }
          if (!TS) {
            const int chunk = c * 4 + q4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + ((chunk ^ sw) << 4)), "r"(w[0]), "r"(w[1]),
                         "r"(w[2]), "r"(w[3])
                         : "memory");
          }
        }
        // keys [32 c, 32 c + 32) of this half -> packed columns [16 c, 16 c + 16) of the half's own S range: those S
        // columns were consumed in this or the previous chunk iteration
        if (TS) tmem_st16(tS + c * 16, pw);
      }
      if (TS) tmem_st_wait(); else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: this thread writes dims [32*hcol, 32*hcol + 32) of its row
    mbar_wait(o_full, (n_kv - 1) & 1, 38);
    tc_fence_after();
    uint32_t ov[32], lv[16];
    tmem_ld32(tO + hcol * 32, ov);
    tmem_ld16(tO + 64, lv);
    tmem_ld_wait();
    const float inv = 1.0f / __uint_as_float(lv[0]);
    const int q = q0 + r;
    if (q < p.seq) {
      uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + ((long long)frame * p.seq + q) * p.ld_o + head * 64 + hcol * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = pack_h2(__uint_as_float(ov[c * 8 + 2 * i]) * inv, __uint_as_float(ov[c * 8 + 2 * i + 1]) * inv);
        *reinterpret_cast<uint4*>(op + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace vb

static int attn3_launch(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v, int64_t ld_v, void* out,
                        int64_t ld_o, int32_t frames, int32_t seq, int32_t heads, void* stream_) {
  using namespace vb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VB_REQUIRE(q && k && v && out, "b200v_attention_spatial_v3: null pointer");
  VB_REQUIRE(frames > 0 && seq > 0 && heads > 0, "b200v_attention_spatial_v3: bad sizes");
  VB_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_o % 8 == 0,
             "b200v_attention_spatial_v3: row strides must be multiples of 8 elements");
  CUtensorMap tm[3];
  const void* ptrs[3] = {q, k, v};
  const int64_t lds[3] = {ld_q, ld_k, ld_v};
  for (int i = 0; i < 3; ++i) {
    VB_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "b200v_attention_spatial_v3: unaligned pointer");
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)seq, (uint64_t)frames};
    uint64_t strides[2] = {(uint64_t)lds[i] * 2, (uint64_t)lds[i] * 2 * seq};
    uint32_t box[3] = {64, 128, 1};
    uint32_t es[3] = {1, 1, 1};
    if (encode_tmap_16bit(&tm[i], ptrs[i], 3, dims, strides, box, es, 0)) return 3;
  }
  Attn2Params p;
  p.seq = seq;
  p.n_kv = (seq + kT2 - 1) / kT2;
  p.ld_o = ld_o;
  p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  const int smem_bytes = 1024 + 3072 + 6 * kT2Bytes;
  static int poly = -1;
  if (poly < 0) {
    poly = VB_ATTN3_POLY_OF_8;
    if (const char* e = getenv("VB_ATTN3_POLY")) poly = atoi(e);   // tuning knob: exponentials (of 8) on the FMA pipe
    VB_CHECK_CUDA(cudaFuncSetAttribute(attn3_spatial_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    VB_CHECK_CUDA(cudaFuncSetAttribute(attn3_spatial_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    VB_CHECK_CUDA(cudaFuncSetAttribute(attn3_spatial_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    VB_CHECK_CUDA(cudaFuncSetAttribute(attn3_spatial_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  dim3 grid((seq + kT2 - 1) / kT2, heads, frames);
  if (poly <= 0) attn3_spatial_kernel<0, false><<<grid, 384, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  else if (poly <= 2) attn3_spatial_kernel<2, false><<<grid, 384, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  else if (poly == 3) attn3_spatial_kernel<3, false><<<grid, 384, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  else attn3_spatial_kernel<4, false><<<grid, 384, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200v_attention_spatial_v3(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v,
                                          int64_t ld_v, void* out, int64_t ld_o, int32_t frames, int32_t seq,
                                          int32_t heads, void* stream) {
  return attn3_launch(q, ld_q, k, ld_k, v, ld_v, out, ld_o, frames, seq, heads, stream);
}
