// Spatial self-attention v6 (head dim 64, non-causal): v5's structure (persistent CTA, two 128-query tiles, P in tensor
// memory, per-tile MMA issuing threads) with TWO softmax threads per query row:
//   warps 0-7   : softmax of tile A — warp w: lane quadrant w & 3, key half (w >> 2) & 1 (keys 0-63 / 64-127 of a block)
//   warps 8-15  : softmax of tile B
//   warp 16     : TMA producer          warps 17, 18 : tcgen05.mma issuing thread of tile A / B (17 owns the TMEM allocation)
// Why: the exponentials (16 384 per block and tile on the 16 ex2/clk/SM MUFU) are the long pole next to the MMAs, and one
// warp per scheduler (v5's 4-warp group in its ping-pong turn) cannot keep the MUFU fed — measured 13.3 elements/clk with
// 8 warps against 16 peak, ~10 with 4 (profiles/r02_sm_probe.md; v5: XU 65 %, tensor pipe 36 %, profiles/r02_ncu_attn5.md).
// With 8 warps per tile a turn holds two warps per scheduler, each thread keeps 64 scores instead of 128 (no spills at
// 104 registers) and the two halves of a row publish their P chunks in parallel.
// Per block: both halves read their 64 scores (tcgen05.ld), exchange the row maximum through shared memory (named
// barrier per tile, parity double-buffered), decide the lazy O rescale tile-wide (a rescale, rare, is followed by a
// tile barrier so that no PV MMA of the block can start on a half-rescaled accumulator), write P = exp2(...) as packed
// fp16 into columns [32 half, 32 half + 32) of the tile's S range (tcgen05.st), and arrive on the p_full barrier of
// each 32-key chunk.  Everything else (rings, aliasing order, ones-column row sum, epilogue) is v5's.
#include <stdlib.h>

#include "../../include/vista_b200.h"
#include "host.cuh"
#include "ptx.cuh"

#ifndef VB_ATTN6_EXP_DEFAULT
#define VB_ATTN6_EXP_DEFAULT 0
#endif

namespace vb {

constexpr int kT6 = 128;
constexpr int kT6Bytes = 128 * 128;   // 16 KB: 128 rows x 64 fp16
constexpr int kNS6 = 3;               // K / V ring depth

struct Attn6Params {
  int seq;
  int n_kv;                 // key blocks of 128
  int n_qb;                 // query blocks of 256
  int heads;
  int n_items;              // frames * heads * n_qb
  long long ld_o;
  void* out;
  float scale_log2;
  int pingpong;             // 1: the two softmax warpgroups take turns on the exponential phase (MUFU at full rate each)
  int chunked;              // 1: P is published per 32-key chunk (PV overlaps the softmax of the same block); 0: per block
};

__device__ __forceinline__ void tmem6_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem6_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3f6(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
// exp2 on the FMA / ALU pipes: x = n + f, f in [-0.5, 0.5] through the magic-number round, degree-3 minimax polynomial
// for 2^f (max relative error 7.7e-5, below the fp16 rounding of P), exponent patched in with one integer multiply-add.
__device__ __forceinline__ float exp2_poly6(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __uint_as_float(__float_as_uint(t) * 8388608u + __float_as_uint(p));
}
__device__ __forceinline__ uint32_t pack6_h2(float lo, float hi) {
  uint32_t p;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
  return p;
}
__device__ __forceinline__ uint32_t ex2_h26(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int EXP>
__global__ void __launch_bounds__(608, 1)
attn6_spatial_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Attn6Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);   // [2] Q buffers
  uint64_t* q_empty = q_full + 2;                         // [2]
  uint64_t* k_full = q_empty + 2;                         // [kNS6]
  uint64_t* k_empty = k_full + kNS6;
  uint64_t* v_full = k_empty + kNS6;
  uint64_t* v_empty = v_full + kNS6;
  uint64_t* s_full = v_empty + kNS6;                      // [2] per tile
  uint64_t* p_full = s_full + 2;                          // [2 tiles][4 chunks of 32 keys]
  uint64_t* o_full = p_full + 8;                          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  int* need_flag = reinterpret_cast<int*>(smem + 256);              // [2 parities][2 tiles][8 warps], 16-byte aligned
  float* xch = reinterpret_cast<float*>(smem + 512);                // [2 parities][2 tiles][2 halves][128 rows] (4 KB)
  uint8_t* sQ = smem + 5120;                   // 2 buffers x 2 tiles (1024-byte aligned)
  uint8_t* sK = sQ + 4 * kT6Bytes;             // kNS6 stages
  uint8_t* sV = sK + kNS6 * kT6Bytes;          // kNS6 stages
  uint8_t* sOnes = sV + kNS6 * kT6Bytes;       // constant B atom: column 0 = 1, rest 0 (behind every V stage: LBO > 0)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = p.n_kv;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);          // one tcgen05.commit per tile issuer
      mbar_init(&s_full[i], 1);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&p_full[i], 128);
    for (int i = 0; i < kNS6; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    fence_barrier_init();
  }
  // ones atom: row k (128 B) holds fp16 1.0 in logical column 0 -> 16-byte chunk 0 lives at slot (0 ^ (k & 7))
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int row = i >> 3, slot = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (slot == (row & 7)) v.x = 0x00003C00u;
    *reinterpret_cast<uint4*>(sOnes + row * 128 + slot * 16) = v;
  }
  fence_proxy_async_smem();
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S_A / P_A [0,128)  S_B / P_B [128,256)  O_A [256,336)  O_B [384,464)

  // work item -> (frame, head, query block); consecutive items share (frame, head): the SMs that run them at the same
  // time read the same K / V from L2
  auto decode = [&](int item, int& frame, int& head, int& q0) {
    const int qb = item % p.n_qb;
    const int fh = item / p.n_qb;
    head = fh % p.heads;
    frame = fh / p.heads;
    q0 = qb * 2 * kT6;
  };

  if (warp == 16) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t kq = 0, kv = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++kq) {
        int frame, head, q0;
        decode(item, frame, head, q0);
        const int qb = kq & 1;
        mbar_wait_relaxed(&q_empty[qb], ((kq >> 1) & 1) ^ 1, 51);
        mbar_expect_tx(&q_full[qb], 2 * kT6Bytes);
        tma_load_3d(sQ + (2 * qb) * kT6Bytes, &tmQ, &q_full[qb], head * 64, q0, frame);
        tma_load_3d(sQ + (2 * qb + 1) * kT6Bytes, &tmQ, &q_full[qb], head * 64, q0 + kT6, frame);
        for (int j = 0; j < n_kv; ++j, ++kv) {
          const int st = kv % kNS6;
          const uint32_t ph = (kv / kNS6) & 1;
          mbar_wait_relaxed(&k_empty[st], ph ^ 1, 52);
          mbar_expect_tx(&k_full[st], kT6Bytes);
          tma_load_3d(sK + st * kT6Bytes, &tmK, &k_full[st], head * 64, j * kT6, frame);
          mbar_wait_relaxed(&v_empty[st], ph ^ 1, 53);
          mbar_expect_tx(&v_full[st], kT6Bytes);
          tma_load_3d(sV + st * kT6Bytes, &tmV, &v_full[st], head * 64, j * kT6, frame);
        }
      }
    }
  } else if (warp == 17 || warp == 18) {
    // ------------------------------------------------------------ MMA issuers: one thread per tile
    // Each tile has its own issuing thread (warp 9: tile A, warp 10: tile B): plain blocking waits, no polling, and the
    // two tiles never gate each other.  Per block: PV of the 32-key chunks as the softmax warpgroup publishes them
    // (the PV MMAs of a block overlap the rest of its softmax), then S of the next block right behind (the aliased P
    // columns are consumed in order by the tensor pipe).  The K / V / Q "empty" barriers count one tcgen05.commit
    // per tile; the issuer of a skipped tile B makes its arrivals without work, in step with the rings.
    if (lane == 0) {
      const int t = warp - 17;
      const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, 80, 0, 0, 1);   // B = [V | ones] MN-major, N = 64 + 16
      const uint32_t ones_base = smem_u32(sOnes);
      uint32_t kq = 0, kv = 0, g = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++kq) {
        int frame, head, q0;
        decode(item, frame, head, q0);
        const bool on = t == 0 || (q0 + kT6 < p.seq);    // tile B entirely beyond the sequence: no work
        const int qb = kq & 1;
        auto issue_s = [&](uint32_t kvi) {
          const uint32_t q_base = smem_u32(sQ + (2 * qb + t) * kT6Bytes);
          const uint32_t k_base = smem_u32(sK + (kvi % kNS6) * kT6Bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + t * 128, make_desc_sw128(q_base + k * 32, 16, 1024),
                     make_desc_sw128(k_base + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
        };
        mbar_wait(&q_full[qb], (kq >> 1) & 1, 54);
        mbar_wait(&k_full[kv % kNS6], (kv / kNS6) & 1, 55);
        tc_fence_after();
        if (on) issue_s(kv);
        umma_commit(&k_empty[kv % kNS6]);
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t cur = kv + j;
          const int st = cur % kNS6;
          mbar_wait(&v_full[st], (cur / kNS6) & 1, 57);
          if (on) {
            const uint32_t v_base = smem_u32(sV + st * kT6Bytes);
            const uint32_t lbo = ones_base - v_base;     // second N atom (columns 64..79) = the ones atom
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              mbar_wait(&p_full[t * 4 + c], g & 1, 56);
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {           // keys 16 k .. 16 k + 15: 8 packed columns of P
                const int k = 2 * c + kk;
                umma_f16_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + k * 8,
                            make_desc_sw128(v_base + k * 2048, lbo, 1024), idesc_o, (j | k) != 0 ? 1u : 0u);
              }
            }
            ++g;
          }
          umma_commit(&v_empty[st]);
          if (j + 1 == n_kv) {
            if (on) umma_commit(&o_full[t]);
          } else {
            const int sn = (cur + 1) % kNS6;
            mbar_wait(&k_full[sn], ((cur + 1) / kNS6) & 1, 58);
            tc_fence_after();
            if (on) issue_s(cur + 1);
            umma_commit(&k_empty[sn]);
          }
        }
        umma_commit(&q_empty[qb]);      // every S MMA of this tile and item has read Q
        kv += n_kv;
      }
    }
  } else {
    // ------------------------------------------------------------ softmax: 8 warps per tile, 2 threads per row
    const int t = warp >> 3;                       // tile 0 / 1
    const int hf = (warp >> 2) & 1;                // key half of the block this thread owns
    const int wq = warp & 3;                       // TMEM lane quadrant (== warp % 4: the hardware's access rule)
    const int r = wq * 32 + lane;                  // row in the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_off;
    const int bar_x = 3 + t, bar_r = 5 + t;        // named barriers of the tile: maximum exchange / after a rescale
    uint32_t g = 0, items_done = 0;
    if (p.pingpong && t == 1) asm volatile("bar.arrive 1, 512;" ::: "memory");     // tile A goes first
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int frame, head, q0;
      decode(item, frame, head, q0);
      if (t == 1 && q0 + kT6 >= p.seq) {          // tile B has no rows: it still passes the turn back, block by block
        if (p.pingpong) {
          for (int j = 0; j < n_kv; ++j) {
            asm volatile("bar.sync 2, 512;" ::: "memory");
            asm volatile("bar.arrive 1, 512;" ::: "memory");
          }
        }
        continue;
      }
      float m_used = -INFINITY;
      for (int j = 0; j < n_kv; ++j, ++g) {
        const int par = g & 1;
        mbar_wait(&s_full[t], par, 59);     // also implies PV_t(j-1) has completed (in-order commits)
        tc_fence_after();
        uint32_t s[64];
        tmem_ld32(tS + hf * 64, *reinterpret_cast<uint32_t(*)[32]>(s));
        tmem_ld32(tS + hf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(s + 32));
        tmem_ld_wait();
        const int kv_left = p.seq - j * kT6 - hf * 64;
        if (kv_left < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= kv_left) s[i] = 0xFF800000u;  // -inf
        }
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mxa[u] = max3f6(mxa[u], __uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1]));
        }
        const float mx_half = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
        // exchange with the thread that holds the other 64 keys of this row; the tile-wide rescale vote rides along
        // (need of a row = need of one of its halves, since the row maximum is the larger half maximum)
        float* xrow = xch + ((par * 2 + t) * 2) * 128;
        xrow[hf * 128 + r] = mx_half;
        const bool need_half = mx_half * p.scale_log2 > m_used + 8.0f;
        const bool warp_need = __any_sync(0xffffffffu, need_half);
        if (lane == 0) need_flag[(par * 2 + t) * 8 + (warp & 7)] = warp_need ? 1 : 0;
        asm volatile("bar.sync %0, 256;" ::"r"(bar_x) : "memory");
        const float mx = fmaxf(mx_half, xrow[(hf ^ 1) * 128 + r]);
        const int4 f0 = *reinterpret_cast<const int4*>(need_flag + (par * 2 + t) * 8);
        const int4 f1 = *reinterpret_cast<const int4*>(need_flag + (par * 2 + t) * 8 + 4);
        const bool tile_need = (f0.x | f0.y | f0.z | f0.w | f1.x | f1.y | f1.z | f1.w) != 0;
        const float m_blk = mx * p.scale_log2;
        // lazy rescale: only when the block maximum exceeds the maximum in use by more than 8 (factor 256)
        if (tile_need) {
          const bool need = m_blk > m_used + 8.0f;
          const float m_new = need ? m_blk : m_used;
          if (j > 0) {
            const float alpha = need ? ex2_f(m_used - m_new) : 1.0f;
            // 80 accumulator columns (64 dims + row sum + 15 unused): half 0 takes 48, half 1 the other 32
            const int c_lo = hf ? 3 : 0, c_hi = hf ? 5 : 3;
            for (int c = c_lo; c < c_hi; ++c) {
              uint32_t ov[16];
              tmem_ld16(tO + c * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem6_st16(tO + c * 16, ov);
            }
            tmem6_st_wait();
            tc_fence_before();
            // no PV MMA of this block may start on a half-rescaled accumulator: both halves of every row first
            asm volatile("bar.sync %0, 256;" ::"r"(bar_r) : "memory");
          }
          m_used = m_new;
        }
        // Ping-pong: the exponentials of the two tiles alternate (named barriers 1 / 2 = "A's turn" / "B's turn").
        if (p.pingpong) {
          if (t == 0) asm volatile("bar.sync 1, 512;" ::: "memory");
          else asm volatile("bar.sync 2, 512;" ::: "memory");
        }
        // P = exp2(s * scale - m_used) -> packed fp16 -> columns [32 hf, 32 hf + 32) of the tile's S range
#pragma unroll
        for (int c = 0; c < 2; ++c) {      // 32 keys -> 16 packed columns: chunk 2 hf + c of the block
          uint32_t pw[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int e = c * 32 + 2 * i;
            const float x0 = fmaf(__uint_as_float(s[e]), p.scale_log2, -m_used);
            const float x1 = fmaf(__uint_as_float(s[e + 1]), p.scale_log2, -m_used);
            if (EXP == 1) {
              pw[i] = ex2_h26(pack6_h2(x0, x1));
            } else {
              const float p0 = ((e & 7) < EXP) ? exp2_poly6(x0) : ex2_f(x0);
              const float p1 = (((e + 1) & 7) < EXP) ? exp2_poly6(x1) : ex2_f(x1);
              pw[i] = pack6_h2(p0, p1);
            }
          }
          if (c > 0 && p.chunked) {        // the first chunk's store completed meanwhile: publish it
            tmem6_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[t * 4 + 2 * hf]);
          }
          tmem6_st16(tS + hf * 32 + c * 16, pw);
        }
        if (p.pingpong) {                  // the other tile's turn
          if (t == 0) asm volatile("bar.arrive 2, 512;" ::: "memory");
          else asm volatile("bar.arrive 1, 512;" ::: "memory");
        }
        tmem6_st_wait();
        tc_fence_before();
        if (!p.chunked) mbar_arrive(&p_full[t * 4 + 2 * hf]);
        mbar_arrive(&p_full[t * 4 + 2 * hf + 1]);
      }
      // epilogue: O / rowsum; each half writes 32 of the 64 dims
      mbar_wait(&o_full[t], items_done & 1, 60);
      ++items_done;
      tc_fence_after();
      uint32_t ov[32], lv[16];
      tmem_ld32(tO + hf * 32, *reinterpret_cast<uint32_t(*)[32]>(ov));
      tmem_ld16(tO + 64, lv);
      tmem_ld_wait();
      const float inv = 1.0f / __uint_as_float(lv[0]);
      const int q = q0 + t * kT6 + r;
      if (q < p.seq) {
        uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + ((long long)frame * p.seq + q) * p.ld_o + head * 64 + hf * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack6_h2(__uint_as_float(ov[c * 8 + 2 * i]) * inv, __uint_as_float(ov[c * 8 + 2 * i + 1]) * inv);
          *reinterpret_cast<uint4*>(op + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vb

extern "C" int b200v_attention_spatial_v6(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v,
                                          int64_t ld_v, void* out, int64_t ld_o, int32_t frames, int32_t seq,
                                          int32_t heads, void* stream_) {
  using namespace vb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VB_REQUIRE(q && k && v && out, "b200v_attention_spatial_v6: null pointer");
  VB_REQUIRE(frames > 0 && seq > 0 && heads > 0, "b200v_attention_spatial_v6: bad sizes");
  VB_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_o % 8 == 0,
             "b200v_attention_spatial_v6: row strides must be multiples of 8 elements");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b200v_attention_spatial_v6: unaligned output");
  CUtensorMap tm[3];
  const void* ptrs[3] = {q, k, v};
  const int64_t lds[3] = {ld_q, ld_k, ld_v};
  for (int i = 0; i < 3; ++i) {
    VB_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "b200v_attention_spatial_v6: unaligned pointer");
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)seq, (uint64_t)frames};
    uint64_t strides[2] = {(uint64_t)lds[i] * 2, (uint64_t)lds[i] * 2 * seq};
    uint32_t box[3] = {64, 128, 1};
    uint32_t es[3] = {1, 1, 1};
    if (encode_tmap_16bit(&tm[i], ptrs[i], 3, dims, strides, box, es, 0)) return 3;
  }
  Attn6Params p;
  p.seq = seq;
  p.n_kv = (seq + kT6 - 1) / kT6;
  p.n_qb = (seq + 2 * kT6 - 1) / (2 * kT6);
  p.heads = heads;
  const long long items = (long long)frames * heads * p.n_qb;
  VB_REQUIRE(items < (1ll << 31), "b200v_attention_spatial_v6: too many work items");
  p.n_items = (int)items;
  p.ld_o = ld_o;
  p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static int chunked = -1;
  if (chunked < 0) chunked = getenv("VB_ATTN6_CHUNKED") ? atoi(getenv("VB_ATTN6_CHUNKED")) : 1;
  p.chunked = chunked;
  static int pingpong = -1;
  if (pingpong < 0) pingpong = getenv("VB_ATTN6_PINGPONG") ? atoi(getenv("VB_ATTN6_PINGPONG")) : 1;
  p.pingpong = pingpong;
  const int smem_bytes = 1024 + 5120 + (4 + 2 * kNS6 + 1) * kT6Bytes;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Attn6Params);
  static const Kern kerns[5] = {attn6_spatial_kernel<0>, attn6_spatial_kernel<1>, attn6_spatial_kernel<2>,
                                attn6_spatial_kernel<3>, attn6_spatial_kernel<4>};
  static bool attr_set[64] = {false};
  int dev = 0;
  VB_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    for (int i = 0; i < 5; ++i)
      VB_CHECK_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set[dev] = true;
  }
  static int mode = -1;
  if (mode < 0) {
    mode = VB_ATTN6_EXP_DEFAULT;
    if (const char* e = getenv("VB_ATTN6_EXP")) mode = atoi(e);
    if (mode < 0 || mode > 4) mode = VB_ATTN6_EXP_DEFAULT;
  }
  int grid = device_sm_count();
  if (items < grid) grid = (int)items;
  kerns[mode]<<<grid, 608, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
