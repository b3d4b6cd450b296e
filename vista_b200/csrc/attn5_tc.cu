// Spatial self-attention v5 (head dim 64, non-causal): persistent, two 128-query tiles per CTA, P in tensor memory.
//   warps 0-3 : softmax warpgroup of tile A (thread = query row = TMEM lane)
//   warps 4-7 : softmax warpgroup of tile B
//   warp 8    : TMA producer (Q of the next work item into the other Q buffer; K / V blocks of 128 keys, NS-deep rings)
//   warps 9, 10 : one tcgen05.mma issuing thread per tile (warp 9 also owns the TMEM allocation)
// One CTA per SM walks work items (frame, head, 256-query block) round robin; the TMA rings, the TMEM allocation and
// the constant "ones" operand atom live across items, so the prologue of an item (Q load, first K block) hides behind
// the tail of the previous one.
// Per KV block and tile:  S = Q K^T (M128 N128 K64, both operands in shared memory) -> the row's thread reads its 128
// scores once (tcgen05.ld), P = exp2(s * scale - m) goes back as packed fp16 into the first 64 columns of the SAME
// tensor-memory range (tcgen05.st: no shared-memory round trip, no proxy fence), then
//   O += P [V | 1]  (M128 N80 K128, A operand in tensor memory, B = the V TMA tile as MN-major atom + a constant "ones"
// atom): the accumulator and the row sum (column 64: the sum of exactly the fp16 P values that multiplied V) stay in
// tensor memory across blocks.  The issue order S_A(j+1) after PV_A(j) keeps the aliasing safe (the tensor pipe executes
// in order).  While warpgroup A runs the softmax of block j the tensor pipe works on tile B and vice versa.
// The running maximum is applied lazily: O (and with it the row sum) is rescaled only when the block maximum exceeds the
// maximum in use by more than 2^8; P stays below 2^8 otherwise (fine in fp16).
// EXP selects how the 16 384 exponentials per block and tile are evaluated (the MUFU, 16 ex2 / clk / SM, is the bound
// of this kernel at head dim 64):  0 = ex2.approx.f32;  1 = ex2.approx.f16x2 on the packed (x0, x1) pair (the argument is
// rounded to fp16: |x| < 2 -> error below the fp16 rounding of P itself);  2..4 = that many of every 8 on the FMA pipe
// (degree-3 polynomial), the rest ex2.approx.f32.
#include <stdlib.h>

#include "../../include/vista_b200.h"
#include "host.cuh"
#include "ptx.cuh"

#ifndef VB_ATTN5_EXP_DEFAULT
#define VB_ATTN5_EXP_DEFAULT 0
#endif

namespace vb {

constexpr int kT5 = 128;
constexpr int kT5Bytes = 128 * 128;   // 16 KB: 128 rows x 64 fp16
constexpr int kNS5 = 3;               // K / V ring depth

struct Attn5Params {
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

__device__ __forceinline__ void tmem5_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem5_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
// exp2 on the FMA / ALU pipes: x = n + f, f in [-0.5, 0.5] through the magic-number round, degree-3 minimax polynomial
// for 2^f (max relative error 7.7e-5, below the fp16 rounding of P), exponent patched in with one integer multiply-add.
__device__ __forceinline__ float exp2_poly5(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __uint_as_float(__float_as_uint(t) * 8388608u + __float_as_uint(p));
}
__device__ __forceinline__ uint32_t pack5_h2(float lo, float hi) {
  uint32_t p;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
  return p;
}
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int EXP>
__global__ void __launch_bounds__(352, 1)
attn5_spatial_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Attn5Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);   // [2] Q buffers
  uint64_t* q_empty = q_full + 2;                         // [2]
  uint64_t* k_full = q_empty + 2;                         // [kNS5]
  uint64_t* k_empty = k_full + kNS5;
  uint64_t* v_full = k_empty + kNS5;
  uint64_t* v_empty = v_full + kNS5;
  uint64_t* s_full = v_empty + kNS5;                      // [2] per tile
  uint64_t* p_full = s_full + 2;                          // [2 tiles][4 chunks of 32 keys]
  uint64_t* o_full = p_full + 8;                          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  uint8_t* sQ = smem + 1024;                   // 2 buffers x 2 tiles
  uint8_t* sK = sQ + 4 * kT5Bytes;             // kNS5 stages
  uint8_t* sV = sK + kNS5 * kT5Bytes;          // kNS5 stages
  uint8_t* sOnes = sV + kNS5 * kT5Bytes;       // constant B atom: column 0 = 1, rest 0 (behind every V stage: LBO > 0)

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
    for (int i = 0; i < kNS5; ++i) {
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
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
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
    q0 = qb * 2 * kT5;
  };

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t kq = 0, kv = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++kq) {
        int frame, head, q0;
        decode(item, frame, head, q0);
        const int qb = kq & 1;
        mbar_wait_relaxed(&q_empty[qb], ((kq >> 1) & 1) ^ 1, 51);
        mbar_expect_tx(&q_full[qb], 2 * kT5Bytes);
        tma_load_3d(sQ + (2 * qb) * kT5Bytes, &tmQ, &q_full[qb], head * 64, q0, frame);
        tma_load_3d(sQ + (2 * qb + 1) * kT5Bytes, &tmQ, &q_full[qb], head * 64, q0 + kT5, frame);
        for (int j = 0; j < n_kv; ++j, ++kv) {
          const int st = kv % kNS5;
          const uint32_t ph = (kv / kNS5) & 1;
          mbar_wait_relaxed(&k_empty[st], ph ^ 1, 52);
          mbar_expect_tx(&k_full[st], kT5Bytes);
          tma_load_3d(sK + st * kT5Bytes, &tmK, &k_full[st], head * 64, j * kT5, frame);
          mbar_wait_relaxed(&v_empty[st], ph ^ 1, 53);
          mbar_expect_tx(&v_full[st], kT5Bytes);
          tma_load_3d(sV + st * kT5Bytes, &tmV, &v_full[st], head * 64, j * kT5, frame);
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------------------------------------------------ MMA issuers: one thread per tile
    // Each tile has its own issuing thread (warp 9: tile A, warp 10: tile B): plain blocking waits, no polling, and the
    // two tiles never gate each other.  Per block: PV of the 32-key chunks as the softmax warpgroup publishes them
    // (the PV MMAs of a block overlap the rest of its softmax), then S of the next block right behind (the aliased P
    // columns are consumed in order by the tensor pipe).  The K / V / Q "empty" barriers count one tcgen05.commit
    // per tile; the issuer of a skipped tile B makes its arrivals without work, in step with the rings.
    if (lane == 0) {
      const int t = warp - 9;
      const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, 80, 0, 0, 1);   // B = [V | ones] MN-major, N = 64 + 16
      const uint32_t ones_base = smem_u32(sOnes);
      uint32_t kq = 0, kv = 0, g = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++kq) {
        int frame, head, q0;
        decode(item, frame, head, q0);
        const bool on = t == 0 || (q0 + kT5 < p.seq);    // tile B entirely beyond the sequence: no work
        const int qb = kq & 1;
        auto issue_s = [&](uint32_t kvi) {
          const uint32_t q_base = smem_u32(sQ + (2 * qb + t) * kT5Bytes);
          const uint32_t k_base = smem_u32(sK + (kvi % kNS5) * kT5Bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + t * 128, make_desc_sw128(q_base + k * 32, 16, 1024),
                     make_desc_sw128(k_base + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
        };
        mbar_wait(&q_full[qb], (kq >> 1) & 1, 54);
        mbar_wait(&k_full[kv % kNS5], (kv / kNS5) & 1, 55);
        tc_fence_after();
        if (on) issue_s(kv);
        umma_commit(&k_empty[kv % kNS5]);
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t cur = kv + j;
          const int st = cur % kNS5;
          mbar_wait(&v_full[st], (cur / kNS5) & 1, 57);
          if (on) {
            const uint32_t v_base = smem_u32(sV + st * kT5Bytes);
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
            const int sn = (cur + 1) % kNS5;
            mbar_wait(&k_full[sn], ((cur + 1) / kNS5) & 1, 58);
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
    // ------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;                       // tile 0 / 1
    const int r = (warp & 3) * 32 + lane;          // row in the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_off;
    uint32_t g = 0, items_done = 0;
    if (p.pingpong && t == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");     // tile A goes first
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int frame, head, q0;
      decode(item, frame, head, q0);
      if (t == 1 && q0 + kT5 >= p.seq) {          // tile B has no rows: it still passes the turn back, block by block
        if (p.pingpong) {
          for (int j = 0; j < n_kv; ++j) {
            asm volatile("bar.sync 2, 256;" ::: "memory");
            asm volatile("bar.arrive 1, 256;" ::: "memory");
          }
        }
        continue;
      }
      float m_used = -INFINITY;
      for (int j = 0; j < n_kv; ++j, ++g) {
        mbar_wait(&s_full[t], g & 1, 59);   // also implies PV_t(j-1) has completed (in-order commits)
        tc_fence_after();
        uint32_t s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(s + c * 32));
        tmem_ld_wait();
        const int kv_left = p.seq - j * kT5;
        if (kv_left < kT5) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_left) s[i] = 0xFF800000u;  // -inf
        }
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mxa[u] = max3f(mxa[u], __uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1]));
        }
        const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
        const float m_blk = mx * p.scale_log2;
        // lazy rescale: only when the block maximum exceeds the maximum in use by more than 8 (factor 256)
        const bool need = m_blk > m_used + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? m_blk : m_used;
          if (j > 0) {
            const float alpha = need ? ex2_f(m_used - m_new) : 1.0f;
#pragma unroll
            for (int c = 0; c < 5; ++c) {   // 80 accumulator columns: 64 dims + row sum (+15 unused)
              uint32_t ov[16];
              tmem_ld16(tO + c * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem5_st16(tO + c * 16, ov);
            }
          }
          m_used = m_new;
        }
        // Ping-pong: the exponentials of the two tiles alternate (named barriers 1 / 2 = "A's turn" / "B's turn"), so that
        // each warpgroup has the MUFU to itself for its 16 384 ex2 while the other one waits for its PV / S MMAs — left
        // alone the two symmetric tiles run in phase and share the MUFU half / half during the same stretch.
        if (p.pingpong) {
          if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        // P = exp2(s * scale - m_used) -> packed fp16 -> columns [0, 64) of this row's own S range
#pragma unroll
        for (int c = 0; c < 4; ++c) {      // 32 keys -> 16 packed columns
          uint32_t pw[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int e = c * 32 + 2 * i;
            const float x0 = fmaf(__uint_as_float(s[e]), p.scale_log2, -m_used);
            const float x1 = fmaf(__uint_as_float(s[e + 1]), p.scale_log2, -m_used);
            if (EXP == 1) {
              pw[i] = ex2_h2(pack5_h2(x0, x1));
            } else {
              const float p0 = ((e & 7) < EXP) ? exp2_poly5(x0) : ex2_f(x0);
              const float p1 = (((e + 1) & 7) < EXP) ? exp2_poly5(x1) : ex2_f(x1);
              pw[i] = pack5_h2(p0, p1);
            }
          }
          // the store of chunk c - 1 (and the O rescale stores above) completed while chunk c was being computed: wait for it
          // BEFORE the next store is issued (tcgen05.wait::st covers every earlier store), then publish chunk c - 1
          if (c > 0 && p.chunked) {
            tmem5_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[t * 4 + c - 1]);
          }
          tmem5_st16(tS + c * 16, pw);
        }
        if (p.pingpong) {                  // the other tile's turn
          if (t == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
          else asm volatile("bar.arrive 1, 256;" ::: "memory");
        }
        tmem5_st_wait();
        tc_fence_before();
        if (!p.chunked) {
#pragma unroll
          for (int c = 0; c < 3; ++c) mbar_arrive(&p_full[t * 4 + c]);
        }
        mbar_arrive(&p_full[t * 4 + 3]);
      }
      // epilogue: O / rowsum
      mbar_wait(&o_full[t], items_done & 1, 60);
      ++items_done;
      tc_fence_after();
      uint32_t ov[64], lv[16];
      tmem_ld32(tO, *reinterpret_cast<uint32_t(*)[32]>(ov));
      tmem_ld32(tO + 32, *reinterpret_cast<uint32_t(*)[32]>(ov + 32));
      tmem_ld16(tO + 64, lv);
      tmem_ld_wait();
      const float inv = 1.0f / __uint_as_float(lv[0]);
      const int q = q0 + t * kT5 + r;
      if (q < p.seq) {
        uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + ((long long)frame * p.seq + q) * p.ld_o + head * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack5_h2(__uint_as_float(ov[c * 8 + 2 * i]) * inv, __uint_as_float(ov[c * 8 + 2 * i + 1]) * inv);
          *reinterpret_cast<uint4*>(op + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vb

extern "C" int b200v_attention_spatial_v5(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v,
                                          int64_t ld_v, void* out, int64_t ld_o, int32_t frames, int32_t seq,
                                          int32_t heads, void* stream_) {
  using namespace vb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VB_REQUIRE(q && k && v && out, "b200v_attention_spatial_v5: null pointer");
  VB_REQUIRE(frames > 0 && seq > 0 && heads > 0, "b200v_attention_spatial_v5: bad sizes");
  VB_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_o % 8 == 0,
             "b200v_attention_spatial_v5: row strides must be multiples of 8 elements");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b200v_attention_spatial_v5: unaligned output");
  CUtensorMap tm[3];
  const void* ptrs[3] = {q, k, v};
  const int64_t lds[3] = {ld_q, ld_k, ld_v};
  for (int i = 0; i < 3; ++i) {
    VB_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "b200v_attention_spatial_v5: unaligned pointer");
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)seq, (uint64_t)frames};
    uint64_t strides[2] = {(uint64_t)lds[i] * 2, (uint64_t)lds[i] * 2 * seq};
    uint32_t box[3] = {64, 128, 1};
    uint32_t es[3] = {1, 1, 1};
    if (encode_tmap_16bit(&tm[i], ptrs[i], 3, dims, strides, box, es, 0)) return 3;
  }
  Attn5Params p;
  p.seq = seq;
  p.n_kv = (seq + kT5 - 1) / kT5;
  p.n_qb = (seq + 2 * kT5 - 1) / (2 * kT5);
  p.heads = heads;
  const long long items = (long long)frames * heads * p.n_qb;
  VB_REQUIRE(items < (1ll << 31), "b200v_attention_spatial_v5: too many work items");
  p.n_items = (int)items;
  p.ld_o = ld_o;
  p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static int chunked = -1;
  if (chunked < 0) chunked = getenv("VB_ATTN5_CHUNKED") ? atoi(getenv("VB_ATTN5_CHUNKED")) : 1;
  p.chunked = chunked;
  static int pingpong = -1;
  if (pingpong < 0) pingpong = getenv("VB_ATTN5_PINGPONG") ? atoi(getenv("VB_ATTN5_PINGPONG")) : 1;
  p.pingpong = pingpong;
  const int smem_bytes = 1024 + 1024 + (4 + 2 * kNS5 + 1) * kT5Bytes;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Attn5Params);
  static const Kern kerns[5] = {attn5_spatial_kernel<0>, attn5_spatial_kernel<1>, attn5_spatial_kernel<2>,
                                attn5_spatial_kernel<3>, attn5_spatial_kernel<4>};
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
    mode = VB_ATTN5_EXP_DEFAULT;
    if (const char* e = getenv("VB_ATTN5_EXP")) mode = atoi(e);
    if (mode < 0 || mode > 4) mode = VB_ATTN5_EXP_DEFAULT;
  }
  int grid = device_sm_count();
  if (items < grid) grid = (int)items;
  kerns[mode]<<<grid, 352, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
