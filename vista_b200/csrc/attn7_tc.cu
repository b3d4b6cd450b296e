// Spatial self-attention v7 (head dim 64, non-causal): v5's persistent two-tile kernel with the S / P aliasing removed so
// that the tensor pipe never waits on a tile's own softmax.
//   warps 0-3 / 4-7 : softmax warpgroup of tile A / B (thread = query row = TMEM lane)
//   warp 8          : TMA producer (Q double-buffered per item, K / V rings of 128-key blocks)
//   warps 9, 10     : one tcgen05.mma issuing thread per tile (warp 9 owns the TMEM allocation)
// What limited v5 (profiles/r02_ncu_attn5.md): at head dim 64 every tcgen05.mma sits on the ~96-cycle per-instruction
// floor (profiles/r02_mma_probe.md), so the 12 MMAs of a (tile, 128-key block) occupy the pipe for ~1200 cycles — more
// than the 1024 cycles its 16 384 exponentials take on the MUFU.  The pipe is the bound, and v5 kept it from running flat
// out: P was written over the consumed scores, so S(j+1) of a tile could only be issued behind PV(j), i.e. after the
// tile's whole softmax, and landed in the OTHER tile's exponential window together with that tile's PV MMAs (demand ~1400
// cycles per 1024-cycle window).  Here S, P and O have their own tensor-memory columns (2 x (128 + 64 + 64) = 512):
//   * S(j+1) is issued as soon as the softmax threads have READ S(j) (s_free barrier, before their exponentials), so it
//     overlaps the tile's own softmax and is long finished when the tile asks for it;
//   * PV(j) follows chunk by chunk as P is published; a pv_done barrier guards P / O against the next block's writes;
//   * the row sum is accumulated by the row's thread (the ones-column of v5 would need 16 more columns), PV has N = 64.
// Per block and tile the pipe still sees 4 + 8 MMAs; what changes is that it always has independent work queued.
#include <stdlib.h>

#include "../../include/vista_b200.h"
#include "host.cuh"
#include "ptx.cuh"

#ifndef VB_ATTN7_EXP_DEFAULT
#define VB_ATTN7_EXP_DEFAULT 0
#endif

namespace vb {

constexpr int kT7 = 128;
constexpr int kT7Bytes = 128 * 128;   // 16 KB: 128 rows x 64 fp16
constexpr int kNS7 = 3;               // K / V ring depth

struct Attn7Params {
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

__device__ __forceinline__ void tmem7_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem7_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3f7(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
// exp2 on the FMA / ALU pipes: x = n + f, f in [-0.5, 0.5] through the magic-number round, degree-3 minimax polynomial
// for 2^f (max relative error 7.7e-5, below the fp16 rounding of P), exponent patched in with one integer multiply-add.
__device__ __forceinline__ float exp2_poly7(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __uint_as_float(__float_as_uint(t) * 8388608u + __float_as_uint(p));
}
__device__ __forceinline__ uint32_t pack7_h2(float lo, float hi) {
  uint32_t p;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
  return p;
}
__device__ __forceinline__ uint32_t ex2_h27(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int EXP>
__global__ void __launch_bounds__(352, 1)
attn7_spatial_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Attn7Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);   // [2] Q buffers
  uint64_t* q_empty = q_full + 2;                         // [2]
  uint64_t* k_full = q_empty + 2;                         // [kNS7]
  uint64_t* k_empty = k_full + kNS7;
  uint64_t* v_full = k_empty + kNS7;
  uint64_t* v_empty = v_full + kNS7;
  uint64_t* s_full = v_empty + kNS7;                      // [2] per tile
  uint64_t* p_full = s_full + 2;                          // [2 tiles][4 chunks of 32 keys]
  uint64_t* o_full = p_full + 8;                          // [2]
  uint64_t* s_free = o_full + 2;                          // [2] the tile's softmax threads have read S(j): S(j+1) may overwrite it
  uint64_t* pv_done = s_free + 2;                         // [2] PV(j) has completed: P and O may be written again
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  uint8_t* sQ = smem + 1024;                   // 2 buffers x 2 tiles
  uint8_t* sK = sQ + 4 * kT7Bytes;             // kNS7 stages
  uint8_t* sV = sK + kNS7 * kT7Bytes;          // kNS7 stages

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = p.n_kv;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);          // one tcgen05.commit per tile issuer
      mbar_init(&s_full[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&p_full[i], 128);
    for (int i = 0; i < kNS7; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    fence_barrier_init();
  }
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
  // TMEM columns: S_A [0,128)  S_B [128,256)  P_A [256,320)  P_B [320,384)  O_A [384,448)  O_B [448,512)

  // work item -> (frame, head, query block); consecutive items share (frame, head): the SMs that run them at the same
  // time read the same K / V from L2
  auto decode = [&](int item, int& frame, int& head, int& q0) {
    const int qb = item % p.n_qb;
    const int fh = item / p.n_qb;
    head = fh % p.heads;
    frame = fh / p.heads;
    q0 = qb * 2 * kT7;
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
        mbar_expect_tx(&q_full[qb], 2 * kT7Bytes);
        tma_load_3d(sQ + (2 * qb) * kT7Bytes, &tmQ, &q_full[qb], head * 64, q0, frame);
        tma_load_3d(sQ + (2 * qb + 1) * kT7Bytes, &tmQ, &q_full[qb], head * 64, q0 + kT7, frame);
        for (int j = 0; j < n_kv; ++j, ++kv) {
          const int st = kv % kNS7;
          const uint32_t ph = (kv / kNS7) & 1;
          mbar_wait_relaxed(&k_empty[st], ph ^ 1, 52);
          mbar_expect_tx(&k_full[st], kT7Bytes);
          tma_load_3d(sK + st * kT7Bytes, &tmK, &k_full[st], head * 64, j * kT7, frame);
          mbar_wait_relaxed(&v_empty[st], ph ^ 1, 53);
          mbar_expect_tx(&v_full[st], kT7Bytes);
          tma_load_3d(sV + st * kT7Bytes, &tmV, &v_full[st], head * 64, j * kT7, frame);
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------------------------------------------------ MMA issuers: one thread per tile
    // Per block j of a tile:  S(j+1) FIRST — it only needs the tile's softmax threads to have read S(j) (s_free) and the
    // next K block — then PV(j) in 32-key chunks as P is published.  S, P and O do not alias, so the only orderings left
    // are data dependencies; the K / V / Q "empty" barriers count one tcgen05.commit per tile (a skipped tile B makes its
    // arrivals without work, in step with the rings).
    if (lane == 0) {
      const int t = warp - 9;
      const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, 64, 0, 0, 1);   // B = V as MN-major atom
      const uint32_t tS = tmem_base + t * 128, tP = tmem_base + 256 + t * 64, tO = tmem_base + 384 + t * 64;
      uint32_t kq = 0, kv = 0, g = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++kq) {
        int frame, head, q0;
        decode(item, frame, head, q0);
        const bool on = t == 0 || (q0 + kT7 < p.seq);    // tile B entirely beyond the sequence: no work
        const int qb = kq & 1;
        auto issue_s = [&](uint32_t kvi) {
          const uint32_t q_base = smem_u32(sQ + (2 * qb + t) * kT7Bytes);
          const uint32_t k_base = smem_u32(sK + (kvi % kNS7) * kT7Bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tS, make_desc_sw128(q_base + k * 32, 16, 1024), make_desc_sw128(k_base + k * 32, 16, 1024), idesc_s,
                     k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
        };
        mbar_wait(&q_full[qb], (kq >> 1) & 1, 54);
        mbar_wait(&k_full[kv % kNS7], (kv / kNS7) & 1, 55);
        // S(0) of this item overwrites the S columns the previous item's last block used: its readers are done (s_free)
        if (on && g > 0) mbar_wait(&s_free[t], (g - 1) & 1, 61);
        tc_fence_after();
        if (on) issue_s(kv);
        umma_commit(&k_empty[kv % kNS7]);
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t cur = kv + j;
          const int st = cur % kNS7;
          if (j + 1 < n_kv) {            // S(j+1): ahead of PV(j)
            const int sn = (cur + 1) % kNS7;
            mbar_wait(&k_full[sn], ((cur + 1) / kNS7) & 1, 58);
            if (on) mbar_wait(&s_free[t], g & 1, 62);
            tc_fence_after();
            if (on) issue_s(cur + 1);
            umma_commit(&k_empty[sn]);
          }
          mbar_wait(&v_full[st], (cur / kNS7) & 1, 57);
          if (on) {
            const uint32_t v_base = smem_u32(sV + st * kT7Bytes);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              mbar_wait(&p_full[t * 4 + c], g & 1, 56);
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {           // keys 16 k .. 16 k + 15: 8 packed columns of P
                const int k = 2 * c + kk;
                umma_f16_ts(tO, tP + k * 8, make_desc_sw128(v_base + k * 2048, 16, 1024), idesc_o, (j | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&pv_done[t]);
            ++g;
          }
          umma_commit(&v_empty[st]);
          if (j + 1 == n_kv && on) umma_commit(&o_full[t]);
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
    const uint32_t tP = tmem_base + 256 + t * 64 + lane_off;
    const uint32_t tO = tmem_base + 384 + t * 64 + lane_off;
    uint32_t g = 0, items_done = 0;
    if (p.pingpong && t == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");     // tile A goes first
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int frame, head, q0;
      decode(item, frame, head, q0);
      if (t == 1 && q0 + kT7 >= p.seq) {          // tile B has no rows: it still passes the turn back, block by block
        if (p.pingpong) {
          for (int j = 0; j < n_kv; ++j) {
            asm volatile("bar.sync 2, 256;" ::: "memory");
            asm volatile("bar.arrive 1, 256;" ::: "memory");
          }
        }
        continue;
      }
      float m_used = -INFINITY;
      float l_sum = 0.f;                             // row sum of P (in units of 2^-m_used), this thread's row
      for (int j = 0; j < n_kv; ++j, ++g) {
        mbar_wait(&s_full[t], g & 1, 59);
        tc_fence_after();
        uint32_t s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(s + c * 32));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);            // the scores are in registers: S(j+1) may land on these columns
        const int kv_left = p.seq - j * kT7;
        if (kv_left < kT7) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_left) s[i] = 0xFF800000u;  // -inf
        }
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mxa[u] = max3f7(mxa[u], __uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1]));
        }
        const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
        const float m_blk = mx * p.scale_log2;
        // lazy rescale: only when the block maximum exceeds the maximum in use by more than 8 (factor 256)
        const bool need = m_blk > m_used + 8.0f;
        // P and O of this tile are free once PV(j-1) has completed (the S MMAs no longer imply it: they run ahead)
        if (g > 0) {
          mbar_wait(&pv_done[t], (g - 1) & 1, 63);
          tc_fence_after();
        }
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? m_blk : m_used;
          if (j > 0) {
            const float alpha = need ? ex2_f(m_used - m_new) : 1.0f;
            l_sum *= alpha;
#pragma unroll
            for (int c = 0; c < 4; ++c) {   // 64 accumulator columns
              uint32_t ov[16];
              tmem_ld16(tO + c * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem7_st16(tO + c * 16, ov);
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
        // P = exp2(s * scale - m_used) -> packed fp16 -> the tile's 64 P columns
#pragma unroll
        for (int c = 0; c < 4; ++c) {      // 32 keys -> 16 packed columns
          uint32_t pw[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int e = c * 32 + 2 * i;
            const float x0 = fmaf(__uint_as_float(s[e]), p.scale_log2, -m_used);
            const float x1 = fmaf(__uint_as_float(s[e + 1]), p.scale_log2, -m_used);
            if (EXP == 1) {
              pw[i] = ex2_h27(pack7_h2(x0, x1));
              const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&pw[i]));
              l_sum += pf.x + pf.y;
            } else {
              const float p0 = ((e & 7) < EXP) ? exp2_poly7(x0) : ex2_f(x0);
              const float p1 = (((e + 1) & 7) < EXP) ? exp2_poly7(x1) : ex2_f(x1);
              pw[i] = pack7_h2(p0, p1);
              l_sum += p0 + p1;
            }
          }
          // the store of chunk c - 1 (and the O rescale stores above) completed while chunk c was being computed: wait for it
          // BEFORE the next store is issued (tcgen05.wait::st covers every earlier store), then publish chunk c - 1
          if (c > 0 && p.chunked) {
            tmem7_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[t * 4 + c - 1]);
          }
          tmem7_st16(tP + c * 16, pw);
        }
        if (p.pingpong) {                  // the other tile's turn
          if (t == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
          else asm volatile("bar.arrive 1, 256;" ::: "memory");
        }
        tmem7_st_wait();
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
      uint32_t ov[64];
      tmem_ld32(tO, *reinterpret_cast<uint32_t(*)[32]>(ov));
      tmem_ld32(tO + 32, *reinterpret_cast<uint32_t(*)[32]>(ov + 32));
      tmem_ld_wait();
      const float inv = 1.0f / l_sum;
      const int q = q0 + t * kT7 + r;
      if (q < p.seq) {
        uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + ((long long)frame * p.seq + q) * p.ld_o + head * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack7_h2(__uint_as_float(ov[c * 8 + 2 * i]) * inv, __uint_as_float(ov[c * 8 + 2 * i + 1]) * inv);
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

extern "C" int b200v_attention_spatial_v7(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v,
                                          int64_t ld_v, void* out, int64_t ld_o, int32_t frames, int32_t seq,
                                          int32_t heads, void* stream_) {
  using namespace vb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VB_REQUIRE(q && k && v && out, "b200v_attention_spatial_v7: null pointer");
  VB_REQUIRE(frames > 0 && seq > 0 && heads > 0, "b200v_attention_spatial_v7: bad sizes");
  VB_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_o % 8 == 0,
             "b200v_attention_spatial_v7: row strides must be multiples of 8 elements");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b200v_attention_spatial_v7: unaligned output");
  CUtensorMap tm[3];
  const void* ptrs[3] = {q, k, v};
  const int64_t lds[3] = {ld_q, ld_k, ld_v};
  for (int i = 0; i < 3; ++i) {
    VB_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "b200v_attention_spatial_v7: unaligned pointer");
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)seq, (uint64_t)frames};
    uint64_t strides[2] = {(uint64_t)lds[i] * 2, (uint64_t)lds[i] * 2 * seq};
    uint32_t box[3] = {64, 128, 1};
    uint32_t es[3] = {1, 1, 1};
    if (encode_tmap_16bit(&tm[i], ptrs[i], 3, dims, strides, box, es, 0)) return 3;
  }
  Attn7Params p;
  p.seq = seq;
  p.n_kv = (seq + kT7 - 1) / kT7;
  p.n_qb = (seq + 2 * kT7 - 1) / (2 * kT7);
  p.heads = heads;
  const long long items = (long long)frames * heads * p.n_qb;
  VB_REQUIRE(items < (1ll << 31), "b200v_attention_spatial_v7: too many work items");
  p.n_items = (int)items;
  p.ld_o = ld_o;
  p.out = out;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static int chunked = -1;
  if (chunked < 0) chunked = getenv("VB_ATTN7_CHUNKED") ? atoi(getenv("VB_ATTN7_CHUNKED")) : 1;
  p.chunked = chunked;
  static int pingpong = -1;
  if (pingpong < 0) pingpong = getenv("VB_ATTN7_PINGPONG") ? atoi(getenv("VB_ATTN7_PINGPONG")) : 1;
  p.pingpong = pingpong;
  const int smem_bytes = 1024 + 1024 + (4 + 2 * kNS7) * kT7Bytes;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Attn7Params);
  static const Kern kerns[5] = {attn7_spatial_kernel<0>, attn7_spatial_kernel<1>, attn7_spatial_kernel<2>,
                                attn7_spatial_kernel<3>, attn7_spatial_kernel<4>};
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
    mode = VB_ATTN7_EXP_DEFAULT;
    if (const char* e = getenv("VB_ATTN7_EXP")) mode = atoi(e);
    if (mode < 0 || mode > 4) mode = VB_ATTN7_EXP_DEFAULT;
  }
  int grid = device_sm_count();
  if (items < grid) grid = (int)items;
  kerns[mode]<<<grid, 352, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
