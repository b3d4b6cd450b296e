// Tap-GEMM: persistent, warp-specialised tcgen05 kernel.
//   warp 8 : TMA producer  (A tile 128 tokens x 64 ch per tap / K-chunk, B tile tile_n x 64)
//   warp 9 : TMEM owner + single-thread tcgen05.mma issuer (M=128, N=tile_n, K=16 per instruction)
//   warps 0-7 : epilogue, two per TMEM lane quadrant (tcgen05.ld -> fused bias / row-vector / activation /
//               GEGLU -> smem transpose -> coalesced residual loads and stores)
// Two accumulator stages in TMEM (2 x 256 columns) let the epilogue of tile i overlap the MMAs of
// tile i+1.  The 3x3 / (3,1,1) convolutions are implicit GEMMs: the A tile of every tap is a
// shifted 4-D TMA box of the token-major activation, zero padding comes from TMA OOB fill.
#include "../../include/vista_b200.h"
#include "host.cuh"
#include "ptx.cuh"

namespace vb {

constexpr int kMaxStages = 8;
constexpr int kABytes = 128 * 64 * 2;  // 16 KB

struct TGParams {
  int a_mode;
  int W, H, NB;
  int BW, BH, BB;
  int tiles_w, tiles_h;
  int m_tiles, n_tiles;
  int ntaps, kc_per_tap;
  int dh[9], dw[9];
  int N, TN;
  int bf16;
  int nstages, stage_bytes;
  long long tokens;
  void* out;
  long long ldo;
  int out_f32, act;
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  int rv_div, rv_mod;
  const void* res1;
  long long ld_res1;
  float s_res1;
  const void* res2;
  long long ld_res2;
  float s_res2;
  float s_acc;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8], int bf16) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bf16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int bf16) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bf16) {
      __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float2 unpack2(uint32_t w, int bf16) {
  if (bf16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
__device__ __forceinline__ uint32_t pack2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
  }
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Column-group schedule of the two epilogue warpgroups.  G 32-column groups per tile; adjacent groups (2i, 2i+1)
// form a 128-byte output line and go to the same warpgroup back to back, pairs alternate between the
// warpgroups; with an odd number of pairs the last pair is split so that both warpgroups get the same load.
__device__ __forceinline__ int epi_group(int G, int wg, int k) {
  const int P = G >> 1;             // full pairs
  const int Pe = P & ~1;            // pairs that are dealt out two by two
  const int pairs_mine = Pe >> 1;   // per warpgroup
  if (k < 2 * pairs_mine) return ((k >> 1) * 2 + wg) * 2 + (k & 1);
  int kk = k - 2 * pairs_mine;
  if (P & 1) {                      // split the last pair
    if (kk == 0) return 2 * Pe + wg;
    --kk;
  }
  if ((G & 1) && kk == 0 && wg == ((P & 1) ? 1 : 0)) return G - 1;   // leftover single group
  return -1;
}

// Phase B of the epilogue: the warp walks its 32 staged rows (32 fp32 columns = 8 chunks of 16 bytes per row,
// 4 rows per instruction) so that every residual load / output store covers a contiguous 64-byte (fp16) or
// 128-byte (fp32) row segment.  The first residual was prefetched into registers before the accumulator was
// ready (u1); a second residual (AlphaBlender GEMMs only) is loaded here, batched ahead of its use.
__device__ __forceinline__ void epilogue_phase_b(const TGParams& p, uint32_t stg, int lane, const int (&tok)[8],
                                                 const bool (&okr)[8], const uint2 (&u1)[8], int n, bool n_ok) {
  const int ch = lane & 7;
  const int rsub = lane >> 3;
  const uint16_t* r2p = reinterpret_cast<const uint16_t*>(p.res2);
  float4 v[8];
  uint2 u2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + rsub;
    const int slot = (ch ^ row) & 7;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                 : "r"(stg + row * 128 + slot * 16));
    u2[i] = make_uint2(0, 0);
    if (r2p && okr[i] && n_ok) u2[i] = __ldg(reinterpret_cast<const uint2*>(r2p + (long long)tok[i] * p.ld_res2 + n));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (!(okr[i] && n_ok)) continue;
    float4 o = v[i];
    if (p.res1) {
      const float2 a = unpack2(u1[i].x, p.bf16), b = unpack2(u1[i].y, p.bf16);
      o.x += p.s_res1 * a.x; o.y += p.s_res1 * a.y; o.z += p.s_res1 * b.x; o.w += p.s_res1 * b.y;
    }
    if (r2p) {
      const float2 a = unpack2(u2[i].x, p.bf16), b = unpack2(u2[i].y, p.bf16);
      o.x += p.s_res2 * a.x; o.y += p.s_res2 * a.y; o.z += p.s_res2 * b.x; o.w += p.s_res2 * b.y;
    }
    if (p.out_f32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)tok[i] * p.ldo + n) = o;
    } else {
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + (long long)tok[i] * p.ldo + n) =
          make_uint2(pack2(o.x, o.y, p.bf16), pack2(o.z, o.w, p.bf16));
    }
  }
}

// 10 warps are allocated as 12 (granularity 4): 65536 / 384 -> at most 168 registers per thread
__global__ void __launch_bounds__(320, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TGParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // 1024 B aligned (SWIZZLE_128B requirement)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);           // [kMaxStages]
  uint64_t* empty = full + kMaxStages;                          // [kMaxStages]
  uint64_t* tfull = empty + kMaxStages;                         // [2]
  uint64_t* tempty = tfull + 2;                                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stages = smem + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int KC = p.ntaps * p.kc_per_tap;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = kABytes + p.TN * 128;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
        const int tw = m_blk % p.tiles_w;
        const int th = (m_blk / p.tiles_w) % p.tiles_h;
        const int tb = m_blk / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH, b0 = tb * p.BB;
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          uint8_t* sA = stages + stage * p.stage_bytes;
          uint8_t* sB = sA + kABytes;
          mbar_expect_tx(&full[stage], tx);
          const int tap = kc / p.kc_per_tap;
          const int c0 = (kc - tap * p.kc_per_tap) * 64;
          if (p.a_mode == 0)
            tma_load_2d(sA, &tmA, &full[stage], c0, w0);
          else
            tma_load_4d(sA, &tmA, &full[stage], c0, w0 + p.dw[tap], h0 + p.dh[tap], b0);
          tma_load_2d(sB, &tmB, &full[stage], kc * 64, n_blk * p.TN);
          if (++stage == p.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t idesc = make_idesc_f16(128, p.TN, p.bf16, 0, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          const uint32_t a_base = smem_u32(stages + stage * p.stage_bytes);
          const uint32_t b_base = a_base + kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bd = make_desc_sw128(b_base + k * 32, 16, 1024);
            umma_f16(d_tmem, ad, bd, idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == p.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[as]);  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0..7)
    // Warp w owns TMEM lane quadrant w % 4 (rows 32*(w%4) ..) and every second 32-column group (w / 4).
    // Phase A: thread = tile row: accumulator -> (+bias)*s_acc + rowvec -> act / GEGLU, fp32, written to the
    //          warp's staging buffer (32 rows x 32 cols, 16-byte chunks XOR-swizzled by row).
    // Phase B: coalesced residual loads / output stores (epilogue_phase_b).
    int as = 0;
    uint32_t aphase = 0;
    const int quad = warp & 3, wg = warp >> 2;
    const int r = quad * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t stg = smem_u32(stages + p.nstages * p.stage_bytes) + warp * (32 * 128);
    const int half = p.TN >> 1;
    const int tile_out_cols = (p.act == 2) ? half : p.TN;
    const int n_out_total = (p.act == 2) ? (p.N >> 1) : p.N;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
      int token_own, valid_own;
      if (p.a_mode == 0) {
        token_own = m_blk * 128 + r;
        valid_own = token_own < p.tokens;
      } else {
        const int tw = m_blk % p.tiles_w;
        const int th = (m_blk / p.tiles_w) % p.tiles_h;
        const int tb = m_blk / (p.tiles_w * p.tiles_h);
        const int ww = r % p.BW, hh = (r / p.BW) % p.BH, bb = r / (p.BW * p.BH);
        const int w = tw * p.BW + ww, h = th * p.BH + hh, b = tb * p.BB + bb;
        valid_own = (w < p.W) && (h < p.H) && (b < p.NB);
        token_own = (b * p.H + h) * p.W + w;
      }
      const float* rv_ptr = p.rowvec ? p.rowvec + (long long)((token_own / p.rv_div) % p.rv_mod) * p.ld_rowvec : nullptr;
      const int n_out_base = n_blk * tile_out_cols;
      // rows this lane serves in phase B (4 rows per instruction, 8 instructions) and their tokens
      int tok[8];
      bool okr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3);
        tok[i] = __shfl_sync(0xffffffffu, token_own, row);
        okr[i] = __shfl_sync(0xffffffffu, valid_own, row) != 0;
      }
      // The first residual is prefetched into registers one column group ahead (group 0 while the MMAs of the
      // tile still run).  Group schedule: see epi_group().
      uint2 rpre[2][8];
      const uint16_t* r1p = reinterpret_cast<const uint16_t*>(p.res1);
      const int G = tile_out_cols >> 5;
      auto prefetch_res = [&](int k, uint2 (&dst)[8]) {
        const int g = epi_group(G, wg, k);
        const int c0 = g * 32;
        const int n = n_out_base + c0 + (lane & 7) * 4;
        const bool n_ok = (g >= 0) && (n + 4 <= n_out_total);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dst[i] = make_uint2(0, 0);
          if (r1p && n_ok && okr[i]) dst[i] = __ldg(reinterpret_cast<const uint2*>(r1p + (long long)tok[i] * p.ld_res1 + n));
        }
      };
      prefetch_res(0, rpre[0]);
      mbar_wait(&tfull[as], aphase, 4);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + as * 256;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int gidx = epi_group(G, wg, k);
        if (gidx < 0) break;
        const int c0 = gidx * 32;
        if (k + 1 < 4) prefetch_res(k + 1, rpre[(k + 1) & 1]);
        // ---------------- phase A
        float f[32];
        if (p.act != 2) {
          uint32_t v[32];
          tmem_ld32(t_row + c0, v);
          tmem_ld_wait();
          const int n0 = n_out_base + c0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), rv = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool in = n0 + g * 4 + 4 <= p.N;
            if (p.bias && in) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + g * 4));
            if (rv_ptr && in) rv = __ldg(reinterpret_cast<const float4*>(rv_ptr + n0 + g * 4));
            f[g * 4 + 0] = (__uint_as_float(v[g * 4 + 0]) + bv.x) * p.s_acc + rv.x;
            f[g * 4 + 1] = (__uint_as_float(v[g * 4 + 1]) + bv.y) * p.s_acc + rv.y;
            f[g * 4 + 2] = (__uint_as_float(v[g * 4 + 2]) + bv.z) * p.s_acc + rv.z;
            f[g * 4 + 3] = (__uint_as_float(v[g * 4 + 3]) + bv.w) * p.s_acc + rv.w;
          }
          if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = silu_f(f[i]);
          }
        } else {
          uint32_t va[32], vg[32];
          tmem_ld32(t_row + c0, va);
          tmem_ld32(t_row + half + c0, vg);
          tmem_ld_wait();
          const int nb = n_blk * p.TN + c0;  // bias index of the value columns (gate: + half)
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) {
              ba = __ldg(reinterpret_cast<const float4*>(p.bias + nb + g * 4));
              bg = __ldg(reinterpret_cast<const float4*>(p.bias + nb + half + g * 4));
            }
            f[g * 4 + 0] = (__uint_as_float(va[g * 4 + 0]) + ba.x) * gelu_erf_fast(__uint_as_float(vg[g * 4 + 0]) + bg.x);
            f[g * 4 + 1] = (__uint_as_float(va[g * 4 + 1]) + ba.y) * gelu_erf_fast(__uint_as_float(vg[g * 4 + 1]) + bg.y);
            f[g * 4 + 2] = (__uint_as_float(va[g * 4 + 2]) + ba.z) * gelu_erf_fast(__uint_as_float(vg[g * 4 + 2]) + bg.z);
            f[g * 4 + 3] = (__uint_as_float(va[g * 4 + 3]) + ba.w) * gelu_erf_fast(__uint_as_float(vg[g * 4 + 3]) + bg.w);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int slot = (j ^ lane) & 7;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + slot * 16), "f"(f[j * 4]),
                       "f"(f[j * 4 + 1]), "f"(f[j * 4 + 2]), "f"(f[j * 4 + 3])
                       : "memory");
        }
        __syncwarp();
        // ---------------- phase B
        {
          const int n = n_out_base + c0 + (lane & 7) * 4;
          epilogue_phase_b(p, stg, lane, tok, okr, rpre[k & 1], n, n + 4 <= n_out_total);
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
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

extern "C" int b200v_gemm(const b200v_gemm_desc* d, void* stream_) {
  using namespace vb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VB_REQUIRE(d && d->a && d->b && d->out, "b200v_gemm: null pointer");
  VB_REQUIRE(d->cin > 0 && d->cin % 64 == 0, "b200v_gemm: cin=%d must be a positive multiple of 64", d->cin);
  VB_REQUIRE(d->ntaps >= 1 && d->ntaps <= 9, "b200v_gemm: ntaps=%d out of range", d->ntaps);
  VB_REQUIRE(d->N > 0 && d->N % 8 == 0, "b200v_gemm: N=%d must be a multiple of 8", d->N);
  VB_REQUIRE(d->tile_n >= 32 && d->tile_n <= 256 && d->tile_n % 32 == 0, "b200v_gemm: tile_n=%d invalid", d->tile_n);
  VB_REQUIRE(d->lda % 8 == 0 && d->ldo % 8 == 0, "b200v_gemm: lda/ldo must be multiples of 8");
  VB_REQUIRE(d->act >= 0 && d->act <= 2, "b200v_gemm: act=%d invalid", d->act);
  VB_REQUIRE(!(d->act == 2 && (d->tile_n % 64 != 0 || d->N % d->tile_n != 0 || d->out_f32)),
             "b200v_gemm: GEGLU needs tile_n %% 64 == 0, N %% tile_n == 0, 16-bit output");
  VB_REQUIRE(!(d->act == 2 && (d->rowvec || d->res1 || d->res2)), "b200v_gemm: GEGLU epilogue takes bias only");
  VB_REQUIRE(!d->res1 || d->ld_res1 % 8 == 0, "b200v_gemm: ld_res1 must be a multiple of 8");
  VB_REQUIRE(!d->res2 || d->ld_res2 % 8 == 0, "b200v_gemm: ld_res2 must be a multiple of 8");
  VB_REQUIRE(!d->rowvec || (d->ld_rowvec % 4 == 0 && d->rv_div > 0 && d->rv_mod > 0), "b200v_gemm: bad rowvec args");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->b) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
             "b200v_gemm: a/b/out must be 16-byte aligned");

  TGParams p;
  memset(&p, 0, sizeof(p));
  p.a_mode = d->a_mode;
  p.tokens = d->tokens;
  const long long K = (long long)d->ntaps * d->cin;
  CUtensorMap tmA, tmB;
  if (d->a_mode == 0) {
    VB_REQUIRE(d->ntaps == 1, "b200v_gemm: linear mode takes one tap");
    VB_REQUIRE(d->tokens > 0 && d->tokens < (1ll << 31) - 256, "b200v_gemm: tokens out of range");
    p.W = (int)d->tokens; p.H = 1; p.NB = 1;
    p.BW = 128; p.BH = 1; p.BB = 1;
    p.tiles_w = (int)((d->tokens + 127) / 128); p.tiles_h = 1;
    p.m_tiles = p.tiles_w;
    uint64_t dims[2] = {(uint64_t)d->cin, (uint64_t)d->tokens};
    uint64_t strides[1] = {(uint64_t)d->lda * 2};
    uint32_t box[2] = {64, 128};
    uint32_t es[2] = {1, 1};
    if (encode_tmap_16bit(&tmA, d->a, 2, dims, strides, box, es, d->bf16)) return 3;
  } else {
    VB_REQUIRE(d->W > 0 && d->H > 0 && d->NB > 0 && (long long)d->W * d->H * d->NB == d->tokens &&
                   d->tokens < (1ll << 31) - 256,
               "b200v_gemm: W*H*NB != tokens (or >= 2^31)");
    VB_REQUIRE(d->box_w > 0 && d->box_h > 0 && d->box_b > 0 && d->box_w * d->box_h * d->box_b == 128,
               "b200v_gemm: box_w*box_h*box_b must be 128");
    p.W = d->W; p.H = d->H; p.NB = d->NB;
    p.BW = d->box_w; p.BH = d->box_h; p.BB = d->box_b;
    p.tiles_w = (d->W + p.BW - 1) / p.BW;
    p.tiles_h = (d->H + p.BH - 1) / p.BH;
    const int tiles_b = (d->NB + p.BB - 1) / p.BB;
    p.m_tiles = p.tiles_w * p.tiles_h * tiles_b;
    uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->NB};
    uint64_t strides[3] = {(uint64_t)d->lda * 2, (uint64_t)d->lda * 2 * d->W, (uint64_t)d->lda * 2 * d->W * d->H};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BB};
    uint32_t es[4] = {1, 1, 1, 1};
    if (encode_tmap_16bit(&tmA, d->a, 4, dims, strides, box, es, d->bf16)) return 3;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)d->N};
    uint64_t strides[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, (uint32_t)d->tile_n};
    uint32_t es[2] = {1, 1};
    if (encode_tmap_16bit(&tmB, d->b, 2, dims, strides, box, es, d->bf16)) return 3;
  }
  p.ntaps = d->ntaps;
  p.kc_per_tap = d->cin / 64;
  for (int i = 0; i < 9; ++i) {
    p.dh[i] = d->dh[i];
    p.dw[i] = d->dw[i];
  }
  p.N = d->N;
  p.TN = d->tile_n;
  p.n_tiles = (d->N + d->tile_n - 1) / d->tile_n;
  p.bf16 = d->bf16;
  p.stage_bytes = kABytes + ((d->tile_n * 128 + 1023) / 1024) * 1024;
  constexpr int kStagingBytes = 8 * 32 * 128;  // epilogue transpose buffers (8 warps x 32 rows x 32 fp32)
  p.nstages = (227 * 1024 - 2048 - kStagingBytes) / p.stage_bytes;
  if (p.nstages > kMaxStages) p.nstages = kMaxStages;
  p.out = d->out; p.ldo = d->ldo; p.out_f32 = d->out_f32; p.act = d->act;
  p.bias = d->bias;
  p.rowvec = d->rowvec; p.ld_rowvec = d->ld_rowvec; p.rv_div = d->rv_div > 0 ? d->rv_div : 1;
  p.rv_mod = d->rv_mod > 0 ? d->rv_mod : 1;
  p.res1 = d->res1; p.ld_res1 = d->ld_res1; p.s_res1 = d->s_res1;
  p.res2 = d->res2; p.ld_res2 = d->ld_res2; p.s_res2 = d->s_res2;
  p.s_acc = d->s_acc;

  const int smem_bytes = 1024 + 1024 + p.nstages * p.stage_bytes + kStagingBytes;
  static bool attr_set = false;
  if (!attr_set) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long total = (long long)p.m_tiles * p.n_tiles;
  int grid = device_sm_count();
  if (total < grid) grid = (int)total;
  tapgemm_kernel<<<grid, 320, smem_bytes, stream>>>(tmA, tmB, p);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
