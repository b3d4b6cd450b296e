// Tap-GEMM: persistent, warp-specialised tcgen05 kernel.
//   warp 4 : TMA producer  (A tile 128 tokens x 64 ch per tap / K-chunk, B tile tile_n x 64)
//   warp 5 : TMEM owner + single-thread tcgen05.mma issuer (M=128, N=tile_n, K=16 per instruction)
//   warps 0-3 : epilogue (tcgen05.ld -> fused bias / row-vector / activation / residuals -> global)
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

__global__ void __launch_bounds__(192, 1)
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
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
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
  } else if (warp == 5) {
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
    // ------------------------------------------------------------ epilogue (warps 0..3)
    // Phase A: thread = tile row (TMEM lane): accumulator -> (+bias)*s_acc + rowvec -> act / GEGLU, fp32,
    //          written to a per-warp staging buffer (32 rows x 64 cols, 16-byte chunks XOR-swizzled by row).
    // Phase B: the warp re-reads the buffer row-segment-wise so that residual loads and output stores are
    //          128-byte contiguous per row (coalesced) instead of one 16-byte piece per row and instruction.
    int as = 0;
    uint32_t aphase = 0;
    const int r = warp * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t stg = smem_u32(stages + p.nstages * p.stage_bytes) + warp * (32 * 256);
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
      mbar_wait(&tfull[as], aphase, 4);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + as * 256;
      for (int c0 = 0; c0 < tile_out_cols; c0 += 64) {
        const int gw = min(64, tile_out_cols - c0);
        // ---------------- phase A
        for (int sub = 0; sub < gw; sub += 32) {
          float f[32];
          if (p.act != 2) {
            uint32_t v[32];
            tmem_ld32(t_row + c0 + sub, v);
            tmem_ld_wait();
            const int n0 = n_out_base + c0 + sub;
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
            tmem_ld32(t_row + c0 + sub, va);
            tmem_ld32(t_row + half + c0 + sub, vg);
            tmem_ld_wait();
            const int nb = n_blk * p.TN + c0 + sub;  // bias index of the value columns (gate: + half)
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
            const int ch = (sub >> 2) + j;
            const int slot = (ch & 8) | ((ch ^ lane) & 7);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 256 + slot * 16), "f"(f[j * 4]),
                         "f"(f[j * 4 + 1]), "f"(f[j * 4 + 2]), "f"(f[j * 4 + 3])
                         : "memory");
          }
        }
        __syncwarp();
        // ---------------- phase B
        const int cpr = gw >> 2;             // 16-byte fp32 chunks per row: 16 or 8
        const int rows_per_it = 32 / cpr;    // 2 or 4
        const int ch = lane % cpr;
        const int n = n_out_base + c0 + ch * 4;
        const bool n_ok = n + 4 <= n_out_total;
        for (int it = 0; it < 32; it += rows_per_it) {
          const int row = it + lane / cpr;
          const int slot = (ch & 8) | ((ch ^ row) & 7);
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(stg + row * 256 + slot * 16));
          const int tok = __shfl_sync(0xffffffffu, token_own, row);
          const int ok = __shfl_sync(0xffffffffu, valid_own, row);
          if (ok && n_ok) {
            if (p.res1) {
              const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.res1) +
                                                                   (long long)tok * p.ld_res1 + n));
              float2 a, b;
              if (p.bf16) {
                a = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u));
                b = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
              } else {
                a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
                b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
              }
              v.x += p.s_res1 * a.x; v.y += p.s_res1 * a.y; v.z += p.s_res1 * b.x; v.w += p.s_res1 * b.y;
            }
            if (p.res2) {
              const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.res2) +
                                                                   (long long)tok * p.ld_res2 + n));
              float2 a, b;
              if (p.bf16) {
                a = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u));
                b = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
              } else {
                a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
                b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
              }
              v.x += p.s_res2 * a.x; v.y += p.s_res2 * a.y; v.z += p.s_res2 * b.x; v.w += p.s_res2 * b.y;
            }
            if (p.out_f32) {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)tok * p.ldo + n) = v;
            } else {
              uint2 o;
              if (p.bf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                o.x = *reinterpret_cast<uint32_t*>(&lo);
                o.y = *reinterpret_cast<uint32_t*>(&hi);
              } else {
                __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
                o.x = *reinterpret_cast<uint32_t*>(&lo);
                o.y = *reinterpret_cast<uint32_t*>(&hi);
              }
              *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + (long long)tok * p.ldo + n) = o;
            }
          }
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
  if (warp == 5) {
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
  constexpr int kStagingBytes = 4 * 32 * 256;  // epilogue transpose buffers (4 warps x 32 rows x 64 fp32)
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
  tapgemm_kernel<<<grid, 192, smem_bytes, stream>>>(tmA, tmB, p);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
