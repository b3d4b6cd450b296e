// Tap-GEMM: persistent, warp-specialised tcgen05 kernel (template tapgemm_kernel<ACT, RV, NRES, GEN, PAIR, NQ>).
//   warps 0 .. 4 NQ - 1 : epilogue, NQ (2 or 4) per TMEM lane quadrant (tcgen05.ld -> smem transpose -> fused bias /
//                         row-vector / SiLU / GEGLU / residuals -> coalesced stores)
//   warp 4 NQ           : TMA producer  (A tile 128 tokens x 64 ch per tap / K-chunk, B tile tile_n x 64)
//   warp 4 NQ + 1       : TMEM owner + single-thread tcgen05.mma issuer (M=128, N=tile_n, K=16 per instruction;
//                         PAIR: cta_group::2, M=256 over the two CTAs of a cluster)
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
  int bw_sh, bh_sh;             // log2(BW), log2(BH): the box extents are powers of two (product 128)
  int tiles_w, tiles_h;
  int m_tiles, n_tiles;
  int m_pairs;                  // ceil(m_tiles / 2): m-tile pairs of the 2-CTA schedule
  int ntaps, kc_per_tap;
  int dh[9], dw[9];
  int N, TN;
  int bf16;
  int nstages, stage_bytes;
  long long tokens;
  void* out;
  int ldo_b;                    // row strides in BYTES (int: one IMAD.WIDE per address)
  int out_f32, act;
  const float* bias;
  const float* rowvec;
  int ld_rowvec_b;
  int rv_div, rv_mod;
  const void* res1;
  int ld_res1_b;
  float s_res1;
  const void* res2;
  int ld_res2_b;
  float s_res2;
  float s_acc;
  // GroupNorm statistics of the OUTPUT, fused (STATS variants): per (128-token tile, TMEM lane quadrant) column sums
  // and sums of squares of the stored values, fp32, at stats[((tile * 4 + quad) * stats_ld + stats_col0 + n) * 2 + {0,1}];
  // the host guarantees that a tile is 128 consecutive tokens (b200v_gemm checks the box)
  float* stats;
  int stats_ld, stats_col0;
};

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

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }
// XOR key of the staging swizzle: 16-byte chunk j of row r sits in slot (j ^ key(r)) mod CH; conflict-free both for
// the row-per-thread writes and for the RPI-rows-per-instruction reads (CH = 8: 128-byte rows, CH = 4: 64-byte rows).
template <int CH>
__device__ __forceinline__ int stage_swz(int row) { return CH == 8 ? row : (row >> 1); }

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

// NQ = 2: 10 warps are allocated as 12 (granularity 4): 65536 / 384 -> at most 168 registers per thread.
// NQ = 4: 18 warps are allocated as 20: the bound is declared as 640 threads so that the compiler stays <= 96.
template <int ACT_, bool RV_, int NRES_, bool GEN, bool PAIR, int NQ, bool STATS = false>
__global__ void __launch_bounds__(NQ == 2 ? 320 : 640, 1)
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
  const int KC = p.ntaps * p.kc_per_tap;
  // Tile schedule.  1-CTA: CTA b walks tiles b, b + grid, ...; tile -> (m_blk, n_blk) = (tile / n_tiles, tile % n_tiles).
  // CTA pair: cluster c walks pair-tiles; pair-tile -> m-tile pair (2 j, 2 j + 1) x n_blk, this CTA takes 2 j + rank
  // (an odd m-tile count leaves one all-out-of-range tile: TMA zero fill, no rows stored).
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int total_tiles = (PAIR ? p.m_pairs : p.m_tiles) * p.n_tiles;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], (PAIR ? 2 : 1) * 4 * NQ);   // one arrival per epilogue warp (of both CTAs in a pair)
    }
    fence_barrier_init();
  }
  if (warp == 4 * NQ && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (PAIR) {
    cluster_sync_all();                       // both CTAs' barriers exist before anything can signal them
    if (warp == 4 * NQ + 1) tmem_alloc_pair<512>(tmem_slot);
  } else {
    if (warp == 4 * NQ + 1) tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4 * NQ) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int b_rows = PAIR ? (p.TN >> 1) : p.TN;           // B rows this CTA stages
      const uint32_t tx = PAIR ? 2u * (kABytes + b_rows * 128) : (uint32_t)(kABytes + b_rows * 128);
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int mq = tile / p.n_tiles, n_blk = tile - mq * p.n_tiles;
        const int m_blk = PAIR ? 2 * mq + (int)cta_rank : mq;
        const int tw = m_blk % p.tiles_w;
        const int th = (m_blk / p.tiles_w) % p.tiles_h;
        const int tb = m_blk / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH, b0 = tb * p.BB;
        const int n0 = n_blk * p.TN + (PAIR ? (int)cta_rank * b_rows : 0);
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait_relaxed(&empty[stage], phase ^ 1, 1);
          uint8_t* sA = stages + stage * p.stage_bytes;
          uint8_t* sB = sA + kABytes;
          const int tap = kc / p.kc_per_tap;
          const int c0 = (kc - tap * p.kc_per_tap) * 64;
          if (PAIR) {
            if (cta_rank == 0) mbar_expect_tx(&full[stage], tx);   // bytes of both CTAs land on the leader's barrier
            if (p.a_mode == 0)
              tma_load_2d_pair(sA, &tmA, &full[stage], c0, w0);
            else
              tma_load_4d_pair(sA, &tmA, &full[stage], c0, w0 + p.dw[tap], h0 + p.dh[tap], b0);
            tma_load_2d_pair(sB, &tmB, &full[stage], kc * 64, n0);
          } else {
            mbar_expect_tx(&full[stage], tx);
            if (p.a_mode == 0)
              tma_load_2d(sA, &tmA, &full[stage], c0, w0);
            else
              tma_load_4d(sA, &tmA, &full[stage], c0, w0 + p.dw[tap], h0 + p.dh[tap], b0);
            tma_load_2d(sB, &tmB, &full[stage], kc * 64, n0);
          }
          if (++stage == p.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 4 * NQ + 1) {
    // ------------------------------------------------------------ MMA issuer (CTA pair: the leader only)
    if (lane == 0 && cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, p.TN, p.bf16, 0, 0);
      // Shared-memory descriptors (make_desc_sw128(addr, 16, 1024)): the high word is constant, the low word is
      // (addr >> 4) | LBO; it advances by stage_bytes >> 4 per pipeline stage and by 2 (32 bytes) per K = 16 step.
      // Kept incrementally so that the single issuing thread spends a handful of instructions per MMA.
      const uint64_t desc_hi = make_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
      const uint32_t a_lo0 = (uint32_t)(make_desc_sw128(smem_u32(stages), 16, 1024) & 0xFFFFFFFFull);
      const uint32_t lo_step = (uint32_t)p.stage_bytes >> 4;
      const uint32_t b_off = kABytes >> 4;
      uint32_t a_lo = a_lo0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        mbar_wait_relaxed(&tempty[as], aphase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = desc_hi | (uint64_t)(a_lo + 2 * k);
            const uint64_t bd = desc_hi | (uint64_t)(a_lo + b_off + 2 * k);
            if (PAIR)
              umma_f16_pair(d_tmem, ad, bd, idesc, (kc | k) != 0 ? 1u : 0u);
            else
              umma_f16(d_tmem, ad, bd, idesc, (kc | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (PAIR) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);
          a_lo += lo_step;
          if (++stage == p.nstages) {
            stage = 0;
            phase ^= 1;
            a_lo = a_lo0;
          }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (PAIR) umma_commit_pair(&tfull[as]); else umma_commit(&tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0 .. 4 NQ - 1)
    // Warp w owns TMEM lane quadrant w % 4 (rows 32*(w%4) ..) and column groups of CG = 64 / NQ columns:
    // NQ = 2 (8 warps): 32-column groups dealt out by epi_group(); NQ = 4 (16 warps): 16-column groups, round robin.
    // Phase A (thread = tile row): tcgen05.ld of the group; GEGLU is evaluated here (value * gelu(gate)), all other
    //          epilogues move the raw accumulator; CG fp32 per row go to the warp's staging buffer (16-byte
    //          chunks XOR-swizzled by row).
    // Phase B (lane = 4 fixed columns, RPI rows per instruction): bias / s_acc / row-vector / SiLU / residuals /
    //          pack / store; per-column operands are loaded once per group, every global access covers a
    //          contiguous row segment of CG output elements.  The first residual is prefetched one group ahead
    //          (group 0 while the MMAs of the tile still run).
    constexpr int CG = 64 / NQ;            // columns per group
    constexpr int CH = CG / 4;             // 16-byte fp32 chunks per staged row
    constexpr int RPI = 32 / CH;           // rows per phase-B instruction
    constexpr int NI = 32 / RPI;           // phase-B instructions per group
    constexpr int MAXK = 4;                // groups per warp and tile (256 / CG / NQ)
    const int act = GEN ? p.act : ACT_;
    const bool has_rv = GEN ? (p.rowvec != nullptr) : RV_;
    const bool has_r1 = GEN ? (p.res1 != nullptr) : (NRES_ >= 1);
    const bool has_r2 = GEN ? (p.res2 != nullptr) : (NRES_ >= 2);
    const bool bf16 = GEN ? (p.bf16 != 0) : false;
    const bool f32o = GEN ? (p.out_f32 != 0) : false;
    int as = 0;
    uint32_t aphase = 0;
    const int quad = warp & 3, wg = warp >> 2;
    const int r = quad * 32 + lane;  // row of the tile == TMEM lane
    const int ch = lane % CH, rsub = lane / CH;
    const uint32_t stg = smem_u32(stages + p.nstages * p.stage_bytes) + warp * (32 * CG * 4);
    const int half = p.TN >> 1;
    const int tile_out_cols = (act == 2) ? half : p.TN;
    const int n_out_total = (act == 2) ? (p.N >> 1) : p.N;
    const int G = tile_out_cols / CG;
    const int out_es = f32o ? 4 : 2;
    const char* r1p = reinterpret_cast<const char*>(p.res1);
    const char* r2p = reinterpret_cast<const char*>(p.res2);
    const bool has_bias = p.bias != nullptr;
    auto group_of = [&](int k) -> int {
      if (NQ == 2) return epi_group(G, wg, k);
      const int g = k * NQ + wg;
      return g < G ? g : -1;
    };
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int mq = tile / p.n_tiles, n_blk = tile - mq * p.n_tiles;
      const int m_blk = PAIR ? 2 * mq + (int)cta_rank : mq;
      int token_own, valid_own;
      if (p.a_mode == 0) {
        token_own = m_blk * 128 + r;
        valid_own = token_own < p.tokens;
      } else {
        const int tw = m_blk % p.tiles_w;
        const int t2 = m_blk / p.tiles_w;
        const int th = t2 % p.tiles_h;
        const int tb = t2 / p.tiles_h;
        const int ww = r & (p.BW - 1), hh = (r >> p.bw_sh) & (p.BH - 1), bb = r >> (p.bw_sh + p.bh_sh);
        const int w = tw * p.BW + ww, h = th * p.BH + hh, b = tb * p.BB + bb;
        valid_own = (w < p.W) && (h < p.H) && (b < p.NB);
        token_own = (b * p.H + h) * p.W + w;
      }
      if (!valid_own) token_own = 0;
      // rows this lane serves in phase B: row(i) = RPI i + rsub
      const uint32_t vrows = __ballot_sync(0xffffffffu, valid_own) >> rsub;   // bit RPI i <-> row(i)
      int tok[NI], rvrow[NI];
      {
        const int rv_own = has_rv ? (token_own / p.rv_div) % p.rv_mod : 0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          tok[i] = __shfl_sync(0xffffffffu, token_own, i * RPI + rsub);
          rvrow[i] = has_rv ? __shfl_sync(0xffffffffu, rv_own, i * RPI + rsub) : 0;
        }
      }
      const int n_out_base = n_blk * tile_out_cols;
      uint2 rpre[2][NI];
      auto prefetch_res = [&](int k, uint2 (&dst)[NI]) {
        if (!has_r1) return;
        const int g = group_of(k);
        const int n = n_out_base + g * CG + ch * 4;
        const bool ok = (g >= 0) && (n + 4 <= n_out_total);
        const char* base = r1p + (long long)n * 2;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          dst[i] = make_uint2(0, 0);
          if (ok && ((vrows >> (RPI * i)) & 1u)) dst[i] = __ldg(reinterpret_cast<const uint2*>(base + (long long)tok[i] * p.ld_res1_b));
        }
      };
      prefetch_res(0, rpre[0]);
      mbar_wait(&tfull[as], aphase, 4);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + as * 256;
#pragma unroll
      for (int k = 0; k < MAXK; ++k) {
        const int gidx = group_of(k);
        if (gidx < 0) break;
        const int c0 = gidx * CG;
        if (k + 1 < MAXK) prefetch_res(k + 1, rpre[(k + 1) & 1]);
        if (n_out_base + c0 >= n_out_total) continue;   // group entirely beyond N (last n-tile of a padded N)
        // ---------------- phase A
        if (act != 2) {
          uint32_t v[CG];
          tmem_ld_cols<CG>(t_row + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int slot = (j ^ stage_swz<CH>(lane)) & (CH - 1);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * (CG * 4) + slot * 16), "r"(v[j * 4]),
                         "r"(v[j * 4 + 1]), "r"(v[j * 4 + 2]), "r"(v[j * 4 + 3])
                         : "memory");
          }
        } else {
          // bias of the group's value | gate columns: one coalesced load per lane, broadcast through the (idle)
          // staging buffer instead of 2 CG / 4 uniform loads with their address arithmetic in every thread
          const int nb = n_blk * p.TN + c0;  // bias index of the value columns (gate: + half)
          if (has_bias) {
            float bv = 0.f;
            if (lane < 2 * CG || CG == 32) {
              if (CG == 32) {
                const float b0 = __ldg(p.bias + nb + lane), b1 = __ldg(p.bias + nb + half + lane);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(stg + lane * 4), "f"(b0) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(stg + (CG + lane) * 4), "f"(b1) : "memory");
              } else {
                bv = __ldg(p.bias + nb + (lane < CG ? lane : half + lane - CG));
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(stg + lane * 4), "f"(bv) : "memory");
              }
            }
            __syncwarp();
          }
          uint32_t va[CG], vg[CG];
          tmem_ld_cols<CG>(t_row + c0, va);
          tmem_ld_cols<CG>(t_row + half + c0, vg);
          float f[CG];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_bias) {
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ba.x), "=f"(ba.y), "=f"(ba.z), "=f"(ba.w) : "r"(stg + j * 16));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bg.x), "=f"(bg.y), "=f"(bg.z), "=f"(bg.w) : "r"(stg + (CG + j * 4) * 4));
            }
            if (j == 0) tmem_ld_wait();
            f[j * 4 + 0] = geglu_fast(__uint_as_float(va[j * 4 + 0]) + ba.x, __uint_as_float(vg[j * 4 + 0]) + bg.x);
            f[j * 4 + 1] = geglu_fast(__uint_as_float(va[j * 4 + 1]) + ba.y, __uint_as_float(vg[j * 4 + 1]) + bg.y);
            f[j * 4 + 2] = geglu_fast(__uint_as_float(va[j * 4 + 2]) + ba.z, __uint_as_float(vg[j * 4 + 2]) + bg.z);
            f[j * 4 + 3] = geglu_fast(__uint_as_float(va[j * 4 + 3]) + ba.w, __uint_as_float(vg[j * 4 + 3]) + bg.w);
          }
          __syncwarp();                              // every lane has read the bias before the buffer is reused
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int slot = (j ^ stage_swz<CH>(lane)) & (CH - 1);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * (CG * 4) + slot * 16), "f"(f[j * 4]),
                         "f"(f[j * 4 + 1]), "f"(f[j * 4 + 2]), "f"(f[j * 4 + 3])
                         : "memory");
          }
        }
        __syncwarp();
        // ---------------- phase B
        {
          const int n = n_out_base + c0 + ch * 4;
          const bool n_ok = n + 4 <= n_out_total;
          float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
          float sa = 1.0f;
          if (act != 2) {
            sa = p.s_acc;
            if (has_bias && n_ok) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
              bs = make_float4(b4.x * sa, b4.y * sa, b4.z * sa, b4.w * sa);
            }
          }
          char* obase = reinterpret_cast<char*>(p.out) + (long long)n * out_es;
          const char* r2base = r2p + (long long)n * 2;
          const char* rvbase = reinterpret_cast<const char*>(p.rowvec) + (long long)n * 4;
          const uint2(&u1)[NI] = rpre[k & 1];
          float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};   // STATS: column sums over this lane's rows
#pragma unroll
          for (int hb = 0; hb < NI / 4; ++hb) {
            float4 v[4], rv[4];
            uint2 u2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = hb * 4 + q;
              const int row = i * RPI + rsub;
              const int slot = (ch ^ stage_swz<CH>(row)) & (CH - 1);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(v[q].x), "=f"(v[q].y), "=f"(v[q].z), "=f"(v[q].w)
                           : "r"(stg + row * (CG * 4) + slot * 16));
              const bool ok = n_ok && ((vrows >> (RPI * i)) & 1u);
              rv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              u2[q] = make_uint2(0, 0);
              if (has_rv && ok) rv[q] = __ldg(reinterpret_cast<const float4*>(rvbase + (long long)rvrow[i] * p.ld_rowvec_b));
              if (has_r2 && ok) u2[q] = __ldg(reinterpret_cast<const uint2*>(r2base + (long long)tok[i] * p.ld_res2_b));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = hb * 4 + q;
              const bool ok = n_ok && ((vrows >> (RPI * i)) & 1u);   // straight-line code, predicated store
              float4 o = v[q];
              if (act != 2) {
                o.x = fmaf(o.x, sa, bs.x); o.y = fmaf(o.y, sa, bs.y); o.z = fmaf(o.z, sa, bs.z); o.w = fmaf(o.w, sa, bs.w);
                if (has_rv) { o.x += rv[q].x; o.y += rv[q].y; o.z += rv[q].z; o.w += rv[q].w; }
                if (act == 1) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
                if (has_r1) {
                  const float2 a = unpack2(u1[i].x, bf16), b = unpack2(u1[i].y, bf16);
                  o.x = fmaf(p.s_res1, a.x, o.x); o.y = fmaf(p.s_res1, a.y, o.y);
                  o.z = fmaf(p.s_res1, b.x, o.z); o.w = fmaf(p.s_res1, b.y, o.w);
                }
                if (has_r2) {
                  const float2 a = unpack2(u2[q].x, bf16), b = unpack2(u2[q].y, bf16);
                  o.x = fmaf(p.s_res2, a.x, o.x); o.y = fmaf(p.s_res2, a.y, o.y);
                  o.z = fmaf(p.s_res2, b.x, o.z); o.w = fmaf(p.s_res2, b.y, o.w);
                }
              }
              if (STATS && ok) {
                cs[0] += o.x; cs[1] += o.y; cs[2] += o.z; cs[3] += o.w;
                cq[0] = fmaf(o.x, o.x, cq[0]); cq[1] = fmaf(o.y, o.y, cq[1]);
                cq[2] = fmaf(o.z, o.z, cq[2]); cq[3] = fmaf(o.w, o.w, cq[3]);
              }
              char* optr = obase + (long long)tok[i] * p.ldo_b;
              if (f32o) {
                if (ok) *reinterpret_cast<float4*>(optr) = o;
              } else {
                const uint2 pk = make_uint2(pack2(o.x, o.y, bf16), pack2(o.z, o.w, bf16));
                if (ok) *reinterpret_cast<uint2*>(optr) = pk;
              }
            }
          }
          if (STATS) {
            // rows of one column live in the RPI lanes ch, ch + CH, ...: fixed-order butterfly, then lane rsub == 0 holds
            // the sums over the warp's 32 rows and writes them (one partial per tile and lane quadrant: bit-reproducible)
#pragma unroll
            for (int off = CH; off < 32; off <<= 1) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                cs[c] += __shfl_xor_sync(0xffffffffu, cs[c], off);
                cq[c] += __shfl_xor_sync(0xffffffffu, cq[c], off);
              }
            }
            if (rsub == 0 && n_ok) {
              float* dst = p.stats + (((long long)m_blk * 4 + quad) * p.stats_ld + p.stats_col0 + n) * 2;
              *reinterpret_cast<float4*>(dst) = make_float4(cs[0], cq[0], cs[1], cq[1]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(cs[2], cq[2], cs[3], cq[3]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                            // the accumulator stage of this CTA is drained
        if (PAIR) mbar_arrive_cluster(&tempty[as], 0); else mbar_arrive(&tempty[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }
  tc_fence_before();
  __syncwarp();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 4 * NQ + 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
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
    VB_REQUIRE(d->ntaps == 1 && d->h_pad == 0, "b200v_gemm: linear mode takes one tap and no halo rows");
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
    VB_REQUIRE(d->h_pad >= 0 && d->h_pad <= 4, "b200v_gemm: h_pad=%d out of range", d->h_pad);
    const uint64_t He = (uint64_t)d->H + 2ull * d->h_pad;     // rows of H in memory (halo rows before and after)
    uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->W, He, (uint64_t)d->NB};
    uint64_t strides[3] = {(uint64_t)d->lda * 2, (uint64_t)d->lda * 2 * d->W, (uint64_t)d->lda * 2 * d->W * He};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BB};
    uint32_t es[4] = {1, 1, 1, 1};
    if (encode_tmap_16bit(&tmA, d->a, 4, dims, strides, box, es, d->bf16)) return 3;
  }
  // CTA pairs (cta_group::2, M = 256 per MMA) halve the shared-memory operand traffic per SM.  Measured on B200
  // (profiles/r01_gemm_pair_vs_single.md) the pair kernel is on par for the large-K convolutions and slower for
  // the small-K projections (the two epilogues of a pair gate each other), so it is opt-in: VB_GEMM_PAIR=1.
  static const bool pair_enabled = getenv("VB_GEMM_PAIR") && atoi(getenv("VB_GEMM_PAIR")) != 0;
  const bool pair = pair_enabled && p.m_tiles >= 2 && !d->stats;
  const int b_rows = pair ? d->tile_n / 2 : d->tile_n;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)d->N};
    uint64_t strides[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, (uint32_t)b_rows};
    uint32_t es[2] = {1, 1};
    if (encode_tmap_16bit(&tmB, d->b, 2, dims, strides, box, es, d->bf16)) return 3;
  }
  p.ntaps = d->ntaps;
  p.kc_per_tap = d->cin / 64;
  for (int i = 0; i < 9; ++i) {
    p.dh[i] = d->dh[i] + (d->a_mode == 1 ? d->h_pad : 0);   // TMA coordinates count from the first halo row
    p.dw[i] = d->dw[i];
  }
  p.N = d->N;
  p.TN = d->tile_n;
  p.n_tiles = (d->N + d->tile_n - 1) / d->tile_n;
  p.bf16 = d->bf16;
  p.m_pairs = (p.m_tiles + 1) / 2;
  p.stage_bytes = kABytes + ((b_rows * 128 + 1023) / 1024) * 1024;
  constexpr int kStagingBytes = 8 * 32 * 128;  // epilogue transpose buffers (8 warps x 32 rows x 32 fp32)
  p.nstages = (227 * 1024 - 2048 - kStagingBytes) / p.stage_bytes;
  if (p.nstages > kMaxStages) p.nstages = kMaxStages;
  const long long kMaxLd = (1ll << 31) - 1;
  VB_REQUIRE(d->ldo * (d->out_f32 ? 4 : 2) <= kMaxLd && d->ld_rowvec * 4 <= kMaxLd && d->ld_res1 * 2 <= kMaxLd &&
                 d->ld_res2 * 2 <= kMaxLd,
             "b200v_gemm: row stride too large");
  for (p.bw_sh = 0; (1 << p.bw_sh) < p.BW; ++p.bw_sh) {}
  for (p.bh_sh = 0; (1 << p.bh_sh) < p.BH; ++p.bh_sh) {}
  p.out = d->out; p.ldo_b = (int)(d->ldo * (d->out_f32 ? 4 : 2)); p.out_f32 = d->out_f32; p.act = d->act;
  p.bias = d->bias;
  p.rowvec = d->rowvec; p.ld_rowvec_b = (int)(d->ld_rowvec * 4); p.rv_div = d->rv_div > 0 ? d->rv_div : 1;
  p.rv_mod = d->rv_mod > 0 ? d->rv_mod : 1;
  p.res1 = d->res1; p.ld_res1_b = (int)(d->ld_res1 * 2); p.s_res1 = d->s_res1;
  p.res2 = d->res2; p.ld_res2_b = (int)(d->ld_res2 * 2); p.s_res2 = d->s_res2;
  p.s_acc = d->s_acc;
  VB_REQUIRE(!(d->res2 && !d->res1), "b200v_gemm: res2 without res1");
  p.stats = d->stats; p.stats_ld = (int)d->stats_ld; p.stats_col0 = d->stats_col0;
  if (d->stats) {
    VB_REQUIRE(!d->bf16 && !d->out_f32 && d->act == 0 && !d->res2 && !(d->rowvec && d->res1),
               "b200v_gemm: fused statistics need fp16 output, act 0, <= 1 residual and no rowvec + residual");
    VB_REQUIRE(d->stats_ld >= d->stats_col0 + d->N && d->stats_col0 % 4 == 0 && d->stats_ld % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(d->stats) & 15) == 0,
               "b200v_gemm: bad stats_ld / stats_col0 / alignment");
    VB_REQUIRE(d->a_mode == 0 || (p.BB == 1 && d->W % p.BW == 0 && d->H % p.BH == 0 && (p.BW == d->W || p.BH == 1)),
               "b200v_gemm: fused statistics need token tiles of 128 consecutive tokens (box %d x %d x %d on %d x %d)",
               p.BW, p.BH, p.BB, d->W, d->H);
  }

  const int smem_bytes = 1024 + 1024 + p.nstages * p.stage_bytes + kStagingBytes;
  // Epilogue variant: the common fp16 feature sets are compiled in (no per-element feature tests), everything
  // else (bf16 operands, fp32 output, unusual combinations) takes the generic instantiation.
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const TGParams);
#define VB_VARIANTS(PAIR, NQ)                                                                                     \
  {tapgemm_kernel<0, false, 0, false, PAIR, NQ>, tapgemm_kernel<0, false, 1, false, PAIR, NQ>,                    \
   tapgemm_kernel<0, false, 2, false, PAIR, NQ>, tapgemm_kernel<0, true, 0, false, PAIR, NQ>,                     \
   tapgemm_kernel<0, true, 1, false, PAIR, NQ>,  tapgemm_kernel<1, false, 0, false, PAIR, NQ>,                    \
   tapgemm_kernel<2, false, 0, false, PAIR, NQ>, tapgemm_kernel<0, true, 2, true, PAIR, NQ>}
  // [pair][16 epilogue warps]: the CTA-pair kernel exists with 8 epilogue warps only
  static const Kern kVariants[2][2][8] = {{VB_VARIANTS(false, 2), VB_VARIANTS(false, 4)},
                                          {VB_VARIANTS(true, 2), VB_VARIANTS(true, 2)}};
#undef VB_VARIANTS
  int variant = 7;
  if (!d->bf16 && !d->out_f32 && !getenv("VB_GEMM_GENERIC")) {
    const int nres = (d->res1 ? 1 : 0) + (d->res2 ? 1 : 0);
    if (d->act == 0 && !d->rowvec) variant = nres;
    else if (d->act == 0 && nres <= 1) variant = 3 + nres;
    else if (d->act == 1 && !d->rowvec && nres == 0) variant = 5;
    else if (d->act == 2) variant = 6;
  }
  // 16 epilogue warps (4 per scheduler) hide the MUFU / dependency latency of the small-K GEGLU epilogue; everything
  // else keeps 8 warps and 32-column groups (fewer per-group fixed costs).  VB_GEMM_NQ=2|4 forces one.
  static const int nq_force = getenv("VB_GEMM_NQ") ? atoi(getenv("VB_GEMM_NQ")) : 0;
  int wide = (variant == 6 && K <= 384) ? 1 : 0;   // measured: L0 GEGLU (K = 320) 0.91 vs 0.94 ms, K = 640 the other way
  if (nq_force == 2) wide = 0;
  if (nq_force == 4) wide = 1;
  if (pair) wide = 0;
  // fused-statistics instantiations (8 epilogue warps, single CTA): plain, one residual, row vector
  static const Kern kStats[3] = {tapgemm_kernel<0, false, 0, false, false, 2, true>,
                                 tapgemm_kernel<0, false, 1, false, false, 2, true>,
                                 tapgemm_kernel<0, true, 0, false, false, 2, true>};
  static bool attr_set[64] = {false};
  static int max_clusters = 0;
  if (vb::first_use_on_device(attr_set)) {
    for (int i = 0; i < 3; ++i)
      VB_CHECK_CUDA(cudaFuncSetAttribute(kStats[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int q = 0; q < 2; ++q)
      for (int w = 0; w < 2; ++w)
        for (int i = 0; i < 8; ++i)
          VB_CHECK_CUDA(cudaFuncSetAttribute(kVariants[q][w][i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // how many CTA pairs the device runs at once (pairs sit on the two SMs of a TPC)
    cudaLaunchConfig_t qc = {};
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    qc.gridDim = dim3(2 * device_sm_count()); qc.blockDim = dim3(320); qc.dynamicSmemBytes = 227 * 1024;
    qc.attrs = qa; qc.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kVariants[1][0][0], &qc) != cudaSuccess || max_clusters <= 0) {
      cudaGetLastError();
      max_clusters = device_sm_count() / 2;
    }
  }
  if (d->stats) {
    const long long total = (long long)p.m_tiles * p.n_tiles;
    int grid = device_sm_count();
    if (total < grid) grid = (int)total;
    kStats[d->rowvec ? 2 : (d->res1 ? 1 : 0)]<<<grid, 320, smem_bytes, stream>>>(tmA, tmB, p);
  } else if (pair) {
    const long long total = (long long)p.m_pairs * p.n_tiles;
    int clusters = max_clusters;
    if (total < clusters) clusters = (int)total;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    VB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kVariants[1][0][variant], tmA, tmB, p));
  } else {
    const long long total = (long long)p.m_tiles * p.n_tiles;
    int grid = device_sm_count();
    if (total < grid) grid = (int)total;
    kVariants[0][wide][variant]<<<grid, wide ? 576 : 320, smem_bytes, stream>>>(tmA, tmB, p);
  }
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
