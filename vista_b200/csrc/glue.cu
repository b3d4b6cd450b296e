// Callers' glue of the hot path as device code, all HBM-bound byte / elementwise work:
//   time_mix_small_u8   AE3DConv.time_mix_conv (temporal_ae.py:90-97) writing the frames the way the reference's output
//                       path stores them: clamp((x + 1) / 2, 0, 1) (sample_utils.py:374) -> 255 * s -> uint8 truncation,
//                       "t c h w -> t h w c" (sample_utils.py:96-126), fused into the decoder's last kernel;
//   rollout_advance     the latent bookkeeping between two rounds of the long-horizon rollout (sample_utils.py:335-365):
//                       sample[0] = z[0] (first round), samples_z[...] = sample[...], fill_latent(sample[-3:], ...);
//   ensemble_reward     reward_utils.py:318-337: exp(-mean_i var_k(sample_k[i])) with the unbiased variance over the
//                       ensemble, fixed-order fp64 reduction (bit-reproducible, no float atomics).
#include "../../include/vista_b200.h"
#include "host.cuh"

namespace vb {

__device__ __forceinline__ uint8_t to_u8(float v) {
  // numpy: (255.0 * clamp((v + 1.0) / 2.0, 0, 1)).astype(uint8) on float32 data — same operation order
  float s = (v + 1.0f) / 2.0f;
  s = fminf(fmaxf(s, 0.0f), 1.0f);
  return (uint8_t)(int)(255.0f * s);
}

// thread = (frame t, pixel).  fp32 NCHW `out` is read for blended frames (the previous chunk left its unblended value
// there) and written for frames >= keep_f32_from (the ones a later chunk will blend with) or when keep_f32_from < 0 (all).
__global__ void time_mix_small_u8_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                         const float* __restrict__ bias, float* __restrict__ out,
                                         uint8_t* __restrict__ out_u8, const int* __restrict__ blend, int T, int HW, int C,
                                         int out_frame0, int skip_frames, int keep_f32_from, long long ldx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)T * HW) return;
  const int t = (int)(i / HW), pix = (int)(i % HW);
  if (t < skip_frames) return;
  float acc[4];
  for (int co = 0; co < C; ++co) acc[co] = bias ? bias[co] : 0.f;
  for (int kt = 0; kt < 3; ++kt) {
    const int tt = t + kt - 1;
    if (tt < 0 || tt >= T) continue;
    const float* xp = x + ((long long)tt * HW + pix) * ldx;
    for (int ci = 0; ci < C; ++ci) {
      const float xv = xp[ci];
      for (int co = 0; co < C; ++co) acc[co] = fmaf(xv, w[(co * C + ci) * 3 + kt], acc[co]);
    }
  }
  const int mode = blend ? blend[t] : 0;
  const bool keep = keep_f32_from < 0 || t >= keep_f32_from;
  uint8_t* up = out_u8 + ((long long)(out_frame0 + t) * HW + pix) * C;
  for (int co = 0; co < C; ++co) {
    float* op = out + ((long long)(out_frame0 + t) * C + co) * HW + pix;
    const float v = mode ? 0.5f * (*op + acc[co]) : acc[co];
    if (keep) *op = v;
    up[co] = to_u8(v);
  }
}

// One launch per round.  sample (T, E) fp32 is the round's result (E = elements per frame); mutated like the reference
// does it: sample[0] = z0 on the first round.  samples_z rows [dst0 + src0, dst0 + T) <- sample rows [src0, T);
// filled (T, E) <- zeros with rows 0..n_cond-1 = the last n_cond rows of sample (fill_latent, sample_utils.py:280-283).
__global__ void rollout_advance_kernel(float* __restrict__ sample, const float* __restrict__ z0, float* __restrict__ samples_z,
                                       float* __restrict__ filled, int T, long long E, int dst0, int src0, int n_cond) {
  const long long n = (long long)T * E;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / E);
    const long long e = i - (long long)t * E;
    float v = sample[i];
    if (z0 != nullptr && t == 0) {
      v = z0[e];
      sample[i] = v;
    }
    if (t >= src0) samples_z[((long long)dst0 + t) * E + e] = v;
    if (filled != nullptr) {
      // row t of `filled` takes sample row T - n_cond + t (t < n_cond), zero otherwise
      filled[i] = t < n_cond ? (z0 != nullptr && T - n_cond + t == 0 ? z0[e] : sample[((long long)(T - n_cond + t)) * E + e]) : 0.0f;
    }
  }
}

constexpr int kRewardBlock = 256;

// partial[b] = sum over the block's elements of var_k (fp64, fixed order: thread-serial over a grid-stride range, then a
// shared-memory tree); the last block (ticket) adds the partials in index order and writes mean variance + reward.
__global__ void ensemble_reward_kernel(const float* const* __restrict__ members, int K, long long n,
                                       double* __restrict__ partial, unsigned int* __restrict__ ticket,
                                       float* __restrict__ out /* [2]: mean variance, reward */) {
  __shared__ double red[kRewardBlock];
  __shared__ bool last;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float u = 0.f;
    for (int k = 0; k < K; ++k) u += members[k][i];       // torch.mean(torch.stack(...), 0): fp32 sum / K
    u /= (float)K;
    float d = 0.f;
    for (int k = 0; k < K; ++k) {
      const float e = members[k][i] - u;
      d += e * e;                                         // diff.add_((each - u) ** 2)
    }
    acc += (double)(d / (float)(K - 1));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kRewardBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = red[0];
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += partial[b];
    const double mv = s / (double)n;
    out[0] = (float)mv;
    out[1] = expf(-(float)mv);
    *ticket = 0;                                          // self-resetting: the next launch needs no memset
  }
}

}  // namespace vb

extern "C" int b200v_time_mix_small_u8(const float* x, int64_t ldx, const float* w, const float* bias, float* out,
                                       uint8_t* out_u8, const int32_t* blend, int32_t T, int32_t HW, int32_t C,
                                       int32_t out_frame0, int32_t skip_frames, int32_t keep_f32_from, void* stream) {
  VB_REQUIRE(x && w && out && out_u8 && C >= 1 && C <= 4, "time_mix_small_u8: bad args");
  const long long total = (long long)T * HW;
  vb::time_mix_small_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, w, bias, out, out_u8, blend, T, HW, C, out_frame0, skip_frames, keep_f32_from, ldx);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200v_rollout_advance(float* sample, const float* z0, float* samples_z, float* filled, int32_t T,
                                     int64_t frame_elems, int32_t dst_frame0, int32_t src_frame0, int32_t n_cond,
                                     void* stream) {
  VB_REQUIRE(sample && samples_z && T > 0 && frame_elems > 0 && src_frame0 >= 0 && src_frame0 <= T && n_cond >= 0 &&
                 n_cond <= T && dst_frame0 >= 0,
             "rollout_advance: bad args");
  const long long n = (long long)T * frame_elems;
  long long blocks = (n + 255) / 256;
  const long long cap = 8ll * vb::device_sm_count();
  if (blocks > cap) blocks = cap;
  vb::rollout_advance_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sample, z0, samples_z, filled, T, frame_elems,
                                                                                 dst_frame0, src_frame0, n_cond);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200v_ensemble_reward_scratch(void) { return 4 * 1024; }   /* doubles of `partial` (>= grid) */

extern "C" int b200v_ensemble_reward(const float* const* members_dev, int32_t K, int64_t n, double* partial,
                                     uint32_t* ticket, float* out2, void* stream) {
  VB_REQUIRE(members_dev && partial && ticket && out2 && K >= 2 && K <= 64 && n > 0, "ensemble_reward: bad args (2 <= K <= 64)");
  long long blocks = (n + vb::kRewardBlock - 1) / vb::kRewardBlock;
  long long cap = 4ll * vb::device_sm_count();
  if (cap > 4 * 1024) cap = 4 * 1024;
  if (blocks > cap) blocks = cap;
  vb::ensemble_reward_kernel<<<(unsigned)blocks, vb::kRewardBlock, 0, (cudaStream_t)stream>>>(members_dev, K, n, partial, ticket, out2);
  VB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
