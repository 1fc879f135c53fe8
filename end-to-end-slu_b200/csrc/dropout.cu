// Dropout keep-mask generator (nn.Dropout at reference models.py:246/276/700, training mode).
// One pass: Philox4x32-10 keyed by a 64-bit seed, counter = index of the float4 -> four Bernoulli(1-p) draws, written
// already scaled by 1/(1-p) as the fp32 mask rows the persistent-GRU kernels stream through their TMA ring.
// HBM-bound: 4 B written per element, nothing read.
#include "common.cuh"
#include "philox.cuh"

namespace {

__device__ __forceinline__ void philox4x32_10(uint64_t ctr, uint64_t seed, uint32_t (&out)[4]) { slu_philox4x32_10(ctr, seed, out); }

// The canonical GRU-layer mask (philox.cuh) written out as a tensor: what slu_gru_{fwd,bwd}_* generate in registers when they are
// given (drop_p, drop_seed) instead of a mask -- for tests and for callers that want to inspect / reuse the mask.
__global__ void __launch_bounds__(256) dropout_mask_gru_kernel(float* __restrict__ mask, int B, int T, uint32_t keep_threshold, float scale,
                                                               uint64_t seed) {
  const int groups = (T + 7) >> 3;
  const long n = (long)B * groups * 256;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i & 255);
    const long r = i >> 8;
    const int tg = (int)(r % groups), b = (int)(r / groups);
    uint32_t w[4];
    slu_gru_mask_draws(b, col, tg, seed, w);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = 8 * tg + k;
      if (t < T) mask[((long)b * T + t) * 256 + col] = slu_gru_mask_draw16(w, k) < keep_threshold ? scale : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ mask, long n, uint32_t keep_threshold, float scale,
                                                           uint64_t seed) {
  const long n4 = (n + 3) >> 2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint64_t)i, seed, r);
    float4 v;
    v.x = r[0] < keep_threshold ? scale : 0.f;     // P(keep) = keep_threshold / 2^32
    v.y = r[1] < keep_threshold ? scale : 0.f;
    v.z = r[2] < keep_threshold ? scale : 0.f;
    v.w = r[3] < keep_threshold ? scale : 0.f;
    if (4 * i + 3 < n) {
      reinterpret_cast<float4*>(mask)[i] = v;
    } else {
      const float t[4] = {v.x, v.y, v.z, v.w};
      for (long j = 4 * i; j < n; ++j) mask[j] = t[j - 4 * i];
    }
  }
}

}  // namespace

extern "C" int slu_dropout_mask(float* mask, long n, float p, unsigned long long seed, void* stream) {
  if (n <= 0) return 0;
  if (!(p >= 0.f && p < 1.f)) return (int)cudaErrorInvalidValue;
  const double keep = 1.0 - (double)p;
  const uint32_t threshold = slu_keep_threshold(p);
  const long n4 = (n + 3) >> 2;
  long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  dropout_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mask, n, threshold, (float)(1.0 / keep), seed);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_dropout_mask_gru(float* mask, int B, int T, float p, unsigned long long seed, void* stream) {
  if (B <= 0 || T <= 0) return 0;
  if (!(p >= 0.f && p < 1.f)) return (int)cudaErrorInvalidValue;
  const long n = (long)B * ((T + 7) >> 3) * 256;
  long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  dropout_mask_gru_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mask, B, T, slu_keep_threshold16(p), (float)(1.0 / (1.0 - (double)p)), seed);
  SLU_CHECK_LAUNCH();
  return 0;
}

// ---- LeakyReLU backward + bias gradient of the conv blocks (autograd of models.py:211 LeakyReLU and the Conv1d bias) ------------
//   dpre[r][c] = y[r][c] > 0 ? gy[r][c] : slope * gy[r][c];     db[c] += sum_r dpre[r][c]
// One pass over the two inputs (HBM-bound: 12 B per element) instead of compare / multiply / select / reduce launches.
// Rows are C floats (C % 4 == 0, C <= 1024); a block is (C/4) x RB threads, so every thread keeps the same 4 columns.
namespace {

__global__ void __launch_bounds__(256) leaky_bwd_bias_kernel(const float* __restrict__ y, const float* __restrict__ gy, float slope,
                                                             float* __restrict__ dpre, float* __restrict__ db, long R, int C4, int RB) {
  extern __shared__ float4 part[];                       // [RB][C4] block partials
  const int tid = threadIdx.x;
  const int cq = tid % C4, rl = tid / C4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < RB) {
    for (long r = (long)blockIdx.x * RB + rl; r < R; r += (long)gridDim.x * RB) {
      const long i = r * C4 + cq;
      const float4 a = __ldg(reinterpret_cast<const float4*>(y) + i), g = __ldg(reinterpret_cast<const float4*>(gy) + i);
      float4 d;
      d.x = a.x > 0.f ? g.x : g.x * slope; d.y = a.y > 0.f ? g.y : g.y * slope;
      d.z = a.z > 0.f ? g.z : g.z * slope; d.w = a.w > 0.f ? g.w : g.w * slope;
      reinterpret_cast<float4*>(dpre)[i] = d;
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
    }
    part[rl * C4 + cq] = acc;
  }
  __syncthreads();
  if (tid < C4) {
    float4 s = part[tid];
    for (int k = 1; k < RB; ++k) {
      const float4 v = part[k * C4 + tid];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    atomicAdd(db + 4 * tid, s.x); atomicAdd(db + 4 * tid + 1, s.y); atomicAdd(db + 4 * tid + 2, s.z); atomicAdd(db + 4 * tid + 3, s.w);
  }
}

}  // namespace

extern "C" int slu_leaky_bwd_bias(const float* y, const float* gy, float slope, float* dpre, float* db, long R, int C, void* stream) {
  if (R <= 0) return 0;
  if (C <= 0 || (C & 3) || C > 1024) return (int)cudaErrorInvalidValue;
  if (((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(dpre)) & 15) != 0)
    return (int)cudaErrorInvalidValue;
  const int C4 = C / 4, RB = 256 / C4;
  long blocks = (R + RB - 1) / RB;
  if (blocks > 148 * 8) blocks = 148 * 8;
  leaky_bwd_bias_kernel<<<(unsigned)blocks, 256, (size_t)RB * C4 * sizeof(float4), (cudaStream_t)stream>>>(y, gy, slope, dpre, db, R, C4, RB);
  SLU_CHECK_LAUNCH();
  return 0;
}
