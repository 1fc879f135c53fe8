// Optimizer step and gradient-bucket glue of the train step (reference training.py:19, 64-66, 96-98:
// `torch.optim.Adam(model.parameters())`, `zero_grad(); backward(); step()`).
//
//   slu_adam_multi   ONE launch updates up to 64 parameter tensors (fp32 or fp64) of any sizes: the tensor table travels
//                    in the kernel parameters (no device-side table, nothing to upload or keep in sync), every CTA owns one
//                    4096-element chunk of one tensor.  The arithmetic is torch.optim.Adam's single-tensor path
//                    (lerp / addcmul / sqrt / addcdiv) with per-tensor step counts, so parameters that were un-frozen later
//                    (models.py:754-795) keep their own bias corrections.  HBM-bound: 4 reads + 3 writes per element.
//   slu_f64_hilo_*   the fp64 SincNet cut-off gradients travel through the single fp32 all-reduce bucket as (hi, lo) float pairs.
#include "common.cuh"

namespace {

constexpr int ADAM_MAX = 64;          // tensors per launch (64 x 48 B + prefix table < 4 KB of kernel parameters)
constexpr int ADAM_CHUNK = 4096;      // elements per CTA
constexpr int ADAM_THREADS = 256;

struct AdamTensor {                   // must match struct SluAdamTensor in include/slu_b200.h
  void* p;
  const void* g;
  void* m;
  void* v;
  long n;
  float step_size;                    // lr / (1 - beta1^t)
  float bc2_sqrt;                     // sqrt(1 - beta2^t)
  int is_f64;
  int pad;
};

struct AdamTable {
  AdamTensor t[ADAM_MAX];
  int first_block[ADAM_MAX + 1];      // CTA index of each tensor's first chunk
  int n;
};

// omb1 = 1 - beta1 and omb2 = 1 - beta2 are formed on the host in double precision (as torch does) and only then rounded.
template <typename T>
__device__ __forceinline__ void adam_elem(T& p, T g, T& m, T& v, T omb1, T beta2, T omb2, T eps, T wd, T step_size, T bc2_sqrt) {
  if (wd != T(0)) g += wd * p;                       // L2 penalty (torch.optim.Adam weight_decay)
  m = m + (g - m) * omb1;                            // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + omb2 * g * g;                      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const T denom = sqrt(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);                   // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_multi_kernel(const __grid_constant__ AdamTable tab, double beta1d, double beta2d,
                                                                  float eps, float wd) {
  const float omb1 = (float)(1.0 - beta1d), beta2 = (float)beta2d, omb2 = (float)(1.0 - beta2d);
  // which tensor does this CTA belong to: first_block is ascending, n <= 64
  int lo = 0, hi = tab.n - 1;
  const int blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.first_block[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  const AdamTensor& t = tab.t[lo];
  const long base = (long)(blk - tab.first_block[lo]) * ADAM_CHUNK;
  const long end = base + ADAM_CHUNK < t.n ? base + ADAM_CHUNK : t.n;
  if (t.is_f64) {
    double* p = (double*)t.p; const double* g = (const double*)t.g; double* m = (double*)t.m; double* v = (double*)t.v;
    for (long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
      double pi = p[i], mi = m[i], vi = v[i];
      adam_elem<double>(pi, g[i], mi, vi, 1.0 - beta1d, beta2d, 1.0 - beta2d, (double)eps, (double)wd, (double)t.step_size, (double)t.bc2_sqrt);
      p[i] = pi; m[i] = mi; v[i] = vi;
    }
    return;
  }
  float* p = (float*)t.p; const float* g = (const float*)t.g; float* m = (float*)t.m; float* v = (float*)t.v;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && (end - base) == ADAM_CHUNK;
  if (vec) {                                         // full, 16-byte aligned chunk: float4 traffic
#pragma unroll
    for (int r = 0; r < ADAM_CHUNK / (4 * ADAM_THREADS); ++r) {
      const long i = base + (long)(r * ADAM_THREADS + threadIdx.x) * 4;
      float4 pi = *reinterpret_cast<float4*>(p + i), mi = *reinterpret_cast<float4*>(m + i), vi = *reinterpret_cast<float4*>(v + i);
      const float4 gi = __ldg(reinterpret_cast<const float4*>(g + i));
      adam_elem<float>(pi.x, gi.x, mi.x, vi.x, omb1, beta2, omb2, eps, wd, t.step_size, t.bc2_sqrt);
      adam_elem<float>(pi.y, gi.y, mi.y, vi.y, omb1, beta2, omb2, eps, wd, t.step_size, t.bc2_sqrt);
      adam_elem<float>(pi.z, gi.z, mi.z, vi.z, omb1, beta2, omb2, eps, wd, t.step_size, t.bc2_sqrt);
      adam_elem<float>(pi.w, gi.w, mi.w, vi.w, omb1, beta2, omb2, eps, wd, t.step_size, t.bc2_sqrt);
      *reinterpret_cast<float4*>(p + i) = pi; *reinterpret_cast<float4*>(m + i) = mi; *reinterpret_cast<float4*>(v + i) = vi;
    }
    return;
  }
  for (long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_elem<float>(pi, g[i], mi, vi, omb1, beta2, omb2, eps, wd, t.step_size, t.bc2_sqrt);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

__global__ void f64_hilo_split_kernel(const double* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double x = src[i];
    const float h = (float)x;
    hi[i] = h;
    lo[i] = (float)(x - (double)h);
  }
}

__global__ void f64_hilo_merge_kernel(double* __restrict__ dst, const float* __restrict__ hi, const float* __restrict__ lo, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)hi[i] + (double)lo[i];
}

__global__ void seed_advance_kernel(unsigned long long* state) {
  state[0] = state[0] * 6364136223846793005ull + 1442695040888963407ull;      // 64-bit LCG step
}

}  // namespace

// The per-step word the dropout seeds are mixed with when a train step runs as a CUDA graph (the by-value seeds are frozen into
// the graph; this kernel is its first node, so every replay draws fresh masks).
extern "C" int slu_seed_advance(unsigned long long* state, void* stream) {
  seed_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_adam_multi(const void* tensors, int n, double beta1, double beta2, float eps, float weight_decay, void* stream) {
  if (n < 0) return (int)cudaErrorInvalidValue;
  const AdamTensor* src = (const AdamTensor*)tensors;
  for (int base = 0; base < n; base += ADAM_MAX) {
    AdamTable tab;
    tab.n = n - base < ADAM_MAX ? n - base : ADAM_MAX;
    int blocks = 0;
    for (int i = 0; i < tab.n; ++i) {
      tab.t[i] = src[base + i];
      if (tab.t[i].n < 0 || !tab.t[i].p || !tab.t[i].g || !tab.t[i].m || !tab.t[i].v) return (int)cudaErrorInvalidValue;
      tab.first_block[i] = blocks;
      blocks += (int)((tab.t[i].n + ADAM_CHUNK - 1) / ADAM_CHUNK);
    }
    for (int i = tab.n; i <= ADAM_MAX; ++i) tab.first_block[i] = blocks;
    if (blocks == 0) continue;
    adam_multi_kernel<<<blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(tab, beta1, beta2, eps, weight_decay);
    SLU_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int slu_f64_hilo_split(const double* src, float* hi, float* lo, int n, void* stream) {
  if (n <= 0) return n == 0 ? 0 : (int)cudaErrorInvalidValue;
  f64_hilo_split_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, hi, lo, n);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_f64_hilo_merge(double* dst, const float* hi, const float* lo, int n, void* stream) {
  if (n <= 0) return n == 0 ? 0 : (int)cudaErrorInvalidValue;
  f64_hilo_merge_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dst, hi, lo, n);
  SLU_CHECK_LAUNCH();
  return 0;
}
