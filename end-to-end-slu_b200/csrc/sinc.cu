// SincNet front end: filter-bank synthesis (fwd + analytic bwd) and the strided sinc
// convolution fused with abs + max-pool(2, ceil) (fwd) / its filter-gradient (bwd).
// Reference behaviour restated: models.py:17-24, 77-110 (SincLayer), 163-168 (Abs), 205 (MaxPool1d).
// Layout: waveform x[B][T] fp32; output frames NLC  out[B][L1][80]  (L0=(T-1)/80+1, L1=ceil(L0/2)).
#include "common.cuh"
#include <math.h>

namespace {

__device__ __forceinline__ float lowpass_tap(float fl, float band, int k) {
  // 2*f*sinc(2*pi*f*fs*t), t = |k-200|/fs ; op order as models.py:18,99  (fp32)
  if (k == SLU_PAD) return 2.0f * fl;
  float n = (float)abs(k - SLU_PAD);
  float t = n / 16000.0f;
  float arg = (6.283185307179586f * band) * t;
  return (2.0f * fl) * (sinf(arg) / arg);
}
__device__ __forceinline__ float hamming_tap(int k) {
  // window = 0.54 - 0.46*cos(2*pi*n/N), n = linspace(0, N, steps=N)  (models.py:91-94)
  float n = (float)k * ((float)SLU_NTAPS / (float)(SLU_NTAPS - 1));
  return 0.54f - 0.46f * cosf((6.283185307179586f * n) / (float)SLU_NTAPS);
}

// block-wide reductions over 512 threads
__device__ float block_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(512) sinc_filters_fwd_kernel(const double* __restrict__ b1, const double* __restrict__ band,
                                                               float* __restrict__ W) {
  __shared__ float red[16];
  const int c = blockIdx.x, k = threadIdx.x;
  const double f1d = fabs(b1[c]) + 50.0 / 16000.0;
  const double f2d = f1d + (fabs(band[c]) + 50.0 / 16000.0);
  const float f1 = (float)f1d, f2 = (float)f2d;
  float bp = -INFINITY;
  if (k < SLU_NTAPS) bp = lowpass_tap(f2, f2 * 16000.0f, k) - lowpass_tap(f1, f1 * 16000.0f, k);
  const float m = block_max(bp, red);
  if (k < SLU_NTAPS) W[c * SLU_NTAPS + k] = (bp / m) * hamming_tap(k);
}

// dL/d(filt_b1), dL/d(filt_band) from dL/dW.  Chain: W = bp/max(bp)*win, bp = lp(f2)-lp(f1),
// d lp(f)[k]/df = 2cos(2*pi*f*fs*t_k) (2 at the centre tap), f1=|b1|+c0, f2=f1+|band|+c0.
// torch.max() (full reduce, models.py:103) spreads its gradient evenly over tied maxima.
__global__ void __launch_bounds__(512) sinc_filters_bwd_kernel(const double* __restrict__ b1, const double* __restrict__ band,
                                                               const float* __restrict__ dW, double* __restrict__ d_b1,
                                                               double* __restrict__ d_band) {
  __shared__ float red[16];
  __shared__ double dred[16];
  const int c = blockIdx.x, k = threadIdx.x;
  const double f1d = fabs(b1[c]) + 50.0 / 16000.0;
  const double f2d = f1d + (fabs(band[c]) + 50.0 / 16000.0);
  const float f1 = (float)f1d, f2 = (float)f2d;
  const bool on = k < SLU_NTAPS;
  float bp = -INFINITY, g = 0.f;
  if (on) {
    bp = lowpass_tap(f2, f2 * 16000.0f, k) - lowpass_tap(f1, f1 * 16000.0f, k);
    g = dW[c * SLU_NTAPS + k] * hamming_tap(k);
  }
  const float m = block_max(bp, red);
  const double ties = block_sum((on && bp == m) ? 1.0 : 0.0, dred);
  const double S = block_sum(on ? (double)g * (double)bp : 0.0, dred);
  double dbp = 0.0;
  if (on) {
    dbp = (double)g / (double)m;
    if (bp == m) dbp -= S / ((double)m * (double)m) / ties;
  }
  double c1 = 2.0, c2 = 2.0;
  if (on && k != SLU_PAD) {
    float t = (float)abs(k - SLU_PAD) / 16000.0f;
    c1 = 2.0 * (double)cosf((6.283185307179586f * (f1 * 16000.0f)) * t);
    c2 = 2.0 * (double)cosf((6.283185307179586f * (f2 * 16000.0f)) * t);
  }
  const double df2 = block_sum(on ? dbp * c2 : 0.0, dred);
  const double df1 = block_sum(on ? -dbp * c1 : 0.0, dred);
  if (k == 0) {
    const double sb1 = (b1[c] > 0) - (b1[c] < 0), sbd = (band[c] > 0) - (band[c] < 0);
    d_b1[c] = (df1 + df2) * sb1;
    d_band[c] = df2 * sbd;
  }
}

// Jacobian banks of the filter synthesis: J[0][c][k] = dW[c][k]/d filt_b1[c], J[1][c][k] = dW[c][k]/d filt_band[c] (fp32 values
// of an fp64 evaluation of the same chain as sinc_filters_bwd_kernel, the max-normalisation term included).  With them
//     dL/d theta[c] = sum_k dL/dW[c][k] J_theta[c][k] = sum_{b,t} g0[b][t][c] * (x * J_theta[c])[b][t]
// i.e. the cut-off gradients are two more strided convolutions of the waveform (banks J_0, J_1) dotted with the routed output
// gradient -- the heavy cancellation between the direct and the normalisation term happens HERE, analytically, instead of
// between the rounding errors of a low-precision dW (sinc_tc.cu).
__global__ void __launch_bounds__(512) sinc_filters_jac_kernel(const double* __restrict__ b1, const double* __restrict__ band,
                                                               float* __restrict__ J) {
  __shared__ float red[16];
  __shared__ double dred[16];
  const int c = blockIdx.x, k = threadIdx.x;
  const double f1d = fabs(b1[c]) + 50.0 / 16000.0;
  const double f2d = f1d + (fabs(band[c]) + 50.0 / 16000.0);
  const float f1 = (float)f1d, f2 = (float)f2d;
  const bool on = k < SLU_NTAPS;
  float bp = -INFINITY;
  if (on) bp = lowpass_tap(f2, f2 * 16000.0f, k) - lowpass_tap(f1, f1 * 16000.0f, k);
  const float m = block_max(bp, red);
  double c1 = 2.0, c2 = 2.0;
  if (on && k != SLU_PAD) {
    float t = (float)abs(k - SLU_PAD) / 16000.0f;
    c1 = 2.0 * (double)cosf((6.283185307179586f * (f1 * 16000.0f)) * t);
    c2 = 2.0 * (double)cosf((6.283185307179586f * (f2 * 16000.0f)) * t);
  }
  const bool tie = on && bp == m;
  const double ties = block_sum(tie ? 1.0 : 0.0, dred);
  const double C1 = block_sum(tie ? c1 : 0.0, dred) / ties;       // d max(bp) / d f1 = -C1, d max(bp) / d f2 = C2
  const double C2 = block_sum(tie ? c2 : 0.0, dred) / ties;
  if (on) {
    const double win = (double)hamming_tap(k), md = (double)m, bpd = (double)bp;
    const double j2 = win * (c2 / md - bpd * C2 / (md * md));       // dW/df2
    const double j1 = win * (-c1 / md + bpd * C1 / (md * md));      // dW/df1
    const double sb1 = (b1[c] > 0) - (b1[c] < 0), sbd = (band[c] > 0) - (band[c] < 0);
    J[(0 * SLU_NFILT + c) * SLU_NTAPS + k] = (float)((j1 + j2) * sb1);   // f1 = |b1| + c0 feeds f2 as well
    J[(1 * SLU_NFILT + c) * SLU_NTAPS + k] = (float)(j2 * sbd);
  }
}

// ---------------- sinc conv + abs + maxpool(2, ceil): CUDA-core fp32 path ------------------------
// CTA = (utterance b, 32 pooled frames = 64 conv frames); filters are symmetric (W[c][k]==W[c][400-k]),
// so each thread folds the waveform window:  out = sum_{k<200} W[k]*(x[a+k]+x[a+400-k]) + W[200]*x[a+200].
constexpr int SC_TP = 32;                 // pooled frames per CTA
constexpr int SC_TF = 2 * SC_TP;          // conv frames per CTA
constexpr int SC_SPAN = (SC_TF - 1) * SLU_STRIDE + SLU_NTAPS;   // 5441 samples
constexpr int SC_HALF = SLU_PAD + 1;      // 201 distinct taps

__global__ void __launch_bounds__(256) sincconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, int B, int T,
                                                           int L0, int L1, float* __restrict__ out, uint8_t* __restrict__ route) {
  extern __shared__ float sm[];
  float* xs = sm;                          // [SC_SPAN] (+pad)
  float* ws = sm + 5504;                   // [201][80]
  const int b = blockIdx.y, jp0 = blockIdx.x * SC_TP, tid = threadIdx.x;
  const long base = (long)jp0 * 2 * SLU_STRIDE - SLU_PAD;   // first sample of the window (may be < 0)
  const float* xb = x + (long)b * T;
  for (int i = tid; i < SC_SPAN; i += 256) {
    long s = base + i;
    xs[i] = (s >= 0 && s < T) ? xb[s] : 0.f;
  }
  for (int i = tid; i < SC_HALF * SLU_NFILT; i += 256) {
    int k = i / SLU_NFILT, c = i % SLU_NFILT;
    ws[i] = W[c * SLU_NTAPS + k];
  }
  __syncthreads();
  const int cg = tid & 15, tg = tid >> 4;       // 5 filters c = cg + 16 i ; 4 frames t = 4 tg + j
  float acc[4][5];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 5; ++i) acc[j][i] = 0.f;
  const float* xt = xs + tg * 4 * SLU_STRIDE;
#pragma unroll 4
  for (int k = 0; k < SLU_PAD; ++k) {
    float s[4], w[5];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = xt[j * SLU_STRIDE + k] + xt[j * SLU_STRIDE + 400 - k];
#pragma unroll
    for (int i = 0; i < 5; ++i) w[i] = ws[k * SLU_NFILT + cg + 16 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 5; ++i) acc[j][i] = fmaf(w[i], s[j], acc[j][i]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 5; ++i) acc[j][i] = fmaf(ws[SLU_PAD * SLU_NFILT + cg + 16 * i], xt[j * SLU_STRIDE + SLU_PAD], acc[j][i]);
  // abs + max-pool over frame pairs (2jp, 2jp+1); odd tail (ceil_mode) keeps the single frame
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int jp = jp0 + tg * 2 + p;
    if (jp >= L1) continue;
    const bool has1 = (2 * jp + 1) < L0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      float v0 = acc[2 * p][i], v1 = acc[2 * p + 1][i];
      float a0 = fabsf(v0), a1 = has1 ? fabsf(v1) : -1.f;
      int sel = (a1 > a0) ? 1 : 0;                       // first max wins on ties (torch max_pool1d)
      float v = sel ? v1 : v0;
      long o = ((long)b * L1 + jp) * SLU_NFILT + cg + 16 * i;
      out[o] = sel ? a1 : a0;
      if (route) route[o] = (uint8_t)(sel | ((v < 0.f) ? 2 : 0) | ((v == 0.f) ? 4 : 0));
    }
  }
}

// dW[c][k] = sum_{b,t} g0[b][t][c] * xpad[b][80 t + k],  g0 = gy routed through maxpool/abs.
// Persistent CTAs: thread k keeps the 80 partial sums of tap k in registers across many tiles.
constexpr int SB_TF = 32;                                      // conv frames per tile
constexpr int SB_SPAN = (SB_TF - 1) * SLU_STRIDE + SLU_NTAPS;  // 2881
__global__ void __launch_bounds__(416) sincconv_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                           const uint8_t* __restrict__ route, int B, int T, int L0, int L1,
                                                           float* __restrict__ dW) {
  __shared__ __align__(16) float xs[2944];
  __shared__ __align__(16) float gs[SB_TF * SLU_NFILT];
  const int tid = threadIdx.x;
  const int tiles_per_utt = (L0 + SB_TF - 1) / SB_TF;
  const long n_tiles = (long)B * tiles_per_utt;
  float acc[SLU_NFILT];
#pragma unroll
  for (int c = 0; c < SLU_NFILT; ++c) acc[c] = 0.f;
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_utt), t0 = (int)(tile % tiles_per_utt) * SB_TF;
    const long base = (long)t0 * SLU_STRIDE - SLU_PAD;
    const float* xb = x + (long)b * T;
    __syncthreads();
    for (int i = tid; i < SB_SPAN; i += 416) {
      long s = base + i;
      xs[i] = (s >= 0 && s < T) ? xb[s] : 0.f;
    }
    for (int i = tid; i < SB_TF * SLU_NFILT; i += 416) {
      int tl = i / SLU_NFILT, c = i % SLU_NFILT, t = t0 + tl;
      float g = 0.f;
      if (t < L0) {
        long o = ((long)b * L1 + (t >> 1)) * SLU_NFILT + c;
        uint8_t r = route[o];
        if ((r & 1) == (t & 1) && !(r & 4)) g = (r & 2) ? -gy[o] : gy[o];
      }
      gs[i] = g;
    }
    __syncthreads();
    if (tid < SLU_NTAPS) {
      for (int tl = 0; tl < SB_TF; ++tl) {
        const float xv = xs[tl * SLU_STRIDE + tid];
        const float4* g4 = reinterpret_cast<const float4*>(gs + tl * SLU_NFILT);
#pragma unroll
        for (int c4 = 0; c4 < SLU_NFILT / 4; ++c4) {
          float4 g = g4[c4];
          acc[4 * c4 + 0] = fmaf(g.x, xv, acc[4 * c4 + 0]);
          acc[4 * c4 + 1] = fmaf(g.y, xv, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(g.z, xv, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(g.w, xv, acc[4 * c4 + 3]);
        }
      }
    }
  }
  if (tid < SLU_NTAPS) {
#pragma unroll
    for (int c = 0; c < SLU_NFILT; ++c) atomicAdd(&dW[c * SLU_NTAPS + tid], acc[c]);
  }
}

}  // namespace

extern "C" int slu_sinc_filters_fwd(const double* filt_b1, const double* filt_band, float* W, void* stream) {
  sinc_filters_fwd_kernel<<<SLU_NFILT, 512, 0, (cudaStream_t)stream>>>(filt_b1, filt_band, W);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_sinc_filters_bwd(const double* filt_b1, const double* filt_band, const float* dW, double* d_b1,
                                    double* d_band, void* stream) {
  sinc_filters_bwd_kernel<<<SLU_NFILT, 512, 0, (cudaStream_t)stream>>>(filt_b1, filt_band, dW, d_b1, d_band);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_sinc_filters_jac(const double* filt_b1, const double* filt_band, float* J, void* stream) {
  sinc_filters_jac_kernel<<<SLU_NFILT, 512, 0, (cudaStream_t)stream>>>(filt_b1, filt_band, J);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_sincconv_fwd_simt(const float* x, const float* W, int B, int T, float* out, uint8_t* route, void* stream) {
  if (B <= 0 || T <= 0) return (int)cudaErrorInvalidValue;
  const int L0 = (T - 1) / SLU_STRIDE + 1, L1 = (L0 + 1) / 2;
  const size_t smem = (5504 + SC_HALF * SLU_NFILT) * sizeof(float);
  SLU_SMEM_ONCE(sincconv_fwd_kernel, smem);
  dim3 grid((L1 + SC_TP - 1) / SC_TP, B);
  sincconv_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, W, B, T, L0, L1, out, route);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_sincconv_bwd_simt(const float* x, const float* gy, const uint8_t* route, int B, int T, float* dW, void* stream) {
  if (B <= 0 || T <= 0) return (int)cudaErrorInvalidValue;
  const int L0 = (T - 1) / SLU_STRIDE + 1, L1 = (L0 + 1) / 2;
  cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * SLU_NFILT * SLU_NTAPS, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  const long n_tiles = (long)B * ((L0 + SB_TF - 1) / SB_TF);
  int grid = (int)(n_tiles < 296 ? n_tiles : 296);
  sincconv_bwd_kernel<<<grid, 416, 0, (cudaStream_t)stream>>>(x, gy, route, B, T, L0, L1, dW);
  SLU_CHECK_LAUNCH();
  return 0;
}
