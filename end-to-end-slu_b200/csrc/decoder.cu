// Seq2seq intent decoder, teacher-forced training path (reference models.py:413-436 Attention,
// 438-484 DecoderRNN = GRUCell stack, 500-556 Seq2SeqDecoder.forward).
//
// The reference runs, per output symbol, ~25 library launches forward (it even re-projects the encoder states to keys / values
// inside every step) and as many again in autograd.  Here everything that does not depend on the recurrence is hoisted into dense
// tcgen05 GEMMs over all symbols at once (keys, values, embeddings, the embedding half of the first cell's input projection, the
// output projection + log-softmax + NLL), and the sequential part of a step is three small fused kernels between four GEMM launches:
//   attn_step_*     one CTA per utterance: scores = keys.q / sqrt(K) -> softmax over the T encoder frames -> context = w.values;
//                   backward accumulates dkeys / dvalues over the steps in place (a CTA owns its utterance: no atomics)
//   grucell_*       the GRUCell gate math (torch.nn.GRUCell: r, z, n; h' = (1-z) n + z h) on pre-computed gi / gh, fused with the
//                   Dropout(0.5) that feeds the next cell (Philox keyed by (seed, step, element)); backward emits dgi / dgh and
//                   the direct dh path
// All tensors fp32; B = utterances, D = decoder width (256), T = encoder frames, K / V = key / value widths.
#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int ATT_THREADS = 256;      // 8 warps: 3-4 encoder frames per warp at T = 25
constexpr int ATT_MAXT = 256;        // encoder frames per utterance (4 s -> 25, 15 s -> 94)

__device__ __forceinline__ float block_sum_128(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < ATT_THREADS / 32; ++i) r += red[i];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max_128(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < ATT_THREADS / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}

// q [B][ldq] (first K used), keys [B][T][K], values [B][T][V] -> w [B][T], ctx [B][V]
__global__ void __launch_bounds__(ATT_THREADS) attn_step_fwd_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ keys,
                                                                    const float* __restrict__ values, int T, int K, int V,
                                                                    float inv_scale, float* __restrict__ w, float* __restrict__ ctx) {
  __shared__ float qs[512], sc[ATT_MAXT], red[ATT_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int k = tid; k < K; k += ATT_THREADS) qs[k] = q[(long)b * ldq + k];
  __syncthreads();
  for (int t = warp; t < T; t += ATT_THREADS / 32) {          // a warp per frame: coalesced key row, shuffle reduction
    const float* kr = keys + ((long)b * T + t) * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(kr[k], qs[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sc[t] = s * inv_scale;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = tid; t < T; t += ATT_THREADS) m = fmaxf(m, sc[t]);
  m = block_max_128(m, red);
  float e = 0.f;
  for (int t = tid; t < T; t += ATT_THREADS) { const float v = expf(sc[t] - m); sc[t] = v; e += v; }
  e = block_sum_128(e, red);
  const float inv = 1.f / e;
  for (int t = tid; t < T; t += ATT_THREADS) { const float v = sc[t] * inv; sc[t] = v; w[(long)b * T + t] = v; }
  __syncthreads();
  for (int v = tid; v < V; v += ATT_THREADS) {                // coalesced over v for every frame
    float a = 0.f;
    const float* vr = values + (long)b * T * V + v;
    for (int t = 0; t < T; ++t) a = fmaf(sc[t], vr[(long)t * V], a);
    ctx[(long)b * V + v] = a;
  }
}

// dctx [B][V], w [B][T], q [B][ldq] -> dq [B][lddq] (first K), dkeys [B][T][K] += , dvalues [B][T][V] +=
__global__ void __launch_bounds__(ATT_THREADS) attn_step_bwd_kernel(const float* __restrict__ dctx, const float* __restrict__ w,
                                                                    const float* __restrict__ q, long ldq, const float* __restrict__ keys,
                                                                    const float* __restrict__ values, int T, int K, int V,
                                                                    float inv_scale, float* __restrict__ dq, long lddq,
                                                                    float* __restrict__ dkeys, float* __restrict__ dvalues) {
  __shared__ float dc[512], qs[512], ws[ATT_MAXT], ds[ATT_MAXT], red[ATT_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int v = tid; v < V; v += ATT_THREADS) dc[v] = dctx[(long)b * V + v];
  for (int k = tid; k < K; k += ATT_THREADS) qs[k] = q[(long)b * ldq + k];
  for (int t = tid; t < T; t += ATT_THREADS) ws[t] = w[(long)b * T + t];
  __syncthreads();
  for (int t = warp; t < T; t += ATT_THREADS / 32) {          // dw[t] = dctx . values[t];  dvalues[t] += w[t] dctx
    const long o = ((long)b * T + t) * V;
    const float wt = ws[t];
    float s = 0.f;
    for (int v = lane; v < V; v += 32) {
      s = fmaf(dc[v], values[o + v], s);
      dvalues[o + v] += wt * dc[v];
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
    if (lane == 0) ds[t] = s;
  }
  __syncthreads();
  float dot = 0.f;
  for (int t = tid; t < T; t += ATT_THREADS) dot += ws[t] * ds[t];
  dot = block_sum_128(dot, red);
  for (int t = tid; t < T; t += ATT_THREADS) ds[t] = ws[t] * (ds[t] - dot) * inv_scale;     // d(scaled score)
  __syncthreads();
  for (int k = tid; k < K; k += ATT_THREADS) {
    float a = 0.f;
    const float qk = qs[k];
    for (int t = 0; t < T; ++t) {
      const long o = ((long)b * T + t) * K + k;
      a = fmaf(ds[t], keys[o], a);
      dkeys[o] += ds[t] * qk;
    }
    dq[(long)b * lddq + k] = a;
  }
}

__device__ __forceinline__ float cell_mask(uint64_t seed, int step, long elem, uint32_t thr, float scale) {
  uint32_t r[4];
  slu_philox4x32_10(((uint64_t)(uint32_t)step << 40) | (uint64_t)(elem >> 2), seed, r);
  const int k = (int)(elem & 3);
  const uint32_t word = k < 2 ? (k == 0 ? r[0] : r[1]) : (k == 2 ? r[2] : r[3]);
  return word < thr ? scale : 0.f;
}

// gi = gi_a[b*lda + .] (+ gi_b[b*ldb + .]), gh [b*ldh + .] (3D wide: r | z | n, biases included), hprev [B][D] (NULL: one row h0[D]
// broadcast) -> h [B][D], stash [B][4D] = r | z | n | gh_n, dropped [B][D] = h * mask (NULL: no dropout output wanted)
__global__ void grucell_fwd_kernel(const float* __restrict__ gi_a, long lda, const float* __restrict__ gi_b, long ldb,
                                   const float* __restrict__ gh, long ldh, const float* __restrict__ hprev, const float* __restrict__ h0,
                                   int B, int D, uint32_t thr, float scale, uint64_t seed_in, const unsigned long long* __restrict__ seed_dev,
                                   int step, float* __restrict__ h, float* __restrict__ stash, float* __restrict__ dropped) {
  const uint64_t seed = seed_dev ? (seed_in ^ (uint64_t)__ldg(seed_dev)) : seed_in;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * D) return;
  const int b = (int)(i / D), j = (int)(i - (long)b * D);
  const float* a = gi_a + (long)b * lda;
  const float* g = gh + (long)b * ldh;
  float ir = a[j], iz = a[D + j], in_ = a[2 * D + j];
  if (gi_b) { const float* c = gi_b + (long)b * ldb; ir += c[j]; iz += c[D + j]; in_ += c[2 * D + j]; }
  const float hn = g[2 * D + j];
  const float r = 1.f / (1.f + expf(-(ir + g[j])));
  const float z = 1.f / (1.f + expf(-(iz + g[D + j])));
  const float n = tanhf(in_ + r * hn);
  const float hp = hprev ? hprev[i] : h0[j];
  const float hv = (1.f - z) * n + z * hp;
  h[i] = hv;
  if (stash) { float* s = stash + (long)b * 4 * D; s[j] = r; s[D + j] = z; s[2 * D + j] = n; s[3 * D + j] = hn; }
  if (dropped) dropped[i] = thr ? hv * cell_mask(seed, step, i, thr, scale) : hv;
}

// dh = (da (* mask if thr) ) + db + dc   (any of db, dc may be NULL)  ->  dgi [B][ldgi] (dr, dz, dn pre-activations),
// dgh [B][ldgh] (dr, dz, dn*r), dh_direct [B][D] = dh * z
__global__ void grucell_bwd_kernel(const float* __restrict__ da, const float* __restrict__ db, const float* __restrict__ dc,
                                   const float* __restrict__ stash, const float* __restrict__ hprev, const float* __restrict__ h0, int B,
                                   int D, uint32_t thr, float scale, uint64_t seed_in, const unsigned long long* __restrict__ seed_dev,
                                   int step, float* __restrict__ dgi, long ldgi, float* __restrict__ dgh, long ldgh,
                                   float* __restrict__ dh_direct) {
  const uint64_t seed = seed_dev ? (seed_in ^ (uint64_t)__ldg(seed_dev)) : seed_in;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * D) return;
  const int b = (int)(i / D), j = (int)(i - (long)b * D);
  float dh = da[i];
  if (thr) dh *= cell_mask(seed, step, i, thr, scale);
  if (db) dh += db[i];
  if (dc) dh += dc[i];
  const float* s = stash + (long)b * 4 * D;
  const float r = s[j], z = s[D + j], n = s[2 * D + j], hn = s[3 * D + j];
  const float hp = hprev ? hprev[i] : h0[j];
  const float dn_pre = dh * (1.f - z) * (1.f - n * n);
  const float dz_pre = dh * (hp - n) * z * (1.f - z);
  const float dr_pre = dn_pre * hn * r * (1.f - r);
  float* gi = dgi + (long)b * ldgi;
  float* gh = dgh + (long)b * ldgh;
  gi[j] = dr_pre; gi[D + j] = dz_pre; gi[2 * D + j] = dn_pre;
  gh[j] = dr_pre; gh[D + j] = dz_pre; gh[2 * D + j] = dn_pre * r;
  dh_direct[i] = dh * z;
}

}  // namespace

extern "C" int slu_attn_step_fwd(const float* q, long ldq, const float* keys, const float* values, int B, int T, int K, int V,
                                 float inv_scale, float* w, float* ctx, void* stream) {
  if (B <= 0 || T <= 0 || T > ATT_MAXT || K <= 0 || K > 512 || V <= 0 || V > 512) return (int)cudaErrorInvalidValue;
  attn_step_fwd_kernel<<<B, ATT_THREADS, 0, (cudaStream_t)stream>>>(q, ldq, keys, values, T, K, V, inv_scale, w, ctx);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_attn_step_bwd(const float* dctx, const float* w, const float* q, long ldq, const float* keys, const float* values, int B,
                                 int T, int K, int V, float inv_scale, float* dq, long lddq, float* dkeys, float* dvalues, void* stream) {
  if (B <= 0 || T <= 0 || T > ATT_MAXT || K <= 0 || K > 512 || V <= 0 || V > 512) return (int)cudaErrorInvalidValue;
  attn_step_bwd_kernel<<<B, ATT_THREADS, 0, (cudaStream_t)stream>>>(dctx, w, q, ldq, keys, values, T, K, V, inv_scale, dq, lddq, dkeys,
                                                                    dvalues);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_grucell_fwd(const float* gi_a, long lda, const float* gi_b, long ldb, const float* gh, long ldh, const float* hprev,
                               const float* h0, int B, int D, float drop_p, unsigned long long drop_seed,
                               const unsigned long long* drop_seed_dev, int step, float* h, float* stash, float* dropped, void* stream) {
  if (B <= 0 || D <= 0 || (!hprev && !h0) || !(drop_p >= 0.f && drop_p < 1.f)) return (int)cudaErrorInvalidValue;
  const uint32_t thr = drop_p > 0.f ? slu_keep_threshold(drop_p) : 0u;
  const long n = (long)B * D;
  grucell_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      gi_a, lda, gi_b, ldb, gh, ldh, hprev, h0, B, D, thr, (float)(1.0 / (1.0 - (double)drop_p)), drop_seed, drop_seed_dev, step, h, stash, dropped);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_grucell_bwd(const float* da, const float* db, const float* dc, const float* stash, const float* hprev, const float* h0,
                               int B, int D, float drop_p, unsigned long long drop_seed, const unsigned long long* drop_seed_dev, int step,
                               float* dgi, long ldgi, float* dgh, long ldgh, float* dh_direct, void* stream) {
  if (B <= 0 || D <= 0 || (!hprev && !h0) || !(drop_p >= 0.f && drop_p < 1.f)) return (int)cudaErrorInvalidValue;
  const uint32_t thr = drop_p > 0.f ? slu_keep_threshold(drop_p) : 0u;
  const long n = (long)B * D;
  grucell_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      da, db, dc, stash, hprev, h0, B, D, thr, (float)(1.0 / (1.0 - (double)drop_p)), drop_seed, drop_seed_dev, step, dgi, ldgi, dgh, ldgh, dh_direct);
  SLU_CHECK_LAUNCH();
  return 0;
}

// ---- skinny GEMM for the decoder's per-symbol projections: C[m][n] = sum_k A[m*lda + k] * W[n*sn + k*sk] (+ bias[n]), M <= 64.
// These are latency-bound GEMV-like products (64 utterances x a few hundred outputs, 28 MFLOP at most): a persistent tensor-core
// pipeline costs more to start than the product takes, so this is an exact-fp32 CUDA-core kernel.  A CTA owns 8 output columns
// (N = 200..868 -> 25..109 CTAs), stages the activation rows and its weight rows through shared memory in K chunks of 128 and every
// thread accumulates 2 outputs with 128-bit shared-memory loads (3 LDS.128 per 8 FMAs: the loop is FMA-, not LSU-bound).
namespace {
constexpr int SK_M = 64, SK_N = 8, SK_KC = 128, SK_LD = SK_KC + 4;      // row pitch 132 floats: 16-byte rows, conflict-free LDS.128

__global__ void __launch_bounds__(256) skinny_gemm_kernel(const float* __restrict__ A, long lda, const float* __restrict__ W, long sn, long sk,
                                                          const float* __restrict__ bias, float* __restrict__ C, long ldc, int M, int N,
                                                          int K) {
  __shared__ __align__(16) float As[SK_M][SK_LD];
  __shared__ __align__(16) float Ws[SK_N][SK_LD];
  const int tid = threadIdx.x, m = tid & 63, cg = tid >> 6;            // 64 rows x 4 groups of 2 columns
  const int n0 = blockIdx.x * SK_N;
  float acc0 = 0.f, acc1 = 0.f;
  for (int k0 = 0; k0 < K; k0 += SK_KC) {
    const int kc = min(SK_KC, K - k0);                                  // K % 4 == 0 -> kc % 4 == 0
    __syncthreads();
    // all loads of the chunk are issued before the first store (8 + 1 (or 4) independent requests per thread in flight: a
    // load -> store loop would pay one L2 round trip per iteration)
    float4 va[SK_M * (SK_KC / 4) / 256];
#pragma unroll
    for (int u = 0; u < SK_M * (SK_KC / 4) / 256; ++u) {               // activation chunk: float4 along k
      const int i = u * 256 + tid, r = i / (SK_KC / 4), k4 = (i - r * (SK_KC / 4)) * 4;
      va[u] = (r < M && k4 < kc) ? __ldg(reinterpret_cast<const float4*>(A + (long)r * lda + k0 + k4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (sk == 1) {                                                      // weight rows contiguous along k: 8 x 32 float4 = 1 per thread
      const int r = tid / (SK_KC / 4), k4 = (tid - r * (SK_KC / 4)) * 4;
      const float4 vw = (n0 + r < N && k4 < kc) ? __ldg(reinterpret_cast<const float4*>(W + (long)(n0 + r) * sn + k0 + k4))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&Ws[r][k4]) = vw;
    } else {                                                            // transposed view: contiguous along n, 4 scalars per thread
      float vw[SK_N * SK_KC / 256];
#pragma unroll
      for (int u = 0; u < SK_N * SK_KC / 256; ++u) {
        const int i = u * 256 + tid, k = i / SK_N, r = i - k * SK_N;
        vw[u] = (n0 + r < N && k < kc) ? __ldg(W + (long)(n0 + r) * sn + (long)(k0 + k) * sk) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < SK_N * SK_KC / 256; ++u) {
        const int i = u * 256 + tid, k = i / SK_N, r = i - k * SK_N;
        Ws[r][k] = vw[u];
      }
    }
#pragma unroll
    for (int u = 0; u < SK_M * (SK_KC / 4) / 256; ++u) {
      const int i = u * 256 + tid, r = i / (SK_KC / 4), k4 = (i - r * (SK_KC / 4)) * 4;
      *reinterpret_cast<float4*>(&As[r][k4]) = va[u];
    }
    __syncthreads();
    const float4* a4 = reinterpret_cast<const float4*>(&As[m][0]);
    const float4* w0 = reinterpret_cast<const float4*>(&Ws[cg * 2][0]);
    const float4* w1 = reinterpret_cast<const float4*>(&Ws[cg * 2 + 1][0]);
#pragma unroll 8
    for (int k4 = 0; k4 < SK_KC / 4; ++k4) {
      const float4 a = a4[k4], x = w0[k4], y = w1[k4];
      acc0 = fmaf(a.x, x.x, acc0); acc0 = fmaf(a.y, x.y, acc0); acc0 = fmaf(a.z, x.z, acc0); acc0 = fmaf(a.w, x.w, acc0);
      acc1 = fmaf(a.x, y.x, acc1); acc1 = fmaf(a.y, y.y, acc1); acc1 = fmaf(a.z, y.z, acc1); acc1 = fmaf(a.w, y.w, acc1);
    }
  }
  if (m < M) {
    const int n = n0 + cg * 2;
    if (n < N) C[(long)m * ldc + n] = acc0 + (bias ? bias[n] : 0.f);
    if (n + 1 < N) C[(long)m * ldc + n + 1] = acc1 + (bias ? bias[n + 1] : 0.f);
  }
}
}  // namespace

extern "C" int slu_skinny_gemm(const float* A, long lda, const float* W, long sn, long sk, const float* bias, float* C, long ldc, int M,
                               int N, int K, void* stream) {
  if (M <= 0 || M > SK_M || N <= 0 || K <= 0 || (K & 3) || (lda & 3) || (reinterpret_cast<uintptr_t>(A) & 15)) return (int)cudaErrorInvalidValue;
  if (sk == 1 && ((sn & 3) || (reinterpret_cast<uintptr_t>(W) & 15))) return (int)cudaErrorInvalidValue;
  skinny_gemm_kernel<<<(N + SK_N - 1) / SK_N, 256, 0, (cudaStream_t)stream>>>(A, lda, W, sn, sk, bias, C, ldc, M, N, K);
  SLU_CHECK_LAUNCH();
  return 0;
}
