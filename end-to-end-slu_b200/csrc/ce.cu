// Frame-wise cross-entropy of the ASR heads (reference models.py:308-314, 321-329:
// `logits = linear(out).view(B*T', V)`, `F.cross_entropy(logits, y, ignore_index=-1)`, masked arg-max accuracy).
//
// The [B*T', V] logits (V = 10 000 words: 3.76 MB per 15 s utterance) are never materialised as a whole.  The host side
// (ops.LinearCE) walks the frames in chunks of a few thousand rows; per chunk the tcgen05 tap-GEMM (gemm_tc.cu) writes the
// logits tile, `ce_rows_kernel` below turns it IN PLACE into dL/dlogits (and emits the per-row loss / hit flags), and the same
// tile feeds the input-gradient GEMM and the weight-gradient GEMM (wgrad_tc.cu) before the next chunk overwrites it.
//   ce_count_kernel   n_valid = #(y != -1) -> {n_valid, 1/n_valid}   (the mean's divisor is needed before the first chunk)
//   ce_rows_kernel    one row per CTA (V > 256) or per warp: the row lives in registers -- one read, one write of the tile
//   ce_finish_kernel  fixed-order reduction of the per-row losses / hits -> {loss, accuracy}   (bit-reproducible)
//   colsum_acc_kernel out[c] += sum_r A[r][c]                         (bias gradient of a chunk)
//   scale_kernel      dst[i] = src[i] * *g                            (backward: the stashed gradients times dL/dloss)
#include <math.h>

#include "common.cuh"

namespace {

constexpr float kNaN = __builtin_nanf("");

__global__ void ce_count_kernel(const long long* __restrict__ y, long M, float* __restrict__ nv) {
  __shared__ int red[256];
  int c = 0;
  for (long i = threadIdx.x; i < M; i += 256) c += y[i] != -1;
  red[threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) { nv[0] = (float)red[0]; nv[1] = 1.f / (float)red[0]; }
}

// TPR threads per row, NV values per thread (V <= TPR * NV).  logits row stride ld.
template <int TPR, int NV>
__global__ void __launch_bounds__(256) ce_rows_kernel(float* __restrict__ logits, long ld, int V, const long long* __restrict__ y,
                                                      long R, const float* __restrict__ nv, int write_grad,
                                                      float* __restrict__ row_loss, float* __restrict__ row_ok) {
  constexpr int RPB = 256 / TPR;                         // rows per CTA
  const int lane = threadIdx.x % TPR;
  const long row = (long)blockIdx.x * RPB + threadIdx.x / TPR;
  __shared__ float s_val[256 / 32];
  __shared__ int s_idx[256 / 32];
  const bool row_on = row < R;
  float* p = logits + (row_on ? row : 0) * ld;
  float v[NV];
  float m = -INFINITY;
  int am = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * TPR + lane;
    v[i] = (row_on && c < V) ? p[c] : -INFINITY;
    if (v[i] > m) { m = v[i]; am = c; }                  // ascending c within a thread: the first maximum wins
  }
  // ---- row max / arg-max (lowest index among equal maxima)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, m, o);
    const int ao = __shfl_xor_sync(0xffffffffu, am, o);
    if (mo > m || (mo == m && ao < am)) { m = mo; am = ao; }
  }
  if (TPR > 32) {
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_val[w] = m; s_idx[w] = am; }
    __syncthreads();
    m = s_val[0]; am = s_idx[0];
    for (int k = 1; k < TPR / 32; ++k)
      if (s_val[k] > m || (s_val[k] == m && s_idx[k] < am)) { m = s_val[k]; am = s_idx[k]; }
    __syncthreads();
  }
  // ---- sum of exponentials
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += __expf(v[i] - m);    // exp(-inf) = 0 for the padding
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (TPR > 32) {
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s_val[w] = s;
    __syncthreads();
    s = 0.f;
    for (int k = 0; k < TPR / 32; ++k) s += s_val[k];
  }
  if (!row_on) return;
  const float lse = m + logf(s);
  const long long t = y[row];
  const bool valid = t != -1;                            // ignore_index = -1 (data.py:506-507 pads the label rows with it)
  const bool bad = valid && (t < 0 || t >= V);           // F.cross_entropy would device-assert: poison the loss instead
  if (lane == 0) {
    row_loss[row] = bad ? kNaN : (valid ? lse - p[t] : 0.f);
    row_ok[row] = (valid && !bad && am == (int)t) ? 1.f : 0.f;
  }
  if (write_grad) {
    if (TPR > 32) __syncthreads();                       // p[t] above is read before any thread overwrites the row
    else __syncwarp();
    const float scale = (valid && !bad) ? nv[1] : 0.f;   // mean over the valid rows
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * TPR + lane;
      if (c < V) p[c] = scale * (__expf(v[i] - lse) - (c == (int)t ? 1.f : 0.f));
    }
  }
}

__global__ void ce_finish_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_ok, long M,
                                 const float* __restrict__ nv, float* __restrict__ out) {
  __shared__ float red[2][256];
  float l = 0.f, a = 0.f;
  for (long i = threadIdx.x; i < M; i += 256) { l += row_loss[i]; a += row_ok[i]; }
  red[0][threadIdx.x] = l; red[1][threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = red[0][0] * nv[1]; out[1] = red[1][0] * nv[1]; }
}

// out[c] += sum_r A[r*ld + c]: a CTA owns 128 columns (coalesced rows) and a strided subset of the rows.
__global__ void __launch_bounds__(256) colsum_acc_kernel(const float* __restrict__ A, long ld, long R, int C, float* __restrict__ out) {
  __shared__ float red[2][128];
  const int c = blockIdx.x * 128 + (threadIdx.x & 127), half = threadIdx.x >> 7;
  float s = 0.f;
  if (c < C)
    for (long r = (long)blockIdx.y * 2 + half; r < R; r += 2 * gridDim.y) s += A[r * ld + c];
  red[half][threadIdx.x & 127] = s;
  __syncthreads();
  if (half == 0 && c < C) atomicAdd(out + c, red[0][threadIdx.x] + red[1][threadIdx.x]);
}

__global__ void scale_kernel(const float* __restrict__ src, float* __restrict__ dst, long n, const float* __restrict__ g) {
  const float s = g[0];
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = src[i] * s;
}

}  // namespace

extern "C" int slu_ce_count(const long long* y, long M, float* nvalid, void* stream) {
  if (M <= 0) return (int)cudaErrorInvalidValue;
  ce_count_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(y, M, nvalid);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_ce_rows(float* logits, long ld, int V, const long long* y, long R, const float* nvalid, int write_grad,
                           float* row_loss, float* row_ok, void* stream) {
  if (R <= 0 || V <= 0 || ld < V) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  if (V <= 256) ce_rows_kernel<32, 8><<<(unsigned)((R + 7) / 8), 256, 0, st>>>(logits, ld, V, y, R, nvalid, write_grad, row_loss, row_ok);
  else if (V <= 256 * 16) ce_rows_kernel<256, 16><<<(unsigned)R, 256, 0, st>>>(logits, ld, V, y, R, nvalid, write_grad, row_loss, row_ok);
  else if (V <= 256 * 48) ce_rows_kernel<256, 48><<<(unsigned)R, 256, 0, st>>>(logits, ld, V, y, R, nvalid, write_grad, row_loss, row_ok);
  else return SLU_ERR_TOO_LARGE;
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_ce_finish(const float* row_loss, const float* row_ok, long M, const float* nvalid, float* loss_acc, void* stream) {
  if (M <= 0) return (int)cudaErrorInvalidValue;
  ce_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(row_loss, row_ok, M, nvalid, loss_acc);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_colsum_acc(const float* A, long ld, long R, int C, float* out, void* stream) {
  if (R <= 0 || C <= 0) return (int)cudaErrorInvalidValue;
  int gy = (int)((R + 63) / 64);
  if (gy > 64) gy = 64;
  colsum_acc_kernel<<<dim3((C + 127) / 128, gy), 256, 0, (cudaStream_t)stream>>>(A, ld, R, C, out);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_scale(const float* src, float* dst, long n, const float* g, void* stream) {
  if (n <= 0) return n == 0 ? 0 : (int)cudaErrorInvalidValue;
  long blocks = (n + 1023) / 1024;
  if (blocks > 1184) blocks = 1184;
  scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n, g);
  SLU_CHECK_LAUNCH();
  return 0;
}
