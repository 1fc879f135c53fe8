// Self-test of the tcgen05 primitives in tc05.cuh: C[128][N] = A[128][K] . B[N][K]^T with the
// 3-pass bf16 split, one CTA.  Exercised by tests/test_gpu_kernels.py against an fp64 numpy product.
#include "common.cuh"
#include "tc05.cuh"

namespace {
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             float* __restrict__ C, int N, int K) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  using namespace tc05;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kc_n = K / 8;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + 128 * K * 2;
  uint8_t* b_hi = a_lo + 128 * K * 2;
  uint8_t* b_lo = b_hi + N * K * 2;
  for (int idx = tid; idx < 128 * kc_n; idx += 128) {
    const int r = idx % 128, kc = idx / 128;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = A[(size_t)r * K + kc * 8 + i];
    uint4 hi, lo; split8(v, hi, lo);
    *reinterpret_cast<uint4*>(a_hi + chunk_off(128, r, kc)) = hi;
    *reinterpret_cast<uint4*>(a_lo + chunk_off(128, r, kc)) = lo;
  }
  for (int idx = tid; idx < N * kc_n; idx += 128) {
    const int r = idx % N, kc = idx / N;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = Bm[(size_t)r * K + kc * 8 + i];
    uint4 hi, lo; split8(v, hi, lo);
    *reinterpret_cast<uint4*>(b_hi + chunk_off(N, r, kc)) = hi;
    *reinterpret_cast<uint4*>(b_lo + chunk_off(N, r, kc)) = lo;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    mma_split3(tmem, smem_u32(a_hi), smem_u32(a_lo), 128, smem_u32(b_hi), smem_u32(b_lo), N, K / 16, idesc_bf16_f32(128, N), false);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) C[(size_t)tid * N + n0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Same product with the A operand resident in TMEM (weights-stationary form used by the GRU kernels);
// B tile uses a padded leading-byte-offset like the GRU's h tile.
__global__ void __launch_bounds__(128, 1) tc_selftest_ts_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                float* __restrict__ C, int N, int K) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  using namespace tc05;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kc_n = K / 8;
  const uint32_t lbo = N * 16 + 16;
  uint8_t* b_hi = smem;
  uint8_t* b_lo = b_hi + kc_n * lbo;
  for (int idx = tid; idx < N * kc_n; idx += 128) {
    const int r = idx % N, kc = idx / N;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = Bm[(size_t)r * K + kc * 8 + i];
    uint4 hi, lo; split8(v, hi, lo);
    *reinterpret_cast<uint4*>(b_hi + kc * lbo + r * 16) = hi;
    *reinterpret_cast<uint4*>(b_lo + kc * lbo + r * 16) = lo;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  {
    float row[128];
    for (int k = 0; k < K; ++k) row[k] = A[(size_t)tid * K + k];
    tmem_store_row_split(tmem + lane_base + 256, tmem + lane_base + 256 + K / 2, row, K);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mma_split3_ts(tmem, tmem + 256, tmem + 256 + K / 2, smem_u32(b_hi), smem_u32(b_lo), lbo, K / 16, idesc_bf16_f32(128, N), false);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 8) {
    float v[8];
    tmem_ld8(tmem + lane_base + n0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) C[(size_t)tid * N + n0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int slu_tc_selftest_ts(const float* A, const float* B, float* C, int N, int K, void* stream) {
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 128) return (int)cudaErrorInvalidValue;
  const size_t smem = (size_t)2 * (K / 8) * (N * 16 + 16);
  int e = slu_set_smem((const void*)tc_selftest_ts_kernel, smem);
  if (e) return e;
  tc_selftest_ts_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, N, K);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_tc_selftest(const float* A, const float* B, float* C, int N, int K, void* stream) {
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 256) return (int)cudaErrorInvalidValue;
  const size_t smem = (size_t)(128 + N) * K * 4;
  if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
  int e = slu_set_smem((const void*)tc_selftest_kernel, smem);
  if (e) return e;
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, N, K);
  SLU_CHECK_LAUNCH();
  return 0;
}
