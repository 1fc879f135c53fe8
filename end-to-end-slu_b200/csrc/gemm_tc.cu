// Dense "tap-GEMM" on tcgen05 tensor cores (sm_100a), fp32 in / fp32 out, 3-pass bf16 split.
//
//   C[m][n] (+)= sum_{tap < taps} sum_{k < K}  A(m, tap, k) * Bop(n, tap, k)   (+ bias[n], optional LeakyReLU)
//
// One kernel family serves every dense contraction on the hot path that is not the recurrence itself:
//   * x-projection  gx = x.W_ih^T + b_ih           (taps=1; replaces the GEMM inside nn.GRU, models.py:232/262/686)
//   * its backward  dX = dgx.W_ih, dW_ih = dgx^T.x, dW_hh = dgh^T.h_prev   (split-K, fp32 atomics)
//   * the CNN tail  Conv1d(k=5,pad=2)+bias+LeakyReLU as 5 accumulating taps over the NLC activations
//                   (models.py:200-220) and its dX / dW
// Operands are read as fp32 from global memory by 4 loader warps, split into bf16 hi + lo in registers and
// staged K-major (no swizzle, padded leading-byte-offset) in shared memory; one thread issues
// hi*hi + hi*lo + lo*hi tcgen05.mma per K=16 step into a [128 x BN] fp32 accumulator in TMEM; 4 epilogue warps
// read it back (tcgen05.ld), transpose through shared memory and store coalesced rows (or red.add for split-K).
// Element addressing is fully strided so that transposed operands (weight gradients) need no transpose pass:
//   A(m,tap,k)   = A[(m + a_row_shift(tap)) * a_sm + k * a_sk]     rows outside the utterance -> 0 (conv padding)
//   Bop(n,tap,k) = B[n * b_sn + k * b_sk + tap * b_stap]
// For the weight-gradient form the reduction index k is the flattened (utterance, frame) row and either operand may
// be time-shifted by `*_kshift` frames inside its utterance (h_{t-1} for dW_hh, x_{t+tap-2} for conv dW).
#include "common.cuh"
#include "tc05.cuh"

namespace {
using namespace tc05;

struct GemmParams {
  const float* A; long a_sm, a_sk;
  const float* B; long b_sn, b_sk, b_stap;
  const float* bias;          // [N] or null
  float* C; long ldc;
  int M, N, K;                // K = reduction length per tap
  int taps, tap_pad;          // A row shift for tap i = i - tap_pad (rows = frames of one utterance of length T)
  int T;                      // frames per utterance (row-boundary for shifted A rows, period for k-shifts); 0 = none
  int a_kshift, b_kshift;     // weight-gradient form: operand(k) taken at frame t + shift (0 outside the utterance)
  int split_k;                // gridDim.z; >1 => atomic accumulation into a zeroed C
  int act;                    // 0 none, 1 LeakyReLU(slope)
  float slope;
};

constexpr int BM = 128, BK = 32, STAGES = 2;
constexpr int LOADERS = 128, THREADS = 256;

template <int BN>
struct Smem {
  static constexpr uint32_t LBO_A = BM * 16 + 16, LBO_B = BN * 16 + 16;
  static constexpr uint32_t A_PART = (BK / 8) * LBO_A, B_PART = (BK / 8) * LBO_B;      // one of hi / lo
  static constexpr uint32_t STAGE = 2 * A_PART + 2 * B_PART;
  static constexpr uint32_t TOTAL = STAGES * STAGE;
};

// 8 consecutive-k values of one operand row.  KC = K-contiguous fast path (two 16-byte loads when possible).
template <bool KCONTIG>
__device__ __forceinline__ void load8(const float* base, long sk, int k0, int K, bool row_ok, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (!row_ok || k0 >= K) return;
  if (KCONTIG) {
    const float* p = base + k0;
    if (k0 + 8 <= K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (k0 + i < K) v[i] = __ldg(p + i);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (k0 + i < K) v[i] = __ldg(base + (long)(k0 + i) * sk);
  }
}

// Weight-gradient form (reduction over frames): 8 consecutive frames k0..k0+7 of column `col`, each frame optionally
// shifted inside its utterance.
__device__ __forceinline__ void load8_frames(const float* base, long s_frame, int k0, int K, int T, int shift, bool col_ok, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = 0.f;
    const int k = k0 + i;
    if (col_ok && k < K) {
      if (shift == 0) v[i] = __ldg(base + (long)k * s_frame);
      else {
        const int t = k % T + shift;
        if (t >= 0 && t < T) v[i] = __ldg(base + (long)(k + shift) * s_frame);
      }
    }
  }
}

template <int BN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using S = Smem<BN>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // K range of this CTA (split-K over k-blocks of the flattened (tap, k) loop)
  const int kb_per_tap = (p.K + BK - 1) / BK;
  const int kb_total = p.taps * kb_per_tap;
  const int kb_chunk = (kb_total + p.split_k - 1) / p.split_k;
  const int kb_begin = blockIdx.z * kb_chunk;
  const int kb_end = min(kb_total, kb_begin + kb_chunk);
  const int nkb = max(0, kb_end - kb_begin);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], LOADERS); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 4) tmem_alloc(&tmem_base, BN < 32 ? 32 : BN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;

  if (warp < 4) {
    // ================= loaders: global fp32 -> bf16 hi/lo K-major tiles =================
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)(((i / STAGES) - 1) & 1));
      const int kb = kb_begin + i;
      const int tap = kb / kb_per_tap, k0 = (kb % kb_per_tap) * BK;
      uint8_t* st = smem + s * S::STAGE;
      uint8_t* a_hi = st; uint8_t* a_lo = st + S::A_PART;
      uint8_t* b_hi = st + 2 * S::A_PART; uint8_t* b_lo = b_hi + S::B_PART;
      // ---- A tile: BM rows x 4 k-chunks
      for (int c = tid; c < BM * (BK / 8); c += LOADERS) {
        int r, kc;
        if (A_KC) { kc = c & 3; r = c >> 2; } else { r = c % BM; kc = c / BM; }
        const int m = m0 + r;
        float v[8];
        if (A_KC) {
          bool ok = m < p.M;
          long row = m;
          if (p.taps > 1 || p.tap_pad) {
            const int sh = tap - p.tap_pad;
            const int t = p.T ? (m % p.T) + sh : 0;
            ok = ok && (p.T == 0 || (t >= 0 && t < p.T));
            row = (long)m + sh;
          }
          load8<true>(p.A + row * p.a_sm, 1, k0 + kc * 8, p.K, ok, v);
        } else {
          load8_frames(p.A + (long)m * p.a_sm, p.a_sk, k0 + kc * 8, p.K, p.T, p.a_kshift + (p.taps > 1 ? tap - p.tap_pad : 0), m < p.M, v);
        }
        uint4 hi, lo; split8(v, hi, lo);
        const uint32_t off = (uint32_t)kc * S::LBO_A + (uint32_t)r * 16;
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        *reinterpret_cast<uint4*>(a_lo + off) = lo;
      }
      // ---- B tile: BN rows x 4 k-chunks
      for (int c = tid; c < BN * (BK / 8); c += LOADERS) {
        int r, kc;
        if (B_KC) { kc = c & 3; r = c >> 2; } else { r = c % BN; kc = c / BN; }
        const int n = n0 + r;
        float v[8];
        const float* base = p.B + (long)n * p.b_sn + (long)tap * p.b_stap;
        if (B_KC) load8<true>(base, 1, k0 + kc * 8, p.K, n < p.N, v);
        else load8_frames(base, p.b_sk, k0 + kc * 8, p.K, p.T ? p.T : 1, p.b_kshift, n < p.N, v);
        uint4 hi, lo; split8(v, hi, lo);
        const uint32_t off = (uint32_t)kc * S::LBO_B + (uint32_t)r * 16;
        *reinterpret_cast<uint4*>(b_hi + off) = hi;
        *reinterpret_cast<uint4*>(b_lo + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(&full_bar[s]);
    }
  } else {
    // ================= MMA issue (one thread) =================
    if (warp == 4 && elect_one()) {
      const uint32_t idesc = idesc_bf16_f32(BM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbar_wait(&full_bar[s], (uint32_t)((i / STAGES) & 1));
        fence_after_sync();
        const int kb = kb_begin + i;
        const int k0 = (kb % kb_per_tap) * BK;
        const int nk16 = min(BK / 16, (p.K - k0 + 15) / 16);
        const uint32_t st = smem_u32(smem + s * S::STAGE);
        const uint64_t ah0 = smem_desc(st, S::LBO_A, 128), al0 = smem_desc(st + S::A_PART, S::LBO_A, 128);
        const uint64_t bh0 = smem_desc(st + 2 * S::A_PART, S::LBO_B, 128), bl0 = smem_desc(st + 2 * S::A_PART + S::B_PART, S::LBO_B, 128);
        uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          if (kk < nk16) {
            const uint64_t ah = desc_advance(ah0, kk * 2 * S::LBO_A), al = desc_advance(al0, kk * 2 * S::LBO_A);
            const uint64_t bh = desc_advance(bh0, kk * 2 * S::LBO_B), bl = desc_advance(bl0, kk * 2 * S::LBO_B);
            mma_bf16(tmem, ah, bh, idesc, acc); acc = 1u;
            mma_bf16(tmem, ah, bl, idesc, 1u);
            mma_bf16(tmem, al, bh, idesc, 1u);
          }
        }
        mma_commit(&empty_bar[s]);
      }
      mma_commit(&acc_bar);
    }
    __syncwarp();
    // ================= epilogue: TMEM -> registers -> smem transpose -> coalesced global rows =================
    if (nkb > 0) {
      mbar_wait(&acc_bar, 0);
      fence_after_sync();
      const int q = warp & 3;                                   // TMEM lane quarter of this warp
      float* tr = reinterpret_cast<float*>(smem) + q * (32 * 33);   // stage buffers are free once acc_bar fired
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;
        float v[32];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0 + 16, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) tr[lane * 33 + i] = v[i];
        __syncwarp();
        const int n = n0 + c0 + lane;
        const float bias = (p.bias && n < p.N && blockIdx.z == 0) ? __ldg(p.bias + n) : 0.f;
        for (int r = 0; r < 32; ++r) {
          const int m = m0 + q * 32 + r;
          if (m < p.M && n < p.N) {
            float x = tr[r * 33 + lane] + bias;
            if (p.act == 1) x = x > 0.f ? x : x * p.slope;
            float* dst = p.C + (long)m * p.ldc + n;
            if (p.split_k > 1) atomicAdd(dst, x); else *dst = x;
          }
        }
        __syncwarp();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, BN < 32 ? 32 : BN);
}

template <int BN, bool A_KC, bool B_KC>
int launch(const GemmParams& p, cudaStream_t stream) {
  const size_t smem = Smem<BN>::TOTAL;
  static int attr = slu_set_smem((const void*)gemm_tc_kernel<BN, A_KC, B_KC>, smem);
  if (attr) return attr;
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, p.split_k);
  gemm_tc_kernel<BN, A_KC, B_KC><<<grid, THREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <bool A_KC, bool B_KC>
int dispatch_bn(const GemmParams& p, cudaStream_t stream) {
  if (p.N <= 64) return launch<64, A_KC, B_KC>(p, stream);
  if (p.N <= 128) return launch<128, A_KC, B_KC>(p, stream);
  return launch<256, A_KC, B_KC>(p, stream);
}

}  // namespace

// Generic entry point (see include/slu_b200.h).  a_kc / b_kc: operand is K-contiguous (stride a_sk / b_sk == 1).
extern "C" int slu_gemm_tc(const float* A, long a_sm, long a_sk, const float* B, long b_sn, long b_sk, long b_stap,
                           const float* bias, float* C, long ldc, int M, int N, int K, int taps, int tap_pad, int T,
                           int a_kshift, int b_kshift, int split_k, int act, float slope, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || taps <= 0 || split_k <= 0) return (int)cudaErrorInvalidValue;
  GemmParams p;
  p.A = A; p.a_sm = a_sm; p.a_sk = a_sk; p.B = B; p.b_sn = b_sn; p.b_sk = b_sk; p.b_stap = b_stap; p.bias = bias;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.taps = taps; p.tap_pad = tap_pad; p.T = T;
  p.a_kshift = a_kshift; p.b_kshift = b_kshift; p.split_k = split_k; p.act = act; p.slope = slope;
  const bool a_kc = (a_sk == 1), b_kc = (b_sk == 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (a_kc && b_kc) return dispatch_bn<true, true>(p, st);
  if (a_kc && !b_kc) return dispatch_bn<true, false>(p, st);
  if (!a_kc && b_kc) return dispatch_bn<false, true>(p, st);
  return dispatch_bn<false, false>(p, st);
}
