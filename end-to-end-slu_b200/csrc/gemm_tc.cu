// Dense "tap-GEMM" on tcgen05 tensor cores (sm_100a), fp32 in / fp32 out, 3-pass bf16 split.
//
//   C[m][n] = sum_{tap < taps} sum_{k < K}  A[(m + tap - tap_pad)][k] * W[n][tap][k]   (+ bias[n], optional LeakyReLU)
//
// Serves every dense contraction with a row-major activation operand on the hot path:
//   * x-projection  gx = x.W_ih^T + b_ih           (taps=1; replaces the GEMM inside nn.GRU, models.py:232/262/686)
//   * input gradients  dX = dgx.W_ih               (taps=1, weights read transposed by the pre-split kernel)
//   * the CNN tail  Conv1d(k=5,pad=2)+bias+LeakyReLU as 5 accumulating taps over the NLC activations
//                   (models.py:200-220) and its dX (taps walked backwards)
// (weight gradients live in wgrad_tc.cu, the SincNet front end in sinc_tc.cu).
// All 8 warps of a CTA stage operands (software pipeline, 2 stages, register prefetch of the next k-block): activations
// are read as fp32, split into bf16 hi + lo in registers and stored K-major (no swizzle, padded leading-byte-offset) in
// shared memory; weights are PRE-SPLIT once per call (slu_presplit_bf16, any strides / tap order) so their tile is a
// plain 16-byte copy.  One elected thread issues hi*hi + hi*lo + lo*hi tcgen05.mma per K=16 step into a [128 x BN] fp32
// accumulator in TMEM (async; it overlaps the staging of the next k-block); the epilogue reads TMEM (tcgen05.ld),
// transposes through shared memory and stores coalesced 16-byte row segments.  Rows shifted out of their utterance
// (conv padding) read as zero.
#include "common.cuh"
#include "tc05.cuh"

namespace {
using namespace tc05;

struct GemmParams {
  const float* A; long lda;          // A[(m + tap - tap_pad) * lda + k]
  const __nv_bfloat16* Wimg;         // pre-split weights: [2 (hi, lo)][taps][N][Kp]
  int Kp;
  const float* bias;                 // [N] or null
  float* C; long ldc;
  int M, N, K;                       // K = reduction length per tap
  int taps, tap_pad;                 // A row shift for tap i = i - tap_pad (rows = frames of one utterance of length T)
  int T;                             // frames per utterance (row boundary for shifted A rows); 0 = none
  int act;                           // 0 none, 1 LeakyReLU(slope)
  float slope;
};

constexpr int BM = 128, BK = 32, STAGES = 2, THREADS = 256;

__host__ __device__ constexpr uint32_t tmem_cols(int bn) { return bn <= 32 ? 32 : (bn <= 64 ? 64 : (bn <= 128 ? 128 : 256)); }

template <int BN>
struct Smem {
  static constexpr uint32_t LBO_A = BM * 16 + 16, LBO_B = BN * 16 + 16;
  static constexpr uint32_t A_PART = (BK / 8) * LBO_A, B_PART = (BK / 8) * LBO_B;      // one of hi / lo
  static constexpr uint32_t STAGE = 2 * A_PART + 2 * B_PART;
  static constexpr uint32_t PIPE = STAGES * STAGE;
  static constexpr uint32_t TRANS = 8 * 32 * 33 * 4;                                    // epilogue transpose buffers
  static constexpr uint32_t TOTAL = PIPE > TRANS ? PIPE : TRANS;
};

// 8 consecutive-k fp32 values of one K-contiguous operand row (two 16-byte loads when possible).
__device__ __forceinline__ void load8_kc(const float* base, int k0, int K, bool row_ok, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (!row_ok || k0 >= K) return;
  const float* p = base + k0;
  if (k0 + 8 <= K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (k0 + i < K) v[i] = __ldg(p + i);
  }
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using S = Smem<BN>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t empty_bar[STAGES], acc_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb_per_tap = (p.K + BK - 1) / BK;
  const int nkb = p.taps * kb_per_tap;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, tmem_cols(BN));
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = idesc_bf16_f32(BM, BN);

  // per-thread chunk assignment (fixed across k-blocks): A 2 chunks, B up to 4 chunks of 8 consecutive k
  constexpr int A_CH = BM * (BK / 8) / THREADS;
  constexpr int B_TOT = BN * (BK / 8);
  constexpr int B_CH = (B_TOT + THREADS - 1) / THREADS;
  int a_r[A_CH], a_kc[A_CH], a_t[A_CH];
#pragma unroll
  for (int u = 0; u < A_CH; ++u) {
    const int c = tid + u * THREADS;
    a_kc[u] = c & 3; a_r[u] = c >> 2;
    a_t[u] = p.T ? (m0 + a_r[u]) % p.T : 0;              // frame of this row inside its utterance (tap boundaries)
  }

  // Register-prefetch pipeline: the global loads of k-block i+1 are issued right after k-block i has been
  // converted into its shared-memory stage, so their latency hides behind the barrier / MMA issue.
  float va[A_CH][8];
  uint4 ib_hi[B_CH], ib_lo[B_CH];
  auto prefetch = [&](int kb) {
    const int tap = kb / kb_per_tap, k0 = (kb % kb_per_tap) * BK;
#pragma unroll
    for (int u = 0; u < A_CH; ++u) {
      const int m = m0 + a_r[u];
      bool ok = m < p.M;
      long row = m;
      if (p.taps > 1 || p.tap_pad) {
        const int sh = tap - p.tap_pad;
        const int t = a_t[u] + sh;
        ok = ok && (p.T == 0 || (t >= 0 && t < p.T));
        row = (long)m + sh;
      }
      load8_kc(p.A + row * p.lda, k0 + a_kc[u] * 8, p.K, ok, va[u]);
    }
#pragma unroll
    for (int u = 0; u < B_CH; ++u) {
      const int c = tid + u * THREADS;
      const int kc = c & 3, r = c >> 2, n = n0 + r;
      ib_hi[u] = make_uint4(0, 0, 0, 0); ib_lo[u] = ib_hi[u];
      if (c < B_TOT && n < p.N) {
        const size_t e = ((size_t)tap * p.N + n) * p.Kp + k0 + kc * 8;
        ib_hi[u] = __ldg(reinterpret_cast<const uint4*>(p.Wimg + e));
        ib_lo[u] = __ldg(reinterpret_cast<const uint4*>(p.Wimg + (size_t)p.taps * p.N * p.Kp + e));
      }
    }
  };
  prefetch(0);

  for (int i = 0; i < nkb; ++i) {
    const int s = i % STAGES;
    if (i >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)(((i / STAGES) - 1) & 1));
    const int k0 = (i % kb_per_tap) * BK;
    uint8_t* st = smem + s * S::STAGE;
    uint8_t* a_hi = st; uint8_t* a_lo = st + S::A_PART;
    uint8_t* b_hi = st + 2 * S::A_PART; uint8_t* b_lo = b_hi + S::B_PART;
#pragma unroll
    for (int u = 0; u < A_CH; ++u) {
      uint4 hi, lo; split8(va[u], hi, lo);
      const uint32_t off = (uint32_t)a_kc[u] * S::LBO_A + (uint32_t)a_r[u] * 16;
      *reinterpret_cast<uint4*>(a_hi + off) = hi;
      *reinterpret_cast<uint4*>(a_lo + off) = lo;
    }
#pragma unroll
    for (int u = 0; u < B_CH; ++u) {
      const int c = tid + u * THREADS;
      if (c < B_TOT) {
        const uint32_t off = (uint32_t)(c & 3) * S::LBO_B + (uint32_t)(c >> 2) * 16;
        *reinterpret_cast<uint4*>(b_hi + off) = ib_hi[u];
        *reinterpret_cast<uint4*>(b_lo + off) = ib_lo[u];
      }
    }
    fence_async_smem();                   // before the prefetch: a proxy fence waits for this thread's outstanding loads
    if (i + 1 < nkb) prefetch(i + 1);
    __syncthreads();
    // ---- one elected thread issues this k-block's MMAs (async: they overlap the staging of the next k-block)
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
        const int nk16 = min(BK / 16, (p.K - k0 + 15) / 16);
        const uint32_t sa = smem_u32(st);
        const uint64_t ah0 = smem_desc(sa, S::LBO_A, 128), al0 = smem_desc(sa + S::A_PART, S::LBO_A, 128);
        const uint64_t bh0 = smem_desc(sa + 2 * S::A_PART, S::LBO_B, 128), bl0 = smem_desc(sa + 2 * S::A_PART + S::B_PART, S::LBO_B, 128);
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
        if (i == nkb - 1) mma_commit(&acc_bar);
      }
      __syncwarp();
    }
  }

  // ================= epilogue: TMEM -> registers -> smem transpose -> coalesced global rows =================
  mbar_wait(&acc_bar, 0);
  fence_after_sync();
  {
    const int q = warp & 3, half = warp >> 2;                      // TMEM lane quarter, column half
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33); // the pipeline buffers are free once acc_bar fired
    constexpr int CW = BN / 2;                                     // columns per warp (32, 64 or 128)
#pragma unroll 1
    for (int cc = 0; cc < CW; cc += 32) {
      const int c0 = half * CW + cc;
      if (n0 + c0 >= p.N) break;
      float v[32];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) tr[lane * 33 + i] = v[i];
      __syncwarp();
      const int rows = min(32, p.M - (m0 + q * 32));
      if (rows == 32 && n0 + c0 + 32 <= p.N && (p.ldc & 3) == 0 && ((n0 + c0) & 3) == 0 &&
          (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
        // fast path: 16-byte stores, a warp instruction writes 4 rows x 128 B
        const int cq = (lane & 7) * 4, r0 = lane >> 3;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + cq));
        float* dst4 = p.C + (long)(m0 + q * 32) * p.ldc + n0 + c0 + cq;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 4 * i;
          float4 x;
          x.x = tr[r * 33 + cq] + b4.x; x.y = tr[r * 33 + cq + 1] + b4.y; x.z = tr[r * 33 + cq + 2] + b4.z; x.w = tr[r * 33 + cq + 3] + b4.w;
          if (p.act == 1) {
            x.x = x.x > 0.f ? x.x : x.x * p.slope; x.y = x.y > 0.f ? x.y : x.y * p.slope;
            x.z = x.z > 0.f ? x.z : x.z * p.slope; x.w = x.w > 0.f ? x.w : x.w * p.slope;
          }
          *reinterpret_cast<float4*>(dst4 + (long)r * p.ldc) = x;
        }
      } else {
        const int n = n0 + c0 + lane;
        if (n < p.N) {
          const float bias = p.bias ? __ldg(p.bias + n) : 0.f;
          float* dst = p.C + (long)(m0 + q * 32) * p.ldc + n;
          for (int r = 0; r < rows; ++r) {
            float x = tr[r * 33 + lane] + bias;
            if (p.act == 1) x = x > 0.f ? x : x * p.slope;
            dst[(long)r * p.ldc] = x;
          }
        }
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols(BN));
}

template <int BN>
int launch(const GemmParams& p, cudaStream_t stream) {
  const size_t smem = Smem<BN>::TOTAL;
  static int attr = slu_set_smem((const void*)gemm_tc_kernel<BN>, smem);
  if (attr) return attr;
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
  gemm_tc_kernel<BN><<<grid, THREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

// fp32 strided weights -> bf16 hi / lo images [2][taps][N][Kp] (zero padded to Kp, a multiple of 32)
__global__ void presplit_kernel(const float* __restrict__ W, long sn, long sk, long stap, int taps, int N, int K, int Kp,
                                int row_len, __nv_bfloat16* __restrict__ img) {
  const size_t total = (size_t)taps * N * Kp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const size_t tn = i / Kp;
    const int n = (int)(tn % N), tap = (int)(tn / N);
    float v = 0.f;
    if (k < K && (row_len == 0 || (long)k * sk + (long)tap * stap < row_len)) v = W[(long)n * sn + (long)k * sk + (long)tap * stap];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    img[i] = h;
    img[total + i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace

// like slu_presplit_bf16, with elements whose in-row offset k*sk + tap*stap reaches row_len read as 0 (used by sinc_tc.cu)
int slu_presplit_rows(const float* W, long sn, long sk, long stap, int taps, int N, int K, int row_len, void* img, void* stream) {
  if (taps <= 0 || N <= 0 || K <= 0) return (int)cudaErrorInvalidValue;
  const int Kp = (K + 31) / 32 * 32;
  const size_t total = (size_t)taps * N * Kp;
  int grid = (int)((total + 255) / 256);
  if (grid > 1184) grid = 1184;
  presplit_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, sn, sk, stap, taps, N, K, Kp, row_len, (__nv_bfloat16*)img);
  return (int)cudaGetLastError();
}

// Pre-split a (strided) fp32 weight operand W(n, tap, k) = W[n*sn + k*sk + tap*stap] into the bf16 hi/lo image the GEMM's
// weight side copies verbatim.  img must hold 2 * taps * N * Kp bf16 values, Kp = K rounded up to a multiple of 32.
extern "C" int slu_presplit_bf16(const float* W, long sn, long sk, long stap, int taps, int N, int K, void* img, void* stream) {
  return slu_presplit_rows(W, sn, sk, stap, taps, N, K, 0, img, stream);
}

// C[m][n] = sum_tap sum_k A[(m + tap - tap_pad)*lda + k] * W(n, tap, k) (+ bias[n]) (LeakyReLU if act == 1); see include/slu_b200.h
extern "C" int slu_gemm_tc(const float* A, long lda, const void* w_img, const float* bias, float* C, long ldc, int M, int N, int K,
                           int taps, int tap_pad, int T, int act, float slope, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || taps <= 0 || !w_img) return (int)cudaErrorInvalidValue;
  GemmParams p;
  p.A = A; p.lda = lda; p.Wimg = (const __nv_bfloat16*)w_img; p.Kp = (K + 31) / 32 * 32; p.bias = bias;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.taps = taps; p.tap_pad = tap_pad; p.T = T; p.act = act; p.slope = slope;
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 64) return launch<64>(p, st);
  if (N <= 128) return launch<128>(p, st);
  return launch<256>(p, st);
}
