// Dense "tap-GEMM" on tcgen05 tensor cores (sm_100a), fp32 in / fp32 out, 3-pass bf16 split.
//
//   C[m][n] (+)= sum_{tap < taps} sum_{k < K}  A(m, tap, k) * Bop(n, tap, k)   (+ bias[n], optional LeakyReLU)
//
// One kernel family serves every dense contraction on the hot path that is not the recurrence itself:
//   * x-projection  gx = x.W_ih^T + b_ih           (taps=1; replaces the GEMM inside nn.GRU, models.py:232/262/686)
//   * its backward  dX = dgx.W_ih, dW_ih = dgx^T.x, dW_hh = dgh^T.h_prev   (split-K, fp32 atomics)
//   * the CNN tail  Conv1d(k=5,pad=2)+bias+LeakyReLU as 5 accumulating taps over the NLC activations
//                   (models.py:200-220) and its dX / dW
// All 8 warps of a CTA stage operands (software pipeline, 2 stages): activations are read as fp32, split into
// bf16 hi + lo in registers and stored K-major (no swizzle, padded leading-byte-offset) in shared memory; weights can
// be PRE-SPLIT once per call (slu_presplit_bf16) so their tile is a plain 16-byte copy.  One elected thread issues
// hi*hi + hi*lo + lo*hi tcgen05.mma per K=16 step into a [128 x BN] fp32 accumulator in TMEM (async; it overlaps
// the staging of the next k-block); the epilogue reads TMEM (tcgen05.ld), transposes through shared memory and
// stores coalesced rows (or red.add for split-K).  Element addressing is fully strided, so transposed operands
// (weight gradients) need no transpose pass:
//   A(m,tap,k)   = A[(m + tap - tap_pad) * a_sm + k]              a_sk == 1; rows leaving their utterance read 0
//                = A[m * a_sm + frame(k + a_kshift) * a_sk]        a_sk != 1: reduction over frames (weight gradients)
//   Bop(n,tap,k) = B[n * b_sn + k * b_sk + tap * b_stap]           (b_sk != 1: frame-shifted by b_kshift like A)
//                = pre-split image  hi/lo[(tap * N + n) * Kp + k]  (b_img given)
#include "common.cuh"
#include "tc05.cuh"

namespace {
using namespace tc05;

struct GemmParams {
  const float* A; long a_sm, a_sk;
  const float* B; long b_sn, b_sk, b_stap;
  const __nv_bfloat16* Bimg;  // pre-split weights: [2][taps][N][Kp] (hi then lo), or null
  int Kp;
  const float* bias;          // [N] or null
  float* C; long ldc;
  int M, N, K;                // K = reduction length per tap
  int taps, tap_pad;          // A row shift for tap i = i - tap_pad (rows = frames of one utterance of length T)
  int T;                      // frames per utterance (row-boundary for shifted A rows, period for k-shifts); 0 = none
  int a_kshift, b_kshift;     // weight-gradient form: operand(k) taken at frame t + shift (0 outside the utterance)
  int split_k;                // gridDim.z; >1 => atomic accumulation into a zeroed C
  int act;                    // 0 none, 1 LeakyReLU(slope)
  float slope;
  // SincNet front end (AMODE 2/3, BMODE 3, EPI 1): rows / reduction indices are frames m = b*L0p + t of the waveform x[B][Ts]
  int Ts, L0, L0p, L1;        // samples per utterance, conv frames, L0 rounded up to even, pooled frames
  const float* gy;            // [B][L1][80] gradient of the pooled output (AMODE 3)
  uint8_t* route;             // [B][L1][80] pooling route bits (written by EPI 1, read by AMODE 3)
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

// Strided form: 8 consecutive reduction indices k0..k0+7 at stride s_k; with shift != 0 the index is a frame number
// inside utterances of T frames and is shifted (0 outside the utterance).
__device__ __forceinline__ void load8_strided(const float* base, long s_k, int k0, int K, int T, int shift, bool ok, float* v) {
  if (shift == 0) {
    const float* p = base + (long)k0 * s_k;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (ok && k0 + i < K) ? __ldg(p + (long)i * s_k) : 0.f;
  } else {
    int t = k0 % T + shift;                 // one modulo per chunk; frames are consecutive, wrap at the utterance end
    const float* p = base + (long)(k0 + shift) * s_k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = (ok && k0 + i < K && t >= 0 && t < T) ? __ldg(p + (long)i * s_k) : 0.f;
      ++t;
      if (t - shift >= T) t -= T;
    }
  }
}

// Strided ("frames") form, vectorised: a 4 (consecutive m) x 8 (consecutive reduction frames) block with one 16-byte load per
// frame -- 4x fewer load instructions than 8 scalar loads per row.  Needs m-contiguous rows (stride 1) and 16-byte alignment.
__device__ __forceinline__ void load_frames_4x8(const float* base, long s_frame, int m, int M, int k0, int K, int T, int shift,
                                                bool on, float (*v)[8]) {
  const bool vec = on && (m + 4 <= M) && (((reinterpret_cast<uintptr_t>(base + m) | (uintptr_t)(s_frame * 4)) & 15) == 0);
  int t = (shift != 0) ? k0 % T + shift : 0;
  const float* p = base + (long)(k0 + shift) * s_frame + m;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool f_ok = on && (k0 + i < K) && (shift == 0 || (t >= 0 && t < T));
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec) {
      if (f_ok) q = __ldg(reinterpret_cast<const float4*>(p + (long)i * s_frame));
    } else if (f_ok) {
      const float* pi = p + (long)i * s_frame;
      if (m < M) q.x = __ldg(pi);
      if (m + 1 < M) q.y = __ldg(pi + 1);
      if (m + 2 < M) q.z = __ldg(pi + 2);
      if (m + 3 < M) q.w = __ldg(pi + 3);
    }
    v[0][i] = q.x; v[1][i] = q.y; v[2][i] = q.z; v[3][i] = q.w;
    if (shift != 0) { ++t; if (t - shift >= T) t -= T; }
  }
}

// SincConv forward A operand: A(m=(b,t), tap, k) = x[b][80*(t+tap) + k - 200]  (zero outside [0,Ts), for pad frames t>=L0
// and for k >= 80).  The waveform itself is the [frames][80] matrix: no im2col.
__device__ __forceinline__ void load8_sinc_rows(const float* xb, int idx0, int k0, int Ts, bool row_ok, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (!row_ok || k0 >= SLU_STRIDE) return;
  const float* p = xb + idx0;
  if (idx0 >= 0 && idx0 + 8 <= Ts && k0 + 8 <= SLU_STRIDE && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (k0 + i < SLU_STRIDE && idx0 + i >= 0 && idx0 + i < Ts) v[i] = __ldg(p + i);
  }
}

// AMODE: 0 = fp32 K-contiguous rows (+ row-shift taps), 1 = fp32 strided (frames), 2 = sinc frames of the waveform,
//        3 = routed pooled-gradient (sinc backward);  BMODE: 0 = fp32 K-contiguous, 1 = fp32 strided, 2 = pre-split bf16
//        image, 3 = waveform samples per frame (sinc backward);  EPI: 0 = bias/act/store or split-K atomics, 1 = abs+maxpool2+route
template <int BN, int AMODE, int BMODE, int EPI>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  constexpr bool A_KC = (AMODE == 0);
  using S = Smem<BN>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t empty_bar[STAGES], acc_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // K range of this CTA (split-K over k-blocks of the flattened (tap, k) loop)
  const int kb_per_tap = (p.K + BK - 1) / BK;
  const int kb_total = p.taps * kb_per_tap;
  const int kb_chunk = (kb_total + p.split_k - 1) / p.split_k;
  const int kb_begin = blockIdx.z * kb_chunk;
  const int nkb = max(0, min(kb_total, kb_begin + kb_chunk) - kb_begin);
  if (nkb == 0) return;                                  // uniform per CTA

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

  // per-thread A chunk assignment (2 chunks): fixed across k-blocks
  constexpr int A_CH = BM * (BK / 8) / THREADS;            // 2
  constexpr int B_TOT = BN * (BK / 8);
  constexpr int B_CH = (B_TOT + THREADS - 1) / THREADS;
  int a_r[A_CH], a_kc[A_CH], a_t[A_CH];
#pragma unroll
  for (int u = 0; u < A_CH; ++u) {
    const int c = tid + u * THREADS;
    if (AMODE == 0 || AMODE == 2) { a_kc[u] = c & 3; a_r[u] = c >> 2; } else { a_r[u] = c % BM; a_kc[u] = c / BM; }
    a_t[u] = (A_KC && p.T) ? (m0 + a_r[u]) % p.T : 0;      // frame of this row inside its utterance (tap boundaries)
    if (AMODE == 2) a_t[u] = (m0 + a_r[u]) % p.L0p;        // sinc: frame t; the utterance is (m0 + r) / L0p
  }

  // Register-prefetch pipeline: the global loads of k-block i+1 are issued right after k-block i has been
  // converted into its shared-memory stage, so their latency hides behind the fence / barrier / MMA issue.
  // frames-form operands (AMODE 1 / BMODE 1) are staged as 4x8 blocks: A blocks by threads 128..255, B blocks by threads < BN
  constexpr bool A_FR = (AMODE == 1), B_FR = (BMODE == 1);
  const bool fa_on = A_FR && tid >= 128, fb_on = B_FR && tid < BN;
  const int fa_m4 = (tid - 128) & 31, fa_kc = (tid - 128) >> 5, fb_m4 = tid % (BN / 4), fb_kc = tid / (BN / 4);
  // when both operands are frames-form (weight gradients, BN <= 128) a thread owns exactly one block: one register array
  constexpr bool FR_BOTH = A_FR && B_FR;
  static_assert(!FR_BOTH || BN <= 128, "frames-form weight-gradient tiles use BN <= 128");
  float fa[A_FR ? 4 : 1][8];
  float fb_own[(B_FR && !FR_BOTH) ? 4 : 1][8];
  float (*fb)[8] = FR_BOTH ? fa : fb_own;
  float va[A_FR ? 1 : A_CH][8];
  float vb[(BMODE == 2 || B_FR) ? 1 : B_CH][8];
  uint4 ib_hi[(BMODE == 2) ? B_CH : 1], ib_lo[(BMODE == 2) ? B_CH : 1];
  auto prefetch = [&](int i) {
    const int kb = kb_begin + i;
    const int tap = kb / kb_per_tap, k0 = (kb % kb_per_tap) * BK;
    if (A_FR && fa_on) load_frames_4x8(p.A, p.a_sk, m0 + fa_m4 * 4, p.M, k0 + fa_kc * 8, p.K, p.T ? p.T : 1, p.a_kshift, fa_on, fa);
    if (B_FR && fb_on) load_frames_4x8(p.B + (long)tap * p.b_stap, p.b_sk, n0 + fb_m4 * 4, p.N, k0 + fb_kc * 8, p.K, p.T ? p.T : 1, p.b_kshift, fb_on, fb);
#pragma unroll
    for (int u = 0; u < (A_FR ? 0 : A_CH); ++u) {
      const int m = m0 + a_r[u];
      if (AMODE == 0) {
        bool ok = m < p.M;
        long row = m;
        if (p.taps > 1 || p.tap_pad) {
          const int sh = tap - p.tap_pad;
          const int t = a_t[u] + sh;
          ok = ok && (p.T == 0 || (t >= 0 && t < p.T));
          row = (long)m + sh;
        }
        load8_kc(p.A + row * p.a_sm, k0 + a_kc[u] * 8, p.K, ok, va[u]);
      } else if (AMODE == 1) {
        load8_strided(p.A + (long)m * p.a_sm, p.a_sk, k0 + a_kc[u] * 8, p.K, p.T ? p.T : 1, p.a_kshift, m < p.M, va[u]);
      } else if (AMODE == 2) {
        const int t = a_t[u], b = (m - t) / p.L0p;
        load8_sinc_rows(p.A + (long)b * p.Ts, SLU_STRIDE * (t + tap) + k0 + a_kc[u] * 8 - SLU_PAD, k0 + a_kc[u] * 8, p.Ts,
                        m < p.M && t < p.L0, va[u]);
      } else {   // AMODE 3: A(m = filter c, k = frame (b,t)) = pooled-output gradient routed back through max-pool and abs
        const int kf = k0 + a_kc[u] * 8;
        int b = kf / p.L0p, t = kf - b * p.L0p;
        uint8_t rb[8]; float gv[8]; bool okf[8]; int par[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {      // all 16 loads are issued unconditionally (clamped address), selected afterwards
          okf[i] = m < p.M && kf + i < p.K && t < p.L0;
          const long o = okf[i] ? ((long)b * p.L1 + (t >> 1)) * SLU_NFILT + m : 0;
          rb[i] = p.route[o];
          gv[i] = __ldg(p.gy + o);
          par[i] = t & 1;
          if (++t == p.L0p) { t = 0; ++b; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool take = okf[i] && ((rb[i] & 1) == par[i]) && !(rb[i] & 4);
          va[u][i] = take ? ((rb[i] & 2) ? -gv[i] : gv[i]) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < (B_FR ? 0 : B_CH); ++u) {
      const int c = tid + u * THREADS;
      if (BMODE == 2) {
        const int kc = c & 3, r = c >> 2, n = n0 + r;
        ib_hi[u] = make_uint4(0, 0, 0, 0); ib_lo[u] = ib_hi[u];
        if (c < B_TOT && n < p.N) {
          const size_t e = ((size_t)tap * p.N + n) * p.Kp + k0 + kc * 8;
          ib_hi[u] = __ldg(reinterpret_cast<const uint4*>(p.Bimg + e));
          ib_lo[u] = __ldg(reinterpret_cast<const uint4*>(p.Bimg + (size_t)p.taps * p.N * p.Kp + e));
        }
      } else {
        int r, kc;
        if (BMODE == 0) { kc = c & 3; r = c >> 2; } else { r = c % BN; kc = c / BN; }
        const int n = n0 + r;
        const bool ok = (c < B_TOT) && (n < p.N);
        const float* base = p.B + (long)n * p.b_sn + (long)tap * p.b_stap;
        if (BMODE == 0) load8_kc(base, k0 + kc * 8, p.K, ok, vb[u]);
        else if (BMODE == 1) load8_strided(base, p.b_sk, k0 + kc * 8, p.K, p.T ? p.T : 1, p.b_kshift, ok, vb[u]);
        else {   // BMODE 3: B(n = tap sample 0..400, k = frame (b,t)) = x[b][80 t + n - 200]
          const int kf = k0 + kc * 8;
          int b = kf / p.L0p, t = kf - b * p.L0p;
#pragma unroll
          for (int i = 0; i < 8; ++i) {      // unconditional (clamped) loads, zeroed afterwards: keeps the 8 loads in flight together
            const int idx = SLU_STRIDE * t + n - SLU_PAD;
            const bool v_ok = ok && kf + i < p.K && t < p.L0 && idx >= 0 && idx < p.Ts;
            const float val = __ldg(p.B + (v_ok ? (long)b * p.Ts + idx : 0));
            vb[u][i] = v_ok ? val : 0.f;
            if (++t == p.L0p) { t = 0; ++b; }
          }
        }
      }
    }
  };
  prefetch(0);

  for (int i = 0; i < nkb; ++i) {
    const int s = i % STAGES;
    if (i >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)(((i / STAGES) - 1) & 1));
    const int k0 = ((kb_begin + i) % kb_per_tap) * BK;
    uint8_t* st = smem + s * S::STAGE;
    uint8_t* a_hi = st; uint8_t* a_lo = st + S::A_PART;
    uint8_t* b_hi = st + 2 * S::A_PART; uint8_t* b_lo = b_hi + S::B_PART;
    if (A_FR && fa_on) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        uint4 hi, lo; split8(fa[r], hi, lo);
        const uint32_t off = (uint32_t)fa_kc * S::LBO_A + (uint32_t)(fa_m4 * 4 + r) * 16;
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        *reinterpret_cast<uint4*>(a_lo + off) = lo;
      }
    }
    if (B_FR && fb_on) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        uint4 hi, lo; split8(fb[r], hi, lo);
        const uint32_t off = (uint32_t)fb_kc * S::LBO_B + (uint32_t)(fb_m4 * 4 + r) * 16;
        *reinterpret_cast<uint4*>(b_hi + off) = hi;
        *reinterpret_cast<uint4*>(b_lo + off) = lo;
      }
    }
#pragma unroll
    for (int u = 0; u < (A_FR ? 0 : A_CH); ++u) {
      uint4 hi, lo; split8(va[u], hi, lo);
      const uint32_t off = (uint32_t)a_kc[u] * S::LBO_A + (uint32_t)a_r[u] * 16;
      *reinterpret_cast<uint4*>(a_hi + off) = hi;
      *reinterpret_cast<uint4*>(a_lo + off) = lo;
    }
#pragma unroll
    for (int u = 0; u < (B_FR ? 0 : B_CH); ++u) {
      const int c = tid + u * THREADS;
      if (c < B_TOT) {
        int r, kc;
        if (BMODE == 1 || BMODE == 3) { r = c % BN; kc = c / BN; } else { kc = c & 3; r = c >> 2; }
        uint4 hi, lo;
        if (BMODE == 2) { hi = ib_hi[u]; lo = ib_lo[u]; } else split8(vb[u], hi, lo);
        const uint32_t off = (uint32_t)kc * S::LBO_B + (uint32_t)r * 16;
        *reinterpret_cast<uint4*>(b_hi + off) = hi;
        *reinterpret_cast<uint4*>(b_lo + off) = lo;
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
      const int n = n0 + c0 + lane;
      const bool n_ok = n < p.N;
      if (EPI == 1) {          // |.| + max over frame pairs (2j, 2j+1) + route bits; rows are frames m = b*L0p + t, t even first
        int m = m0 + q * 32;
        int b = m / p.L0p, t = m - b * p.L0p;
        for (int r = 0; r < 32; r += 2) {
          if (m + r < p.M && t < p.L0 && n_ok) {
            const float v0 = tr[r * 33 + lane], v1 = tr[(r + 1) * 33 + lane];
            const float a0 = fabsf(v0), a1 = (t + 1 < p.L0) ? fabsf(v1) : -1.f;
            const int sel = a1 > a0 ? 1 : 0;
            const float v = sel ? v1 : v0;
            const long o = ((long)b * p.L1 + (t >> 1)) * SLU_NFILT + n;
            p.C[o] = sel ? a1 : a0;
            if (p.route) p.route[o] = (uint8_t)(sel | ((v < 0.f) ? 2 : 0) | ((v == 0.f) ? 4 : 0));
          }
          t += 2;
          if (t >= p.L0p) { t -= p.L0p; ++b; }
        }
        __syncwarp();
        continue;
      }
      const int rows = min(32, p.M - (m0 + q * 32));
      if (p.split_k == 1 && rows == 32 && n0 + c0 + 32 <= p.N && (p.ldc & 3) == 0 && ((n0 + c0) & 3) == 0 &&
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
        const float bias = (p.bias && n_ok && blockIdx.z == 0) ? __ldg(p.bias + n) : 0.f;
        float* dst = p.C + (long)(m0 + q * 32) * p.ldc + n;
        if (n_ok) {
          if (rows == 32) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              float x = tr[r * 33 + lane] + bias;
              if (p.act == 1) x = x > 0.f ? x : x * p.slope;
              if (p.split_k > 1) atomicAdd(dst + (long)r * p.ldc, x); else dst[(long)r * p.ldc] = x;
            }
          } else {
            for (int r = 0; r < rows; ++r) {
              float x = tr[r * 33 + lane] + bias;
              if (p.act == 1) x = x > 0.f ? x : x * p.slope;
              if (p.split_k > 1) atomicAdd(dst + (long)r * p.ldc, x); else dst[(long)r * p.ldc] = x;
            }
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

template <int BN, int AMODE, int BMODE, int EPI>
int launch(const GemmParams& p, cudaStream_t stream) {
  const size_t smem = Smem<BN>::TOTAL;
  static int attr = slu_set_smem((const void*)gemm_tc_kernel<BN, AMODE, BMODE, EPI>, smem);
  if (attr) return attr;
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, p.split_k);
  gemm_tc_kernel<BN, AMODE, BMODE, EPI><<<grid, THREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <int AMODE, int BMODE>
int dispatch_bn(const GemmParams& p, cudaStream_t stream) {
  if (p.N <= 64) return launch<64, AMODE, BMODE, 0>(p, stream);
  if (p.N <= 128 || (AMODE == 1 && BMODE == 1)) return launch<128, AMODE, BMODE, 0>(p, stream);   // weight gradients: BN <= 128
  if constexpr (!(AMODE == 1 && BMODE == 1)) return launch<256, AMODE, BMODE, 0>(p, stream);
  return (int)cudaErrorInvalidValue;
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

// Pre-split a (strided) fp32 weight operand into the bf16 hi/lo image the GEMM's B side can copy verbatim.
// img must hold 2 * taps * N * Kp bf16 values, Kp = K rounded up to a multiple of 32.
int slu_presplit_rows(const float* W, long sn, long sk, long stap, int taps, int N, int K, int row_len, void* img, void* stream);
extern "C" int slu_presplit_bf16(const float* W, long sn, long sk, long stap, int taps, int N, int K, void* img, void* stream) {
  return slu_presplit_rows(W, sn, sk, stap, taps, N, K, 0, img, stream);
}
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

// Generic entry point (see include/slu_b200.h).  b_img != NULL selects the pre-split weight image for the B operand.
extern "C" int slu_gemm_tc(const float* A, long a_sm, long a_sk, const float* B, long b_sn, long b_sk, long b_stap,
                           const void* b_img, const float* bias, float* C, long ldc, int M, int N, int K, int taps,
                           int tap_pad, int T, int a_kshift, int b_kshift, int split_k, int act, float slope, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || taps <= 0 || split_k <= 0) return (int)cudaErrorInvalidValue;
  GemmParams p;
  p.A = A; p.a_sm = a_sm; p.a_sk = a_sk; p.B = B; p.b_sn = b_sn; p.b_sk = b_sk; p.b_stap = b_stap;
  p.Bimg = (const __nv_bfloat16*)b_img; p.Kp = (K + 31) / 32 * 32; p.bias = bias;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.taps = taps; p.tap_pad = tap_pad; p.T = T;
  p.a_kshift = a_kshift; p.b_kshift = b_kshift; p.split_k = split_k; p.act = act; p.slope = slope;
  p.Ts = p.L0 = p.L0p = p.L1 = 0; p.gy = nullptr; p.route = nullptr;
  const bool a_kc = (a_sk == 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (b_img) return a_kc ? dispatch_bn<0, 2>(p, st) : dispatch_bn<1, 2>(p, st);
  if (b_sk == 1) return a_kc ? dispatch_bn<0, 0>(p, st) : dispatch_bn<1, 0>(p, st);
  return a_kc ? dispatch_bn<0, 1>(p, st) : dispatch_bn<1, 1>(p, st);
}
