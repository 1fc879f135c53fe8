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
//
// Persistent, warp-specialised kernel: one CTA per SM walks [128 x BN] output tiles (n fastest, so the activation rows of
// a tile row are re-read from L2).  Roles (18 warps):
//   warps 13-16 activation loaders: every k-block's [128 rows x 128 B] fp32 tile is fetched with 16-byte cp.async copies
//              (coalesced full lines, zero-fill for rows shifted out of their utterance / the K tail) into a ring of
//              staging slots -- several k-blocks in flight, no registers held; completion arrives on an mbarrier;
//   warps 4-11 converters (2 groups of 4 taking alternate k-blocks): staging slot -> bf16 hi + lo, stored K-major (no swizzle,
//              padded leading-byte-offset) into the operand ring;
//   warp 12    weight loader: the weights are PRE-SPLIT once per call (slu_presplit_bf16) into a k-chunk-major image, so
//              a stage's weight tile is 8 TMA bulk copies, no thread work;
//   warp 17    MMA issuer: hi*hi + hi*lo + lo*hi tcgen05.mma per K=16 step into one of TWO [128 x BN] fp32 accumulators
//              in TMEM; tcgen05.commit frees the operand stage / publishes the accumulator;
//   warps 0-3  epilogue: tcgen05.ld, bias / LeakyReLU, transpose through shared memory, coalesced 16-byte row stores;
//              the accumulator is released right after its last tcgen05.ld, so tile i+1's MMAs overlap tile i's stores.
// Contract of the 16-byte copies: A 16-byte aligned, lda and K multiples of 4 floats (every call site of the model).
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
  int dbg;                           // developer switch (slu_debug_gemm_mode): 1 = no MMAs, 2 = no conversion, 4 = no global stores, 8 = no epilogue
};
int g_gemm_dbg = 0;
// The ablation switches cost ~7 % of this kernel's time in the train step even when off (measured): they exist only in builds with
// -DSLU_KERNEL_DEBUG (SLU_KERNEL_DEBUG=1 python __graft_entry__.py), which is what tools/gemm_rate.py needs.
#ifdef SLU_KERNEL_DEBUG
#define SLU_DBG(p) ((p).dbg)
#else
#define SLU_DBG(p) 0
#endif

constexpr int BM = 128, BK = 32;

__host__ __device__ constexpr uint32_t tmem_cols(int bn) { return bn <= 32 ? 32 : (bn <= 64 ? 64 : (bn <= 128 ? 128 : 256)); }

// warps 0-3 epilogue (TMEM lane quarters); then CONV_GROUPS groups of 4 converter warps (group g takes k-blocks g, g + groups, ..:
// converting one k-block is a chain of latencies -- barrier wait, LDS, split, STS, proxy fence, arrive -- so two groups in flight
// double the rate of the activation path); weight loader; 4 activation loader warps; MMA issuer
constexpr int EPI_WARPS = 4, CONV_GROUPS = 2, CONV_WARPS = 4 * CONV_GROUPS;
constexpr int ALOAD_WARPS = 4;
constexpr int WLOAD_WARP = EPI_WARPS + CONV_WARPS, ALOAD_WARP = WLOAD_WARP + 1, MMA_WARP = ALOAD_WARP + ALOAD_WARPS;
constexpr int THREADS = (MMA_WARP + 1) * 32;                          // 576
constexpr int CONV_THREADS = 128, EPI_THREADS = EPI_WARPS * 32, ALOAD_THREADS = ALOAD_WARPS * 32;   // CONV_THREADS: per group
constexpr uint32_t STG_ROW = BK * 4 + 16;                             // staged fp32 row segment + pad: conflict-free 16-byte reads
constexpr uint32_t STG_SLOT = BM * STG_ROW;

template <int BN>
struct Smem {
  static constexpr int STAGES = BN > 128 ? 3 : 4;                     // operand ring (bf16 hi/lo A + B tiles)
  static constexpr int NSTG = BN > 128 ? 3 : 4;                       // fp32 staging ring of the activation loader
                                                                      // (the activation path is a latency ring: k-blocks per round trip = slots)
  static constexpr uint32_t LBO_A = BM * 16 + 16, LBO_B = BN * 16 + 16;
  static constexpr uint32_t A_PART = (BK / 8) * LBO_A, B_PART = (BK / 8) * LBO_B;      // one of hi / lo
  static constexpr uint32_t STAGE = 2 * A_PART + 2 * B_PART;
  static constexpr uint32_t PIPE = STAGES * STAGE;
  static constexpr uint32_t STG = NSTG * STG_SLOT;
  static constexpr uint32_t TRANS = EPI_WARPS * 32 * 33 * 4;                            // epilogue transpose buffers
  static constexpr uint32_t TOTAL = PIPE + STG + TRANS;
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using S = Smem<BN>;
  constexpr int STAGES = S::STAGES, NSTG = S::NSTG;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_a[STAGES], full_b[STAGES], empty_bar[STAGES], stg_full[NSTG], stg_empty[NSTG], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int kb_per_tap = (p.K + BK - 1) / BK;
  const int nkb = p.taps * kb_per_tap;
  const int ntiles_n = (p.N + BN - 1) / BN;
  const int ntiles = ((p.M + BM - 1) / BM) * ntiles_n;
  uint8_t* const stg_base = smem + S::PIPE;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_a[s], CONV_THREADS);
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&stg_full[s], ALOAD_THREADS);       // one (asynchronous) arrival per loader thread
      mbar_init(&stg_empty[s], CONV_THREADS);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], EPI_THREADS);
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 2 * tmem_cols(BN));
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;

  if (warp >= ALOAD_WARP && warp < MMA_WARP) {
    // ================= activation loaders: [128 rows x 128 B] per k-block as 16-byte cp.async copies =================
    const int lt = tid - ALOAD_WARP * 32;
    constexpr int PER = BM * (BK / 4) / ALOAD_THREADS;               // 8 pieces of 16 B per thread and k-block
    const int piece = lt & 7, r0 = lt >> 3;                          // rows r0 + 16 u: a warp covers 4 full 128-byte lines
    const bool shifted = p.taps > 1 || p.tap_pad != 0;
    const uint32_t dst0 = smem_u32(stg_base) + (uint32_t)r0 * STG_ROW + (uint32_t)piece * 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m0 = (tile / ntiles_n) * BM;
      // per-tile row state: everything that does not depend on the k-block stays out of the inner loop
      const float* rowp[PER];
      int a_t[PER];
      bool okm[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int m = m0 + r0 + 16 * u;
        okm[u] = m < p.M;
        a_t[u] = (shifted && p.T) ? m % p.T : 0;                      // frame of the row inside its utterance
        rowp[u] = p.A + (long)m * p.lda + piece * 4;
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int slot = it % NSTG;
        if (it >= NSTG) mbar_wait(&stg_empty[slot], (uint32_t)(((it / NSTG) - 1) & 1));
        const int tap = kb / kb_per_tap, k0 = (kb - tap * kb_per_tap) * BK;
        const int sh = tap - p.tap_pad;
        const long off = (long)sh * p.lda + k0;                      // same for every row of the k-block
        const uint32_t nb = (p.K - k0 - piece * 4) > 0 ? 16u : 0u;    // K % 4 == 0: a piece is inside the row or beyond it
        const uint32_t dst = dst0 + (uint32_t)slot * STG_SLOT;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          bool ok = okm[u];
          if (shifted && p.T) { const int t = a_t[u] + sh; ok = ok && t >= 0 && t < p.T; }
          const uint32_t nbytes = ok ? nb : 0u;
          cp_async16_s(dst + (uint32_t)(16 * u) * STG_ROW, nbytes ? rowp[u] + off : p.A, nbytes);
        }
        cp_async_mbar_arrive_noinc(&stg_full[slot]);
      }
    }
  } else if (warp >= EPI_WARPS && warp < WLOAD_WARP) {
    // ================= converters: fp32 staging slot -> bf16 hi/lo K-major operand stage =================
    const int ptid = (tid - EPI_THREADS) & (CONV_THREADS - 1), grp = (tid - EPI_THREADS) / CONV_THREADS;
    constexpr int A_CH = BM * (BK / 8) / CONV_THREADS;              // 4 chunks of 8 consecutive k per thread and k-block
    int a_r[A_CH], a_kc[A_CH];
#pragma unroll
    for (int u = 0; u < A_CH; ++u) {
      const int c = ptid + u * CONV_THREADS;
      a_kc[u] = c & 3; a_r[u] = c >> 2;
    }
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;
    const int total = my_tiles * nkb;                                 // k-blocks of this CTA, in pipeline order
    {
      for (int it = grp; it < total; it += CONV_GROUPS) {
        const int s = it % STAGES, slot = it % NSTG;
        mbar_wait(&stg_full[slot], (uint32_t)((it / NSTG) & 1));                       // this k-block's rows have landed
        const uint8_t* src = stg_base + slot * STG_SLOT;
        float va[A_CH][8];
#pragma unroll
        for (int u = 0; u < A_CH; ++u) {
          const float4 x0 = *reinterpret_cast<const float4*>(src + a_r[u] * STG_ROW + a_kc[u] * 32);
          const float4 x1 = *reinterpret_cast<const float4*>(src + a_r[u] * STG_ROW + a_kc[u] * 32 + 16);
          va[u][0] = x0.x; va[u][1] = x0.y; va[u][2] = x0.z; va[u][3] = x0.w;      // zero-filled by the loaders where invalid
          va[u][4] = x1.x; va[u][5] = x1.y; va[u][6] = x1.z; va[u][7] = x1.w;
        }
        // Values are on their way to registers: the loader may refill the slot.  An mbarrier arrive does not wait for the thread's
        // outstanding shared-memory loads (measured in wgrad_tc.cu, whose TMA refill overtook them), but here the refill is made of
        // cp.async instructions that enter the SM's load/store queue behind these loads and write shared memory a global-memory
        // latency later; two rounds of parity tests and sanitizer runs never saw a stale row here.
        mbar_arrive(&stg_empty[slot]);
        if (it >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)(((it / STAGES) - 1) & 1));
        if (SLU_DBG(p) & 2) { mbar_arrive(&full_a[s]); continue; }
        uint8_t* a_hi = smem + s * S::STAGE;
        uint8_t* a_lo = a_hi + S::A_PART;
#pragma unroll
        for (int u = 0; u < A_CH; ++u) {
          uint4 hi, lo; split8(va[u], hi, lo);
          const uint32_t off = (uint32_t)a_kc[u] * S::LBO_A + (uint32_t)a_r[u] * 16;
          *reinterpret_cast<uint4*>(a_hi + off) = hi;
          *reinterpret_cast<uint4*>(a_lo + off) = lo;
        }
        fence_async_smem();               // generic-proxy stores -> visible to the tensor core
        mbar_arrive(&full_a[s]);
      }
    }
  } else if (warp == WLOAD_WARP) {
    // ================= weight loader: 8 TMA bulk copies per stage from the k-chunk-major image =================
    const int KC = p.Kp / 8;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int n0 = (tile % ntiles_n) * BN;
      const int rows = min(BN, p.N - n0);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)(((it / STAGES) - 1) & 1));
        if (elect_one()) {
          const int tap = kb / kb_per_tap, kc0 = (kb % kb_per_tap) * (BK / 8);
          uint8_t* b_hi = smem + s * S::STAGE + 2 * S::A_PART;
          mbar_arrive_expect_tx(&full_b[s], (uint32_t)(2 * (BK / 8) * rows * 16));
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int kc = 0; kc < BK / 8; ++kc) {
              const size_t e = ((((size_t)part * p.taps + tap) * KC + kc0 + kc) * p.N + n0) * 8;
              tma_load_1d(b_hi + part * S::B_PART + kc * S::LBO_B, p.Wimg + e, (uint32_t)rows * 16u, &full_b[s]);
            }
        }
        __syncwarp();
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_bf16_f32(BM, BN);
    int it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      if (ti >= 2) mbar_wait(&acc_empty[buf], (uint32_t)(((ti >> 1) - 1) & 1));   // epilogue has drained this accumulator
      fence_after_sync();
      const uint32_t d_tmem = tmem + (uint32_t)buf * tmem_cols(BN);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)((it / STAGES) & 1);
        mbar_wait(&full_a[s], ph);
        mbar_wait(&full_b[s], ph);
        fence_after_sync();
        if (elect_one()) {
          const int k0 = (kb % kb_per_tap) * BK;
          const int nk16 = min(BK / 16, (p.K - k0 + 15) / 16);
          const uint32_t sa = smem_u32(smem + s * S::STAGE);
          const uint64_t ah0 = smem_desc(sa, S::LBO_A, 128), al0 = smem_desc(sa + S::A_PART, S::LBO_A, 128);
          const uint64_t bh0 = smem_desc(sa + 2 * S::A_PART, S::LBO_B, 128), bl0 = smem_desc(sa + 2 * S::A_PART + S::B_PART, S::LBO_B, 128);
          uint32_t acc = kb > 0 ? 1u : 0u;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            if (kk < nk16 && !(SLU_DBG(p) & 1)) {
              const uint64_t ah = desc_advance(ah0, kk * 2 * S::LBO_A), al = desc_advance(al0, kk * 2 * S::LBO_A);
              const uint64_t bh = desc_advance(bh0, kk * 2 * S::LBO_B), bl = desc_advance(bl0, kk * 2 * S::LBO_B);
              mma_bf16(d_tmem, ah, bh, idesc, acc); acc = 1u;
              mma_bf16(d_tmem, ah, bl, idesc, 1u);
              mma_bf16(d_tmem, al, bh, idesc, 1u);
            }
          }
          mma_commit(&empty_bar[s]);                        // stage s may be refilled once these MMAs have read it
          if (kb == nkb - 1) mma_commit(&acc_full[buf]);    // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> smem transpose -> coalesced global rows =================
    const int q = warp;                                            // TMEM lane quarter == rows 32q .. 32q+31 of the tile
    float* tr = reinterpret_cast<float*>(smem + S::PIPE + S::STG) + warp * (32 * 33);
    const bool vec_ok = (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      const int m0 = (tile / ntiles_n) * BM, n0 = (tile % ntiles_n) * BN;
      const int ncols = min(BN, p.N - n0);
      const int nchunks = (ncols + 31) / 32;
      mbar_wait(&acc_full[buf], (uint32_t)((ti >> 1) & 1));
      fence_after_sync();
      const uint32_t t_acc = tmem + (uint32_t)buf * tmem_cols(BN) + ((uint32_t)(q * 32) << 16);
      const int rows = min(32, p.M - (m0 + q * 32));               // may be <= 0 for the tail tile
#pragma unroll 1
      for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * 32;
        // fast path: 16-byte stores, a warp instruction writes 4 rows x 128 B.  Its bias vector is requested BEFORE the accumulator
        // read: fetched where it is used it exposes an L2 round trip per chunk (tools/gemm_rate.py: -50 % on the x-projection's epilogue)
        const bool fast = rows == 32 && c0 + 32 <= ncols && vec_ok && ((n0 + c0) & 3) == 0;
        const int cq = (lane & 7) * 4, r0 = lane >> 3;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fast && p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + cq));
        float v[32];
        tmem_ld16(t_acc + c0, v);
        tmem_ld16(t_acc + c0 + 16, v + 16);
        tmem_ld_wait();
        if (ch == nchunks - 1) {          // last read of this accumulator: hand it back to the MMA warp
          fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
        }
        if (SLU_DBG(p) & 8) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) tr[lane * 33 + i] = v[i];
        __syncwarp();
        if (fast) {
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
            if (!(SLU_DBG(p) & 4)) *reinterpret_cast<float4*>(dst4 + (long)r * p.ldc) = x;
          }
        } else {
          const int n = n0 + c0 + lane;
          if (c0 + lane < ncols) {
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
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 2 * tmem_cols(BN));
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN>
int launch(const GemmParams& p, cudaStream_t stream) {
  const size_t smem = Smem<BN>::TOTAL;
  SLU_SMEM_ONCE(gemm_tc_kernel<BN>, smem);
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tc_kernel<BN><<<grid, THREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

// fp32 strided weights -> bf16 hi / lo images [2][taps][N][Kp] (zero padded to Kp, a multiple of 32)
// CHUNK_MAJOR: image [2][taps][Kp/8][N][8] (every 8-element k-chunk of all N rows contiguous = what one TMA bulk copy of the
// GEMM's weight loader moves); otherwise row-major [2][taps][N][Kp].
template <bool CHUNK_MAJOR>
__global__ void presplit_kernel(const float* __restrict__ W, long sn, long sk, long stap, int taps, int N, int K, int Kp,
                                int row_len, __nv_bfloat16* __restrict__ img) {
  const size_t total = (size_t)taps * N * Kp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int k, n, tap;
    if (CHUNK_MAJOR) {
      const int e = (int)(i & 7);
      const size_t r = i >> 3;
      n = (int)(r % N);
      const size_t c = r / N;
      k = (int)(c % (Kp / 8)) * 8 + e;
      tap = (int)(c / (Kp / 8));
    } else {
      k = (int)(i % Kp);
      const size_t tn = i / Kp;
      n = (int)(tn % N); tap = (int)(tn / N);
    }
    float v = 0.f;
    if (k < K && (row_len == 0 || (long)k * sk + (long)tap * stap < row_len)) v = W[(long)n * sn + (long)k * sk + (long)tap * stap];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    img[i] = h;
    img[total + i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// Up to 16 pre-split jobs in ONE launch (blockIdx.y = job): every weight image a forward/backward pass needs, made at once.
struct PresplitJob { const float* W; long sn, sk, stap; int taps, N, K, pad_; __nv_bfloat16* img; };
struct PresplitTable { PresplitJob job[16]; };

__global__ void presplit_multi_kernel(const __grid_constant__ PresplitTable tab) {
  const PresplitJob& j = tab.job[blockIdx.y];
  const int Kp = (j.K + 31) / 32 * 32, KC = Kp / 8;
  const size_t total = (size_t)j.taps * j.N * Kp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    const size_t r = i >> 3;
    const int n = (int)(r % j.N);
    const size_t c = r / j.N;
    const int k = (int)(c % KC) * 8 + e, tap = (int)(c / KC);
    const float v = k < j.K ? j.W[(long)n * j.sn + (long)k * j.sk + (long)tap * j.stap] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    j.img[i] = h;
    j.img[total + i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace

// like slu_presplit_bf16, with elements whose in-row offset k*sk + tap*stap reaches row_len read as 0 (sinc_tc.cu's TMA-fed filter bank)
int slu_presplit_rows_cm(const float* W, long sn, long sk, long stap, int taps, int N, int K, int row_len, void* img, void* stream) {
  if (taps <= 0 || N <= 0 || K <= 0) return (int)cudaErrorInvalidValue;
  const int Kp = (K + 31) / 32 * 32;
  const size_t total = (size_t)taps * N * Kp;
  int grid = (int)((total + 255) / 256);
  if (grid > 1184) grid = 1184;
  presplit_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(W, sn, sk, stap, taps, N, K, Kp, row_len, (__nv_bfloat16*)img);
  return (int)cudaGetLastError();
}

// Pre-split a (strided) fp32 weight operand W(n, tap, k) = W[n*sn + k*sk + tap*stap] into the k-chunk-major bf16 hi/lo image the
// GEMM's weight loader copies verbatim.  img must hold 2 * taps * N * Kp bf16 values, Kp = K rounded up to a multiple of 32.
extern "C" int slu_presplit_bf16(const float* W, long sn, long sk, long stap, int taps, int N, int K, void* img, void* stream) {
  if (taps <= 0 || N <= 0 || K <= 0) return (int)cudaErrorInvalidValue;
  const int Kp = (K + 31) / 32 * 32;
  const size_t total = (size_t)taps * N * Kp;
  int grid = (int)((total + 255) / 256);
  if (grid > 1184) grid = 1184;
  presplit_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(W, sn, sk, stap, taps, N, K, Kp, 0, (__nv_bfloat16*)img);
  return (int)cudaGetLastError();
}

// n <= 16 jobs of slu_presplit_bf16 in one launch.  `jobs` is a HOST array of {W, sn, sk, stap, taps, N, K, (pad), img}.
extern "C" int slu_presplit_multi(const void* jobs, int n, void* stream) {
  if (n <= 0) return 0;
  if (n > 16 || !jobs) return (int)cudaErrorInvalidValue;
  PresplitTable tab;
  const PresplitJob* src = (const PresplitJob*)jobs;
  size_t most = 0;
  for (int i = 0; i < n; ++i) {
    tab.job[i] = src[i];
    if (src[i].taps <= 0 || src[i].N <= 0 || src[i].K <= 0 || !src[i].W || !src[i].img) return (int)cudaErrorInvalidValue;
    const size_t total = (size_t)src[i].taps * src[i].N * ((src[i].K + 31) / 32 * 32);
    most = total > most ? total : most;
  }
  for (int i = n; i < 16; ++i) tab.job[i] = tab.job[0];
  int gx = (int)((most + 255) / 256);
  if (gx > 296) gx = 296;
  presplit_multi_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(tab);
  return (int)cudaGetLastError();
}

// C[m][n] = sum_tap sum_k A[(m + tap - tap_pad)*lda + k] * W(n, tap, k) (+ bias[n]) (LeakyReLU if act == 1); see include/slu_b200.h
extern "C" int slu_gemm_tc(const float* A, long lda, const void* w_img, const float* bias, float* C, long ldc, int M, int N, int K,
                           int taps, int tap_pad, int T, int act, float slope, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || taps <= 0 || !w_img) return (int)cudaErrorInvalidValue;
  GemmParams p;
  p.A = A; p.lda = lda; p.Wimg = (const __nv_bfloat16*)w_img; p.Kp = (K + 31) / 32 * 32; p.bias = bias;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.taps = taps; p.tap_pad = tap_pad; p.T = T; p.act = act; p.slope = slope; p.dbg = g_gemm_dbg;
  // TMA source alignment: 16-byte aligned operands, row pitch and K in whole 16-byte units
  if ((reinterpret_cast<uintptr_t>(w_img) & 15) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0 || (lda & 3) != 0 || (K & 3) != 0)
    return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  // Column-tile width from a measured cost model: a k-block costs ~1100 cycles on the activation path (independent of BN) plus
  // ~1.7 cycles per output column; pick the BN with the fewest (waves of the persistent grid) x (per-k-block cost).
  const int mt = (M + BM - 1) / BM, sms = sm_count();
  int best_bn = 0;
  long best_cost = 0;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (bn > 64 && N <= bn / 2) continue;                     // tile would be mostly padding
    const long tiles = (long)mt * ((N + bn - 1) / bn);
    const long cost = ((tiles + sms - 1) / sms) * (1100 + (17 * bn) / 10);
    if (best_bn == 0 || cost < best_cost) { best_bn = bn; best_cost = cost; }
  }
  if (best_bn == 64) return launch<64>(p, st);
  if (best_bn == 128) return launch<128>(p, st);
  return launch<256>(p, st);
}

// Developer switch for tools/gemm_rate.py: bit 0 skips the MMAs, bit 1 the operand conversion, bit 2 the global stores of the fast
// epilogue path, bit 3 the whole epilogue after the accumulator read.  Results are meaningless with any bit set.
extern "C" int slu_debug_gemm_mode(int mode) {
  g_gemm_dbg = mode;
#ifdef SLU_KERNEL_DEBUG
  return 0;
#else
  return mode ? (int)cudaErrorNotSupported : 0;      // built without the switches
#endif
}
