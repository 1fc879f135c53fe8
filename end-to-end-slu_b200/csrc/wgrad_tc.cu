// Weight-gradient ("reduction over frames") GEMM on tcgen05 with MN-major operands and TMEM-resident accumulators.
//
//   dW[tap][m][n] += sum over utterances b and frames t of  G[b][t][m] * X[b][t + shift0 + tap][n]      (0 outside [0, T))
//
// Both operands are stored frame-major in HBM (the reduction index is the slow one), which is exactly the MN-major
// operand form of tcgen05: a [frames][8-wide chunk] image with 16 bytes per frame row and chunk columns SBO apart.  A
// persistent CTA walks 64-frame tiles of its share of the (utterance, tile) list, stages the two images once per tile
// (coalesced 32-byte row segments -> bf16 hi/lo split -> one 16-byte shared-memory store each), issues
// hi*hi + hi*lo + lo*hi MMAs (M = 128 rows of G, N <= 128 columns of X, K = 16 frames) for every tap -- a tap is a
// frame shift, i.e. +16 bytes on the B descriptor start address, so the X image is staged once for all taps -- and
// keeps the NTAPS accumulators in tensor memory across ALL its tiles.  One flush with fp32 atomics at the end.
// Used for: dW_ih = dgx^T.x, dW_hh = [dr,dz|dhn]^T.h_{t-+1} (shift -1 / +1), Conv1d weight gradients (5 taps, shift -2).
// Replaces cuDNN's RNN / conv backward-weights (autograd of models.py:200, 232/262/686).
#include "common.cuh"
#include "tc05.cuh"

namespace {
using namespace tc05;

constexpr int THREADS = 256;
constexpr int TFW = 64;                        // frames per tile
constexpr int HALO = 8;                        // extra staged frame rows (taps reach at most +-4 frames)
constexpr int XR = TFW + HALO;                 // 72 staged X rows
constexpr uint32_t SBO_A = TFW * 16 + 16;      // 1040: chunk-column stride of the G image
constexpr uint32_t SBO_B = XR * 16 + 16;       // 1168
constexpr uint32_t A_PART = 16 * SBO_A;        // 16 chunks = 128 rows of G
__host__ __device__ constexpr uint32_t b_part(int nch) { return (uint32_t)nch * SBO_B; }

struct WgradParams {
  const float* G; long ldg;        // G[(b*T + t) * ldg + m], m in [0, m_valid)
  const float* X; long ldx;        // X[(b*T + t) * ldx + n], n in [0, n_valid)
  int m_valid, n_valid;
  int B, T, shift0, tiles_per_utt;
  float* out; long s_m, s_n, s_tap;     // out[m*s_m + n*s_n + tap*s_tap] += D_tap[m][n]
};

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) { return idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16); }

// 8 consecutive fp32 of one row (zero beyond `valid` elements), two 16-byte loads when aligned and complete
__device__ __forceinline__ void load_row8(const float* p, int valid, bool row_ok, float* v) {
  if (row_ok && valid >= 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (row_ok && i < valid) ? __ldg(p + i) : 0.f;     // predicated loads: never touches p when off
  }
}

template <int NTAPS, int NCH>          // NCH = N / 8 chunks of the X image (N = 8*NCH, a multiple of 16)
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  constexpr int N = 8 * NCH;
  static_assert(N % 16 == 0 && NTAPS * N <= 512, "accumulators must fit the 512 TMEM columns");
  constexpr uint32_t B_PART = b_part(NCH);
  constexpr uint32_t STAGE = 2 * A_PART + 2 * B_PART;
  constexpr uint32_t TCOLS = NTAPS * N <= 32 ? 32 : (NTAPS * N <= 64 ? 64 : (NTAPS * N <= 128 ? 128 : (NTAPS * N <= 256 ? 256 : 512)));
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t empty_bar[2], acc_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.z * N;
  const int mv = min(128, p.m_valid - m0), nv = min(N, p.n_valid - n0);       // valid rows / columns of this tile
  const int n_tiles = p.B * p.tiles_per_utt;

  if (tid == 0) { mbar_init(&empty_bar[0], 1); mbar_init(&empty_bar[1], 1); mbar_init(&acc_bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, TCOLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = idesc_mn(128, N);
  const int a_chunks = (mv + 7) / 8;        // chunks of G actually staged; the rest of the 16 are zeroed once below
  for (int s = 0; s < 2; ++s)
    for (int part = 0; part < 2; ++part) {
      uint8_t* g = smem + s * STAGE + part * A_PART + a_chunks * SBO_A;
      for (int i = tid * 16; i < (int)((16 - a_chunks) * SBO_A); i += THREADS * 16) *reinterpret_cast<uint4*>(g + i) = make_uint4(0, 0, 0, 0);
    }
  __syncthreads();

  // Register-prefetch pipeline over this CTA's tiles: the loads of tile i+1 are issued before tile i is converted, so
  // the memory latency hides behind the conversion, the barrier and the (asynchronous) MMAs.
  constexpr int XROWS = TFW + NTAPS - 1;
  constexpr int A_PER = 4, B_PER = (XROWS * NCH + THREADS - 1) / THREADS;
  static_assert(TFW * 16 <= A_PER * THREADS && B_PER <= 4, "staging task counts");
  const int a_tasks = TFW * a_chunks;
  int a_r[A_PER], a_cc[A_PER], b_r[B_PER], b_cc[B_PER];
#pragma unroll
  for (int u = 0; u < A_PER; ++u) { const int task = u * THREADS + tid; a_r[u] = task / a_chunks; a_cc[u] = task - a_r[u] * a_chunks; }
#pragma unroll
  for (int u = 0; u < B_PER; ++u) { const int task = u * THREADS + tid; b_r[u] = task / NCH; b_cc[u] = task - b_r[u] * NCH; }
  float va[A_PER][8], vb[B_PER][8];
  auto issue_loads = [&](int tile) {
    const int b = tile / p.tiles_per_utt, t0 = (tile - b * p.tiles_per_utt) * TFW;
#pragma unroll
    for (int u = 0; u < A_PER; ++u) {
      const int t = t0 + a_r[u];
      load_row8(p.G + ((long)b * p.T + min(t, p.T - 1)) * p.ldg + m0 + a_cc[u] * 8, mv - a_cc[u] * 8,
                u * THREADS + tid < a_tasks && t < p.T, va[u]);
    }
#pragma unroll
    for (int u = 0; u < B_PER; ++u) {
      const int t = t0 + p.shift0 + b_r[u];
      const bool ok = u * THREADS + tid < XROWS * NCH && t >= 0 && t < p.T;
      load_row8(p.X + ((long)b * p.T + (ok ? t : 0)) * p.ldx + n0 + b_cc[u] * 8, nv - b_cc[u] * 8, ok, vb[u]);
    }
  };
  if ((int)blockIdx.x < n_tiles) issue_loads(blockIdx.x);

  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    if (it >= 2) mbar_wait(&empty_bar[s], (uint32_t)(((it >> 1) - 1) & 1));
    uint8_t* st = smem + s * STAGE;
    uint8_t* a_hi = st; uint8_t* a_lo = st + A_PART; uint8_t* b_hi = st + 2 * A_PART; uint8_t* b_lo = b_hi + B_PART;
#pragma unroll
    for (int u = 0; u < A_PER; ++u) {
      if (u * THREADS + tid < a_tasks) {
        uint4 h, l; split8(va[u], h, l);
        const uint32_t off = (uint32_t)a_cc[u] * SBO_A + (uint32_t)a_r[u] * 16;
        *reinterpret_cast<uint4*>(a_hi + off) = h;
        *reinterpret_cast<uint4*>(a_lo + off) = l;
      }
    }
#pragma unroll
    for (int u = 0; u < B_PER; ++u) {
      if (u * THREADS + tid < XROWS * NCH) {
        uint4 h, l; split8(vb[u], h, l);
        const uint32_t off = (uint32_t)b_cc[u] * SBO_B + (uint32_t)b_r[u] * 16;
        *reinterpret_cast<uint4*>(b_hi + off) = h;
        *reinterpret_cast<uint4*>(b_lo + off) = l;
      }
    }
    fence_async_smem();                                   // before the prefetch: the proxy fence would wait for those loads
    if (tile + (int)gridDim.x < n_tiles) issue_loads(tile + gridDim.x);
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
        // MN-major, no swizzle: LBO = stride between 8-frame K groups (128 B), SBO = stride between 8-wide MN chunks
        const uint64_t ah0 = smem_desc(smem_u32(a_hi), 128, SBO_A), al0 = smem_desc(smem_u32(a_lo), 128, SBO_A);
        const uint64_t bh0 = smem_desc(smem_u32(b_hi), 128, SBO_B), bl0 = smem_desc(smem_u32(b_lo), 128, SBO_B);
#pragma unroll
        for (int kk = 0; kk < TFW / 16; ++kk) {
          const uint64_t ah = desc_advance(ah0, kk * 256), al = desc_advance(al0, kk * 256);
          const uint32_t acc = (it | kk) ? 1u : 0u;
#pragma unroll
          for (int tap = 0; tap < NTAPS; ++tap) {
            const uint64_t bh = desc_advance(bh0, kk * 256 + tap * 16), bl = desc_advance(bl0, kk * 256 + tap * 16);
            const uint32_t d = tmem + tap * N;
            mma_bf16(d, ah, bh, idesc, acc);
            mma_bf16(d, ah, bl, idesc, 1u);
            mma_bf16(d, al, bh, idesc, 1u);
          }
        }
        mma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
  }
  if (it > 0) {
    if (warp == 0 && elect_one()) mma_commit(&acc_bar);
    __syncwarp();
    mbar_wait(&acc_bar, 0);
    fence_after_sync();
    // flush: per warp 32 rows x 16 columns through a shared-memory transpose so that every red.add instruction covers
    // two rows x 16 consecutive columns (coalesced when s_n == 1)
    const int q = warp & 3, half = warp >> 2;
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 17);        // stage buffers are idle after acc_bar
    for (int c0 = half * 16; c0 < NTAPS * N; c0 += 32) {                  // two warps per lane quarter alternate 16-column groups
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) tr[lane * 17 + i] = v[i];
      __syncwarp();
      const int tap = c0 / N, nb = c0 - tap * N + (lane & 15);
      if (nb < nv) {
        float* dst = p.out + (long)(n0 + nb) * p.s_n + (long)tap * p.s_tap;
#pragma unroll 4
        for (int rr = 0; rr < 16; ++rr) {
          const int m = q * 32 + 2 * rr + (lane >> 4);
          if (m < mv) atomicAdd(dst + (long)(m0 + m) * p.s_m, tr[(2 * rr + (lane >> 4)) * 17 + (lane & 15)]);
        }
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int NTAPS, int NCH>
int launch(const WgradParams& p, int m_tiles, int n_tiles_n, cudaStream_t st) {
  constexpr uint32_t smem = 2 * (2 * A_PART + 2 * b_part(NCH));
  static int attr = slu_set_smem((const void*)wgrad_tc_kernel<NTAPS, NCH>, smem);
  if (attr) return attr;
  const long n_tiles = (long)p.B * p.tiles_per_utt;
  int gx = 148 / (m_tiles * n_tiles_n);                  // CTAs per output tile: fill the SMs ...
  if (gx > n_tiles / 6) gx = (int)(n_tiles / 6);        // ... but give every CTA >= 6 frame tiles per atomic flush
  if (gx < 1) gx = 1;
  wgrad_tc_kernel<NTAPS, NCH><<<dim3(gx, m_tiles, n_tiles_n), THREADS, smem, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

// out[m*s_m + n*s_n + tap*s_tap] += sum_{b,t} G[(b*T+t)*ldg + m] * X[(b*T + t + shift0 + tap)*ldx + n]   (frames outside [0,T) read 0)
// M, N arbitrary (tiled 128 x <=128); taps in {1, 5}.  `out` accumulates (zero it first for a plain gradient).
extern "C" int slu_wgrad_tc(const float* G, long ldg, int M, const float* X, long ldx, int N, int B, int T, int taps, int shift0,
                            float* out, long s_m, long s_n, long s_tap, void* stream) {
  if (M <= 0 || N <= 0 || B <= 0 || T <= 0 || (taps != 1 && taps != 5)) return (int)cudaErrorInvalidValue;
  if (taps > 1 && (shift0 < -4 || shift0 + taps - 1 > 4)) return (int)cudaErrorInvalidValue;
  WgradParams p;
  p.G = G; p.ldg = ldg; p.X = X; p.ldx = ldx; p.m_valid = M; p.n_valid = N; p.B = B; p.T = T; p.shift0 = shift0;
  p.tiles_per_utt = (T + TFW - 1) / TFW; p.out = out; p.s_m = s_m; p.s_n = s_n; p.s_tap = s_tap;
  const int m_tiles = (M + 127) / 128;
  cudaStream_t st = (cudaStream_t)stream;
  if (taps == 5) {
    if (N <= 64) return launch<5, 8>(p, m_tiles, 1, st);
    if (N <= 80) return launch<5, 10>(p, m_tiles, 1, st);
    return (int)cudaErrorInvalidValue;
  }
  if (N <= 64) return launch<1, 8>(p, m_tiles, 1, st);
  return launch<1, 16>(p, m_tiles, (N + 127) / 128, st);
}
