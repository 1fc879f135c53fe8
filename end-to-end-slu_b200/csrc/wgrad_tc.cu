// Weight-gradient ("reduction over frames") GEMM on tcgen05 with MN-major operands and TMEM-resident accumulators.
//
//   dW[tap][m][n] += sum over utterances b and frames t of  G[b][t][m] * X[b][t + shift0 + tap][n]      (0 outside [0, T))
//
// Both operands are stored frame-major in HBM (the reduction index is the slow one), which is exactly the MN-major
// operand form of tcgen05: a [8-wide column chunk][frame] image with 16 bytes per frame and chunk columns SBO apart.
// A persistent, warp-specialised CTA owns one [MT*128 x N] block of dW and walks 16-frame tiles of its share of the
// (utterance, tile) list:
//   warps 8-11 loaders: the tile's G rows and X rows arrive as 16-byte cp.async copies (coalesced full lines, zero-fill
//              for frames outside the utterance and for the column tails) in a 3-slot fp32 staging ring; completion is
//              an asynchronous mbarrier arrival, no registers are held across the memory latency;
//   warps 0-7  converters (two groups of 4 warps taking alternate tiles): staging slot -> bf16 hi + lo MN-major images (lanes walk frames: conflict-free 16-byte
//              shared-memory reads and stores), 2-stage operand ring;
//   warp 12    MMA issuer: per tile hi*hi + hi*lo + lo*hi for every 128-row block of G and every tap (a tap is a frame
//              shift = +16 bytes on the B descriptor, so X is staged once for all taps); the MT * NTAPS accumulators
//              [128 x N] stay in tensor memory across ALL tiles of the CTA.
// One flush with fp32 atomics at the end.  G may come from two tensors (rows [0, m_split) from G0, the rest from G1):
// dW_hh = [dr, dz | dhn]^T . h_{t-+1} is ONE launch per direction.
// Used for: dW_ih = dgx^T.x, dW_hh (shift -1 / +1), Conv1d weight gradients (5 taps, shift -2).
// Replaces cuDNN's RNN / conv backward-weights (autograd of models.py:200, 232/262/686).
// Contract of the 16-byte copies: operands 16-byte aligned; ldg, ldx, M, m_split, N multiples of 4 floats.
#include "common.cuh"
#include "tc05.cuh"

namespace {
using namespace tc05;

constexpr int TF = 16;                         // frames per tile (= one K step of the MMA)
constexpr int NSTG = 3, NOPS = 2;              // fp32 staging slots, bf16 operand stages
constexpr int CONV_GROUPS = 2;                 // converter groups of 4 warps taking alternate tiles (a tile's conversion is a latency chain)
constexpr int CONV_WARPS = 4 * CONV_GROUPS, LOAD_WARPS = 4, MMA_WARP = CONV_WARPS + LOAD_WARPS;
constexpr int THREADS = (MMA_WARP + 1) * 32;   // 416
constexpr int CONV_THREADS = 128, LOAD_THREADS = LOAD_WARPS * 32;      // CONV_THREADS: per group

struct WgradParams {
  const float* G0; long ldg0;      // G rows [0, m_split):   G0[(b*T + t) * ldg0 + m]
  const float* G1; long ldg1;      // G rows [m_split, M):   G1[(b*T + t) * ldg1 + (m - m_split)]
  int m_split;
  const float* X; long ldx;        // X[(b*T + t) * ldx + n], n in [0, n_valid)
  int m_valid, n_valid;
  int B, T, shift0, tiles_per_utt;
  float* out; long s_m, s_n, s_tap;     // out[m*s_m + n*s_n + tap*s_tap] += D_tap[m][n]
};

// fp32 x4 vector reduction to global memory (sm_90+): one L2 operation instead of four
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) { return idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16); }

template <int MT, int NCH, int NTAPS>
struct Cfg {
  static constexpr int MC = MT * 128, N = 8 * NCH;                   // staged G columns, X columns
  static constexpr int XR = TF + NTAPS - 1;                          // staged X rows
  static constexpr uint32_t PITCH_G = MC * 4 + 16, PITCH_X = N * 4 + 16;          // staging row pitch = 16 mod 128: lanes walking
  static constexpr uint32_t STG_G = TF * PITCH_G, STG_X = XR * PITCH_X;           // frames read 16 bytes conflict-free
  static constexpr uint32_t STG_SLOT = STG_G + STG_X;
  static constexpr uint32_t SBO_A = TF * 16 + 16, SBO_B = XR * 16 + 16;           // chunk-column strides of the bf16 images
  static constexpr uint32_t A_PART = (MC / 8) * SBO_A, B_PART = NCH * SBO_B;     // one of hi / lo
  static constexpr uint32_t OP_STAGE = 2 * A_PART + 2 * B_PART;
  static constexpr uint32_t PIPE = NOPS * OP_STAGE;
  static constexpr uint32_t FLUSH = 8 * 32 * 17 * 4;                              // transposers of the final flush (alias PIPE)
  static constexpr uint32_t TOTAL = (PIPE > FLUSH ? PIPE : FLUSH) + NSTG * STG_SLOT;
  static constexpr int ACC = MT * NTAPS * N;
  static constexpr uint32_t TCOLS = ACC <= 32 ? 32 : (ACC <= 64 ? 64 : (ACC <= 128 ? 128 : (ACC <= 256 ? 256 : 512)));
  static_assert(N % 16 == 0 && N <= 256 && ACC <= 512, "accumulators must fit the 512 TMEM columns");
};

template <int MT, int NCH, int NTAPS>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  using C = Cfg<MT, NCH, NTAPS>;
  constexpr int N = C::N, MC = C::MC, XR = C::XR;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t stg_full[NSTG], stg_empty[NSTG], op_full[NOPS], op_empty[NOPS], acc_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.y * MC, n0 = blockIdx.z * N;
  const int mv = min(MC, p.m_valid - m0), nv = min(N, p.n_valid - n0);          // valid G / X columns of this block
  const int n_tiles = p.B * p.tiles_per_utt;
  uint8_t* const stg_base = smem + (C::PIPE > C::FLUSH ? C::PIPE : C::FLUSH);

  if (tid == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(&stg_full[s], LOAD_THREADS); mbar_init(&stg_empty[s], CONV_THREADS); }
    for (int s = 0; s < NOPS; ++s) { mbar_init(&op_full[s], CONV_THREADS); mbar_init(&op_empty[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, C::TCOLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++my_tiles;

  if (warp >= CONV_WARPS && warp < MMA_WARP) {
    // ================= loaders: one tile = TF rows of G (MC columns) + XR rows of X (N columns), 16 bytes per copy =================
    // A warp walks whole frame rows (frames lw, lw+4, ..), lanes walk 16-byte column pieces: one cp.async instruction
    // moves 512 contiguous bytes.  Column state (source tensor, validity) is fixed per thread and hoisted out of the loops.
    const int lw = warp - CONV_WARPS;                         // loader warp 0..3
    constexpr int GJ = MC / 128, XJ = (N / 4 + 31) / 32;     // pieces per lane and row: G (= MT), X
    constexpr int GF = TF / LOAD_WARPS, XF = (XR + LOAD_WARPS - 1) / LOAD_WARPS;
    const float* gcol[GJ]; long gld[GJ]; bool gok[GJ];
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      const int c = (lane + 32 * j) * 4, m = m0 + c;           // block-relative / global G column of this piece
      gok[j] = c < mv;
      const bool first = m < p.m_split;
      gcol[j] = first ? p.G0 + m : p.G1 + (m - p.m_split);
      gld[j] = first ? p.ldg0 : p.ldg1;
    }
    const float* xcol[XJ]; bool xok[XJ];
#pragma unroll
    for (int j = 0; j < XJ; ++j) {
      const int c = (lane + 32 * j) * 4;
      xok[j] = c < nv && c < N;
      xcol[j] = p.X + n0 + c;
    }
    const uint32_t stg0 = smem_u32(stg_base);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int slot = it % NSTG;
      if (it >= NSTG) mbar_wait(&stg_empty[slot], (uint32_t)(((it / NSTG) - 1) & 1));
      const int b = tile / p.tiles_per_utt, t0 = (tile - b * p.tiles_per_utt) * TF;
      const long fr0 = (long)b * p.T;
      const uint32_t dst = stg0 + (uint32_t)slot * C::STG_SLOT;
#pragma unroll
      for (int fi = 0; fi < GF; ++fi) {
        const int f = lw + LOAD_WARPS * fi, t = t0 + f;
        const bool okf = t < p.T;
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
          const bool ok = okf && gok[j];
          cp_async16_s(dst + (uint32_t)f * C::PITCH_G + (uint32_t)(lane + 32 * j) * 16, ok ? gcol[j] + (fr0 + t) * gld[j] : p.G0, ok ? 16u : 0u);
        }
      }
      const uint32_t dstx = dst + C::STG_G;
#pragma unroll
      for (int fi = 0; fi < XF; ++fi) {
        const int f = lw + LOAD_WARPS * fi, t = t0 + p.shift0 + f;
        const bool okf = f < XR && t >= 0 && t < p.T;
#pragma unroll
        for (int j = 0; j < XJ; ++j) {
          if ((lane + 32 * j) * 4 < N && f < XR) {
            const bool ok = okf && xok[j];
            cp_async16_s(dstx + (uint32_t)f * C::PITCH_X + (uint32_t)(lane + 32 * j) * 16, ok ? xcol[j] + (fr0 + t) * p.ldx : p.X, ok ? 16u : 0u);
          }
        }
      }
      cp_async_mbar_arrive_noinc(&stg_full[slot]);
    }
  } else if (warp < CONV_WARPS) {
    // ================= converters: fp32 staging -> bf16 hi/lo MN-major images (lanes walk frames) =================
    constexpr int GT = (MC / 8) * TF, XT = NCH * XR;        // (chunk, frame) tasks
    const int grp = warp >> 2, ctid = tid & (CONV_THREADS - 1);
    for (int it = grp; it < my_tiles; it += CONV_GROUPS) {
      const int slot = it % NSTG, s = it % NOPS;
      mbar_wait(&stg_full[slot], (uint32_t)((it / NSTG) & 1));
      if (it >= NOPS) mbar_wait(&op_empty[s], (uint32_t)(((it / NOPS) - 1) & 1));
      const uint8_t* src = stg_base + slot * C::STG_SLOT;
      uint8_t* a_hi = smem + s * C::OP_STAGE;
      uint8_t* a_lo = a_hi + C::A_PART;
      uint8_t* b_hi = a_hi + 2 * C::A_PART;
      uint8_t* b_lo = b_hi + C::B_PART;
#pragma unroll 4
      for (int i = ctid; i < GT; i += CONV_THREADS) {
        const int c = i / TF, f = i - c * TF;
        const float4 x0 = *reinterpret_cast<const float4*>(src + f * C::PITCH_G + c * 32);
        const float4 x1 = *reinterpret_cast<const float4*>(src + f * C::PITCH_G + c * 32 + 16);
        const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint4 h, l; split8(v, h, l);
        const uint32_t off = (uint32_t)c * C::SBO_A + (uint32_t)f * 16;
        *reinterpret_cast<uint4*>(a_hi + off) = h;
        *reinterpret_cast<uint4*>(a_lo + off) = l;
      }
      const uint8_t* srcx = src + C::STG_G;
#pragma unroll 4
      for (int i = ctid; i < XT; i += CONV_THREADS) {
        const int c = i / XR, f = i - c * XR;
        const float4 x0 = *reinterpret_cast<const float4*>(srcx + f * C::PITCH_X + c * 32);
        const float4 x1 = *reinterpret_cast<const float4*>(srcx + f * C::PITCH_X + c * 32 + 16);
        const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint4 h, l; split8(v, h, l);
        const uint32_t off = (uint32_t)c * C::SBO_B + (uint32_t)f * 16;
        *reinterpret_cast<uint4*>(b_hi + off) = h;
        *reinterpret_cast<uint4*>(b_lo + off) = l;
      }
      mbar_arrive(&stg_empty[slot]);            // all reads of the slot are done (the stores above depend on them)
      fence_async_smem();                       // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&op_full[s]);
    }
  } else {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_mn(128, N);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % NOPS;
      mbar_wait(&op_full[s], (uint32_t)((it / NOPS) & 1));
      fence_after_sync();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * C::OP_STAGE);
        // MN-major, no swizzle: LBO = stride between 8-frame K groups (128 B), SBO = stride between 8-wide MN chunks
        const uint64_t bh0 = smem_desc(sa + 2 * C::A_PART, 128, C::SBO_B), bl0 = smem_desc(sa + 2 * C::A_PART + C::B_PART, 128, C::SBO_B);
        const uint32_t acc = it ? 1u : 0u;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          const uint64_t ah = smem_desc(sa + mi * 16 * C::SBO_A, 128, C::SBO_A), al = smem_desc(sa + C::A_PART + mi * 16 * C::SBO_A, 128, C::SBO_A);
#pragma unroll
          for (int tap = 0; tap < NTAPS; ++tap) {
            const uint64_t bh = desc_advance(bh0, tap * 16), bl = desc_advance(bl0, tap * 16);
            const uint32_t d = tmem + (uint32_t)((mi * NTAPS + tap) * N);
            mma_bf16(d, ah, bh, idesc, acc);
            mma_bf16(d, ah, bl, idesc, 1u);
            mma_bf16(d, al, bh, idesc, 1u);
          }
        }
        mma_commit(&op_empty[s]);
        if (it == my_tiles - 1) mma_commit(&acc_bar);
      }
      __syncwarp();
    }
  }

  // ================= flush: TMEM -> transposed through shared memory -> fp32 atomics (coalesced when s_n == 1) =================
  if (my_tiles > 0 && warp < 8) {
    mbar_wait(&acc_bar, 0);
    fence_after_sync();
    const int q = warp & 3, half = warp >> 2;          // two warps per TMEM lane quarter alternate 16-column groups
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 17);        // operand stages are idle once acc_bar fired
    const bool vec_ok = p.s_n == 1 && (p.s_m & 3) == 0 && (p.s_tap & 3) == 0 && (n0 & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int c0 = half * 16; c0 < C::ACC; c0 += 32) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) tr[lane * 17 + i] = v[i];
      __syncwarp();
      const int blk = c0 / N, mi = blk / NTAPS, tap = blk - mi * NTAPS;
      if (vec_ok) {
        // 16-byte vector reductions (red.global.add.v4.f32): a lane adds 4 consecutive columns of one row
        const int nb = c0 - blk * N + (lane & 3) * 4;
        if (nb < nv) {
          float* dst = p.out + (long)(n0 + nb) + (long)tap * p.s_tap;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int r = rr * 8 + (lane >> 2), m = mi * 128 + q * 32 + r;
            const float* t4 = tr + r * 17 + (lane & 3) * 4;
            if (m < mv) red_add_v4(dst + (long)(m0 + m) * p.s_m, t4[0], t4[1], t4[2], t4[3]);
          }
        }
      } else {
        const int nb = c0 - blk * N + (lane & 15);
        if (nb < nv) {
          float* dst = p.out + (long)(n0 + nb) * p.s_n + (long)tap * p.s_tap;
#pragma unroll 4
          for (int rr = 0; rr < 16; ++rr) {
            const int m = mi * 128 + q * 32 + 2 * rr + (lane >> 4);
            if (m < mv) atomicAdd(dst + (long)(m0 + m) * p.s_m, tr[(2 * rr + (lane >> 4)) * 17 + (lane & 15)]);
          }
        }
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TCOLS);
}

int sm_count_w() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int MT, int NCH, int NTAPS>
int launch(const WgradParams& p, cudaStream_t st) {
  using C = Cfg<MT, NCH, NTAPS>;
  SLU_SMEM_ONCE((wgrad_tc_kernel<MT, NCH, NTAPS>), C::TOTAL);
  const int m_groups = (p.m_valid + C::MC - 1) / C::MC, n_groups = (p.n_valid + C::N - 1) / C::N;
  const long n_tiles = (long)p.B * p.tiles_per_utt;
  int gx = sm_count_w() / (m_groups * n_groups);        // CTAs per output block: fill the SMs ...
  if (gx > n_tiles / 8) gx = (int)(n_tiles / 8);        // ... but give every CTA >= 8 frame tiles per (vector-atomic) flush
  if (gx < 1) gx = 1;
  wgrad_tc_kernel<MT, NCH, NTAPS><<<dim3(gx, m_groups, n_groups), THREADS, C::TOTAL, st>>>(p);
  return (int)cudaGetLastError();
}

int run(const WgradParams& p, int taps, cudaStream_t st) {
  const int M = p.m_valid, N = p.n_valid;
  if (taps == 5) {
    if (M > 128) return (int)cudaErrorInvalidValue;
    if (N <= 64) return launch<1, 8, 5>(p, st);
    if (N <= 80) return launch<1, 10, 5>(p, st);
    return (int)cudaErrorInvalidValue;
  }
  if (N <= 64) return M <= 128 ? launch<1, 8, 1>(p, st) : (M <= 256 ? launch<2, 8, 1>(p, st) : launch<3, 8, 1>(p, st));
  if (N <= 128) return M <= 128 ? launch<1, 16, 1>(p, st) : (M <= 256 ? launch<2, 16, 1>(p, st) : launch<3, 16, 1>(p, st));
  return M <= 128 ? launch<1, 32, 1>(p, st) : launch<2, 32, 1>(p, st);       // N > 256: blockIdx.z walks 256-column groups
}

bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

}  // namespace

// out[m*s_m + n*s_n + tap*s_tap] += sum_{b,t} G[(b*T+t)][m] * X[(b*T + t + shift0 + tap)*ldx + n]   (frames outside [0,T) read 0)
// with G rows [0, m_split) taken from G0 (pitch ldg0) and rows [m_split, M) from G1 (pitch ldg1); see include/slu_b200.h.
extern "C" int slu_wgrad2_tc(const float* G0, long ldg0, int m_split, const float* G1, long ldg1, int M, const float* X, long ldx, int N,
                             int B, int T, int taps, int shift0, float* out, long s_m, long s_n, long s_tap, void* stream) {
  if (M <= 0 || N <= 0 || B <= 0 || T <= 0 || (taps != 1 && taps != 5) || m_split < 0 || m_split > M) return (int)cudaErrorInvalidValue;
  if (taps > 1 && (shift0 < -4 || shift0 + taps - 1 > 4)) return (int)cudaErrorInvalidValue;
  if (m_split < M && !G1) return (int)cudaErrorInvalidValue;
  if (!aligned16(G0) || !aligned16(G1) || !aligned16(X) || (ldg0 & 3) || (ldg1 & 3) || (ldx & 3) || (M & 3) || (N & 3) || (m_split & 3))
    return (int)cudaErrorInvalidValue;
  WgradParams p;
  p.G0 = G0; p.ldg0 = ldg0; p.G1 = G1 ? G1 : G0; p.ldg1 = G1 ? ldg1 : ldg0; p.m_split = m_split;
  p.X = X; p.ldx = ldx; p.m_valid = M; p.n_valid = N; p.B = B; p.T = T; p.shift0 = shift0;
  p.tiles_per_utt = (T + TF - 1) / TF; p.out = out; p.s_m = s_m; p.s_n = s_n; p.s_tap = s_tap;
  return run(p, taps, (cudaStream_t)stream);
}

// Single-source form.
extern "C" int slu_wgrad_tc(const float* G, long ldg, int M, const float* X, long ldx, int N, int B, int T, int taps, int shift0,
                            float* out, long s_m, long s_n, long s_tap, void* stream) {
  return slu_wgrad2_tc(G, ldg, M, nullptr, 0, M, X, ldx, N, B, T, taps, shift0, out, s_m, s_n, s_tap, stream);
}
