// SincNet front end on tcgen05 tensor cores, "the waveform IS the matrix" formulation.
//
// With stride 80 and 401 = 5*80 + 1 taps, the padded waveform of one utterance viewed as a row-major [frames][80]
// matrix F (F[j][r] = xpad[80 j + r], xpad = x shifted by 200 zeros) gives
//     conv[t][c] = sum_{a=0..5} sum_{r<80} F[t+a][r] * W[c][80a + r]            (taps beyond 400 are zero)
// i.e. six accumulating K=80 MMAs over ONE staged tile whose row offset is the tap index.  The tile is staged once
// per 128 output frames as bf16 hi/lo in the K-major "chunk-column" image  byte(row, k) = (k/8)*LBO + row*16 + (k%8)*2,
// in which a row shift of `a` frames is just +16*a bytes on the descriptor start address: no im2col, no re-staging.
//   forward : D[128 frames x 80 filters] = sum_a  F(rows a..a+127) . W_a^T ; epilogue = abs + max-pool(2) + route bits
//   backward: dW[c][80a + r] = sum_t g0[t][c] * F[t+a][r]  -- the same image read as an MN-major B operand (K = frames),
//             the routed gradient g0 as an MN-major A operand; six [128 x 80] accumulators live in TMEM (480 columns)
//             across all tiles of a persistent CTA and are flushed once with fp32 atomics.
// Every product is bf16 hi*hi + hi*lo + lo*hi with fp32 accumulation (see tc05.cuh).
// Reference behaviour: models.py:108 (F.conv1d stride 80 pad 200), :163-168 (Abs), :205 (MaxPool1d(2, ceil)).
#include "common.cuh"
#include "tc05.cuh"

namespace {
using namespace tc05;

constexpr int THREADS = 256;
constexpr int TF = 128;                         // output frames per tile
constexpr int XROWS = TF + 8;                   // staged frame rows (5 halo rows, padded to a multiple of 8)
constexpr int KC = SLU_STRIDE / 8;              // 10 sample chunks per frame
constexpr uint32_t LBO_X = XROWS * 16 + 16;     // 2192: chunk-column stride of the waveform image
constexpr uint32_t X_PART = KC * LBO_X;         // 21920 B (one of hi / lo)
constexpr uint32_t LBO_W = SLU_NFILT * 16 + 16; // 1296: chunk-column stride of one tap's filter tile [80 n][80 k]
constexpr uint32_t W_PART = KC * LBO_W;         // 12960 B

// Instruction descriptor with selectable operand majors (bit 15: A is MN-major, bit 16: B is MN-major).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return idesc_bf16_f32(M, N) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
}

// Stage the waveform slice of one tile: rows r = 0..XROWS-1 are frames t0 + r, 80 samples each, starting at sample
// 80*t0 - 200 of utterance xb (zero outside [0, T)).
__device__ __forceinline__ void stage_wave_image(uint8_t* hi, uint8_t* lo, const float* xb, int t0, int T, int tid) {
  for (int base = 0; base < XROWS * KC; base += 4 * THREADS) {       // batches of 4 tasks: 8 loads in flight per thread
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int task = base + u * THREADS + tid;
      const int r = task / KC, kc = task - r * KC;
      const int idx0 = SLU_STRIDE * (t0 + r) + kc * 8 - SLU_PAD;
      const float* p = xb + idx0;
      const bool on = task < XROWS * KC;
      if (on && idx0 >= 0 && idx0 + 8 <= T && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = (on && idx0 + i >= 0 && idx0 + i < T) ? __ldg(p + i) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int task = base + u * THREADS + tid;
      if (task < XROWS * KC) {
        const int r = task / KC, kc = task - r * KC;
        uint4 h, l; split8(v[u], h, l);
        const uint32_t off = (uint32_t)kc * LBO_X + (uint32_t)r * 16;
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t FWD_SMEM = 2 * X_PART + 2 * 2 * W_PART;      // waveform image (hi, lo) + 2-slot ring of filter taps
constexpr int WKC = 96 / 8;                                     // k-chunks per tap in the pre-split bank image (Kp = 96)

// Filter taps arrive through a 2-slot TMA ring (the pre-split bank image is k-chunk-major: one bulk copy per 8-sample chunk
// of all 80 filters), issued by one elected thread of warp 1 from the very start of the CTA -- they land while all warps
// stage the waveform image; warp 0 issues the MMAs.  No block-wide barrier inside the tap loop.
// GRAD = false: forward (epilogue = abs + max-pool + route bits).
// GRAD = true : cut-off gradients.  `wimg` holds gridDim.z stacked banks (the Jacobian banks of slu_sinc_filters_jac, image rows
//               z*80 .. z*80+79); the epilogue multiplies the convolution by the routed output gradient (read from gy_in / route)
//               and reduces over frames: dsum[z*80 + c] += sum_t g0[t][c] * conv_z[t][c]   (fp64 atomics, one per filter per CTA).
template <bool GRAD>
__global__ void __launch_bounds__(THREADS, 2)
sincconv_fwd_tc_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ wimg, int T, int L0, int L1,
                       float* __restrict__ out, uint8_t* __restrict__ route, const float* __restrict__ gy_in,
                       double* __restrict__ dsum) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[2], empty_bar[2], acc_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * TF;
  uint8_t* x_hi = smem; uint8_t* x_lo = smem + X_PART;
  uint8_t* w_ring = smem + 2 * X_PART;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = idesc_bf16(128, SLU_NFILT, false, false);

  // tap `tap` of the bank -> ring slot tap & 1: 10 chunks x (hi, lo), 1280 B each.  Tap 5 has a single non-zero sample
  // (k = 0): one K=16 step = chunks 0 and 1 (chunk 1 is all zero in the image).
  auto tma_tap = [&](int tap) {
    const int slot = tap & 1, nch = tap == 5 ? 2 : KC;
    uint8_t* w_hi = w_ring + slot * 2 * W_PART;
    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(2 * nch * SLU_NFILT * 16));
    const int n_img = SLU_NFILT * (int)gridDim.z;              // rows of the (stacked) bank image
    for (int part = 0; part < 2; ++part)
      for (int kc = 0; kc < nch; ++kc) {
        const size_t e = ((((size_t)part * 6 + tap) * WKC + kc) * n_img + (size_t)blockIdx.z * SLU_NFILT) * 8;
        tma_load_1d(w_hi + part * W_PART + kc * LBO_W, wimg + e, SLU_NFILT * 16, &full_bar[slot]);
      }
  };
  if (warp == 1 && elect_one()) { tma_tap(0); tma_tap(1); }
  __syncwarp();

  stage_wave_image(x_hi, x_lo, x + (size_t)b * T, t0, T, tid);
  fence_async_smem();
  __syncthreads();

  if (warp == 1) {
    for (int tap = 2; tap < 6; ++tap) {
      mbar_wait(&empty_bar[tap & 1], (uint32_t)(((tap >> 1) - 1) & 1));      // the MMAs of tap - 2 have read the slot
      if (elect_one()) tma_tap(tap);
      __syncwarp();
    }
  } else if (warp == 0) {
    for (int tap = 0; tap < 6; ++tap) {
      const int slot = tap & 1;
      mbar_wait(&full_bar[slot], (uint32_t)((tap >> 1) & 1));
      fence_after_sync();
      if (elect_one()) {
        uint8_t* w_hi = w_ring + slot * 2 * W_PART; uint8_t* w_lo = w_hi + W_PART;
        // rows tap .. tap+127 of the staged image: start address + 16 B per row
        const uint64_t ah0 = smem_desc(smem_u32(x_hi) + tap * 16, LBO_X, 128), al0 = smem_desc(smem_u32(x_lo) + tap * 16, LBO_X, 128);
        const uint64_t bh0 = smem_desc(smem_u32(w_hi), LBO_W, 128), bl0 = smem_desc(smem_u32(w_lo), LBO_W, 128);
        const int nk = tap == 5 ? 1 : SLU_STRIDE / 16;
#pragma unroll 1
        for (int kk = 0; kk < nk; ++kk) {
          const uint64_t ah = desc_advance(ah0, kk * 2 * LBO_X), al = desc_advance(al0, kk * 2 * LBO_X);
          const uint64_t bh = desc_advance(bh0, kk * 2 * LBO_W), bl = desc_advance(bl0, kk * 2 * LBO_W);
          mma_bf16(tmem, ah, bh, idesc, (tap | kk) ? 1u : 0u);
          mma_bf16(tmem, ah, bl, idesc, 1u);
          mma_bf16(tmem, al, bh, idesc, 1u);
        }
        mma_commit(&empty_bar[slot]);
        if (tap == 5) mma_commit(&acc_bar);
      }
      __syncwarp();
    }
  }

  mbar_wait(&acc_bar, 0);
  fence_after_sync();
  if (GRAD) {
    // ---- epilogue (cut-off gradients): conv[t][c] * routed gradient g0[t][c], summed over the 128 frames of the tile
    const int q = warp & 3, half = warp >> 2;
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 17);    // image buffers are free once acc_bar fired
    double* part = reinterpret_cast<double*>(smem + 8 * 32 * 17 * 4); // [4 quarters][80] partial sums
    const int t = t0 + q * 32 + lane;                                 // this thread's frame (TMEM lane)
    const bool t_on = t < L0;
    const size_t o = ((size_t)b * L1 + (t_on ? (t >> 1) : 0)) * SLU_NFILT;
    for (int c0 = half * 48; c0 < (half ? SLU_NFILT : 48); c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      uint4 rt = make_uint4(0, 0, 0, 0);
      float4 g4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t_on) {
        rt = __ldg(reinterpret_cast<const uint4*>(route + o + c0));
#pragma unroll
        for (int i = 0; i < 4; ++i) g4[i] = __ldg(reinterpret_cast<const float4*>(gy_in + o + c0) + i);
      }
      const uint32_t rw[4] = {rt.x, rt.y, rt.z, rt.w};
      const float gv[16] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w,
                            g4[2].x, g4[2].y, g4[2].z, g4[2].w, g4[3].x, g4[3].y, g4[3].z, g4[3].w};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t rb = (rw[i >> 2] >> (8 * (i & 3))) & 0xffu;
        const bool take = t_on && ((rb & 1u) == (uint32_t)(t & 1)) && !(rb & 4u);   // this frame won the pair and |.| has a gradient
        tr[lane * 17 + i] = take ? ((rb & 2u) ? -gv[i] : gv[i]) * v[i] : 0.f;
      }
      __syncwarp();
      if (lane < 16) {                                                // lane i sums column c0 + i over this warp's 32 frames
        double s = 0.0;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) s += (double)tr[r * 17 + lane];
        part[q * SLU_NFILT + c0 + lane] = s;
      }
      __syncwarp();
    }
    __syncthreads();
    if (tid < SLU_NFILT)
      atomicAdd(dsum + blockIdx.z * SLU_NFILT + tid, (part[tid] + part[SLU_NFILT + tid]) + (part[2 * SLU_NFILT + tid] + part[3 * SLU_NFILT + tid]));
  } else
  // ---- epilogue: |.| + max over frame pairs + route bits, coalesced along the 80 filters
  {
    const int q = warp & 3, half = warp >> 2;                         // TMEM lane quarter; columns [0,48) / [48,80) in 16-col steps
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 17);    // image buffers are free once acc_bar fired
    for (int c0 = half * 48; c0 < (half ? SLU_NFILT : 48); c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) tr[lane * 17 + i] = v[i];
      __syncwarp();
      // lane -> (pair jp = lane / 2 of this warp's 16 pairs, 8 of the 16 columns)
      const int jp = lane >> 1, cb = (lane & 1) * 8;
      const int t = t0 + q * 32 + 2 * jp;
      if (t < L0) {
        const bool has1 = t + 1 < L0;
        const size_t o = ((size_t)b * L1 + (t >> 1)) * SLU_NFILT + c0 + cb;
        float r_out[8]; uint8_t r_rt[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v0 = tr[(2 * jp) * 17 + cb + i], v1 = tr[(2 * jp + 1) * 17 + cb + i];
          const float a0 = fabsf(v0), a1 = has1 ? fabsf(v1) : -1.f;
          const int sel = a1 > a0 ? 1 : 0;
          const float vs = sel ? v1 : v0;
          r_out[i] = sel ? a1 : a0;
          r_rt[i] = (uint8_t)(sel | ((vs < 0.f) ? 2 : 0) | ((vs == 0.f) ? 4 : 0));
        }
        *reinterpret_cast<float4*>(out + o) = make_float4(r_out[0], r_out[1], r_out[2], r_out[3]);
        *reinterpret_cast<float4*>(out + o + 4) = make_float4(r_out[4], r_out[5], r_out[6], r_out[7]);
        if (route) {
          uint2 pk;
          pk.x = r_rt[0] | (r_rt[1] << 8) | (r_rt[2] << 16) | ((uint32_t)r_rt[3] << 24);
          pk.y = r_rt[4] | (r_rt[5] << 8) | (r_rt[6] << 16) | ((uint32_t)r_rt[7] << 24);
          *reinterpret_cast<uint2*>(route + o) = pk;
        }
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------------------------
// persistent, warp-specialised variant of the kernel above (same math, same operand images)
// ---------------------------------------------------------------------------------------------------------------
// One CTA per SM walks the 128-frame tiles (tile = blockIdx.x, + gridDim.x, ...).  14 warps:
//   warps 0-7   STAGERS   waveform slice -> bf16 hi/lo image, two image slots: tile i+1 is staged while tile i is multiplied
//   warp  8     MMA       six taps x (5 K-steps x 3 split passes) into one of TWO TMEM accumulators
//   warp  9     BANK      streams the pre-split filter taps from L2 through a 4-slot TMA ring (tap g of the CTA -> slot g & 3)
//   warps 10-13 EPILOGUE  tcgen05.ld of accumulator i (abs / max-pool / route stores, or the routed-gradient dot product) while
//                         the MMAs of tile i+1 run into the other accumulator
// so per tile the SM pays max(staging, MMA, epilogue) instead of their sum; the MMA phase (78 M128 x N80 x K16 instructions,
// ~3 300 cycles) is the floor.  All hand-offs are mbarriers (bounded spins: a protocol bug traps instead of hanging).
// The filter taps use their own operand image here: [bank][tap][hi | lo][10 k-chunks][80 filters][8] bf16, i.e. ONE contiguous
// 25 600-byte block per tap, fetched with ONE bulk copy (the per-tile kernel's image needs 20 copies per tap, and bulk copies
// are operation-rate limited: 120 per tile made the bank loader the bottleneck of the first persistent version).
constexpr int P_STAGE_WARPS = 8, P_WARPS = 14, P_THREADS = P_WARPS * 32;
constexpr int P_BANK_SLOTS = 4;
constexpr uint32_t P_IMG = 2 * X_PART;                                  // one image slot (hi, lo)
constexpr uint32_t P_LBO_W = SLU_NFILT * 16;                            // 1280: unpadded chunk stride (TMA writes it, no thread stores)
constexpr uint32_t P_W_PART = KC * P_LBO_W;                             // 12 800 B
constexpr uint32_t P_BANK = 2 * P_W_PART;                               // one bank slot = one tap (hi, lo) = one bulk copy

// fp32 bank(s) W[nb][80][401] -> the per-tap image above (taps beyond 400 are zero)
__global__ void __launch_bounds__(256) sinc_bank_image_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ img) {
  const int bank = blockIdx.x / 6, tap = blockIdx.x % 6;
  __nv_bfloat16* dst = img + (size_t)blockIdx.x * (P_BANK / 2);
  for (int i = threadIdx.x; i < KC * SLU_NFILT * 8; i += 256) {
    const int e = i & 7, n = (i >> 3) % SLU_NFILT, kc = i / (8 * SLU_NFILT);
    const int k = SLU_STRIDE * tap + kc * 8 + e;
    const float v = k < SLU_NTAPS ? W[((size_t)bank * SLU_NFILT + n) * SLU_NTAPS + k] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    dst[i] = h;
    dst[P_W_PART / 2 + i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
constexpr uint32_t P_TR = 4 * 32 * 17 * 4;                              // epilogue transpose buffers
constexpr uint32_t P_SMEM = 2 * P_IMG + P_BANK_SLOTS * P_BANK + P_TR;   // 200 064 B

// Staging split in two so that a stager's global loads for tile i+1 are in flight while it waits for the image slot (the MMAs
// of tile i-1): prefetch = issue the loads of this thread's 6 (frame row, 8-sample chunk) tasks into registers, commit = split
// into bf16 hi / lo and store the 16-byte operand chunks.  (Loading only after the slot is free exposed one HBM round trip per
// tile: the first persistent version ran at 2x its MMA floor with the stagers stalled on the loads.)
constexpr int P_TASKS = (XROWS * KC + P_STAGE_WARPS * 32 - 1) / (P_STAGE_WARPS * 32);     // 6

// Developer tool (tools/sinc_trace.py): when set, CTA (0,0) records clock64() at the hand-off points of its first 16 tiles:
// trace[tile][0..7] = stager: slot free, image committed | MMA: image ready, accumulator free, last tap issued |
//                     epilogue (quarter 0): accumulator full, tile drained | bank loader: tap 5 of the tile issued
__device__ long long* g_sinc_trace = nullptr;
#ifdef SLU_KERNEL_DEBUG      // SLU_KERNEL_DEBUG=1 python __graft_entry__.py; the default build carries no trace branches
#define STRACE(tile, ev) do { if (trace && (tile) < 16) trace[(tile) * 8 + (ev)] = clock64(); } while (0)
#else
#define STRACE(tile, ev) do { } while (0)
#endif

__device__ __forceinline__ void stage_prefetch(float (&v)[P_TASKS][8], const float* xb, int t0, int T, int tid) {
#pragma unroll
  for (int u = 0; u < P_TASKS; ++u) {
    const int task = u * (P_STAGE_WARPS * 32) + tid;
    const int r = task / KC, kc = task - r * KC;
    const int idx0 = SLU_STRIDE * (t0 + r) + kc * 8 - SLU_PAD;
    const float* p = xb + idx0;
    const bool on = task < XROWS * KC;
    if (on && idx0 >= 0 && idx0 + 8 <= T && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u][i] = (on && idx0 + i >= 0 && idx0 + i < T) ? __ldg(p + i) : 0.f;
    }
  }
}

__device__ __forceinline__ void stage_commit(uint8_t* hi, uint8_t* lo, const float (&v)[P_TASKS][8], int tid) {
#pragma unroll
  for (int u = 0; u < P_TASKS; ++u) {
    const int task = u * (P_STAGE_WARPS * 32) + tid;
    if (task < XROWS * KC) {
      const int r = task / KC, kc = task - r * KC;
      uint4 h, l; split8(v[u], h, l);
      const uint32_t off = (uint32_t)kc * LBO_X + (uint32_t)r * 16;
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(P_THREADS, 1)
sincconv_tc_persistent_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ wimg, int T, int L0, int L1, int tiles_per_utt,
                              int n_tiles, float* __restrict__ out, uint8_t* __restrict__ route, const float* __restrict__ gy_in,
                              double* __restrict__ dsum) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t img_full[2], img_empty[2], acc_full[2], acc_empty[2], bank_full[P_BANK_SLOTS], bank_empty[P_BANK_SLOTS];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  uint8_t* bank_ring = smem + 2 * P_IMG;
  float* tr_all = reinterpret_cast<float*>(smem + 2 * P_IMG + P_BANK_SLOTS * P_BANK);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&img_full[s], P_STAGE_WARPS); mbar_init(&img_empty[s], 1); mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4);
    }
    for (int s = 0; s < P_BANK_SLOTS; ++s) { mbar_init(&bank_full[s], 1); mbar_init(&bank_empty[s], 1); }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int n_mine = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  long long* trace = (g_sinc_trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? g_sinc_trace : nullptr;

  if (warp < P_STAGE_WARPS) {
    // ---- stagers ------------------------------------------------------------------------------------------------------
    float v[P_TASKS][8];
    auto prefetch = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TF;
      stage_prefetch(v, x + (size_t)b * T, t0, T, tid);
    };
    if (n_mine > 0) prefetch(0);
    for (int it = 0; it < n_mine; ++it) {
      const int slot = it & 1;
      if (it >= 2) mbar_wait(&img_empty[slot], (uint32_t)(((it >> 1) - 1) & 1));      // the MMAs of tile it-2 have read this slot
      if (warp == 0) STRACE(it, 0);
      uint8_t* hi = smem + slot * P_IMG;
      stage_commit(hi, hi + X_PART, v, tid);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&img_full[slot]);
      if (warp == 0) STRACE(it, 1);
      if (it + 1 < n_mine) prefetch(it + 1);                                          // in flight across the next slot wait
    }
  } else if (warp == P_STAGE_WARPS) {
    // ---- MMA issuer ---------------------------------------------------------------------------------------------------
    // Measured (tools/sinc_trace.py): the 78 MMAs of a tile take ~4 900 cycles = 63 cycles each although an M128 x N80 x K16
    // MMA is ~42 cycles of tensor work: both operands come from shared memory (4 KB + 2.5 KB per MMA) and 128 B/cycle of
    // shared-memory bandwidth is the limit (507 KB per tile).  Splitting N into 64 + 16 made it worse (A is read twice as often).
    const uint32_t idesc = idesc_bf16(128, SLU_NFILT, false, false);
    for (int it = 0; it < n_mine; ++it) {
      const int slot = it & 1;
      mbar_wait(&img_full[slot], (uint32_t)((it >> 1) & 1));
      STRACE(it, 2);
      if (it >= 2) mbar_wait(&acc_empty[slot], (uint32_t)(((it >> 1) - 1) & 1));      // the epilogue has drained this accumulator
      STRACE(it, 3);
      fence_after_sync();
      const uint32_t x_hi = smem_u32(smem + slot * P_IMG), x_lo = x_hi + X_PART;
      const uint32_t dacc = tmem + (uint32_t)slot * 128u;
      for (int tap = 0; tap < 6; ++tap) {
        const int g = it * 6 + tap, bs = g & (P_BANK_SLOTS - 1);
        mbar_wait(&bank_full[bs], (uint32_t)((g / P_BANK_SLOTS) & 1));
        fence_after_sync();
        if (elect_one()) {
          const uint32_t w_hi = smem_u32(bank_ring + bs * P_BANK), w_lo = w_hi + P_W_PART;
          const uint64_t ah0 = smem_desc(x_hi + tap * 16, LBO_X, 128), al0 = smem_desc(x_lo + tap * 16, LBO_X, 128);
          const uint64_t bh0 = smem_desc(w_hi, P_LBO_W, 128), bl0 = smem_desc(w_lo, P_LBO_W, 128);
          const int nk = tap == 5 ? 1 : SLU_STRIDE / 16;
#pragma unroll 1
          for (int kk = 0; kk < nk; ++kk) {
            const uint64_t ah = desc_advance(ah0, kk * 2 * LBO_X), al = desc_advance(al0, kk * 2 * LBO_X);
            const uint64_t bh = desc_advance(bh0, kk * 2 * P_LBO_W), bl = desc_advance(bl0, kk * 2 * P_LBO_W);
            mma_bf16(dacc, ah, bh, idesc, (tap | kk) ? 1u : 0u);
            mma_bf16(dacc, ah, bl, idesc, 1u);
            mma_bf16(dacc, al, bh, idesc, 1u);
          }
          mma_commit(&bank_empty[bs]);
          if (tap == 5) { mma_commit(&img_empty[slot]); mma_commit(&acc_full[slot]); }
        }
        __syncwarp();
      }
      STRACE(it, 4);
    }
  } else if (warp == P_STAGE_WARPS + 1) {
    // ---- filter-bank loader (TMA ring) --------------------------------------------------------------------------------
    for (int g = 0; g < n_mine * 6; ++g) {
      const int bs = g & (P_BANK_SLOTS - 1), tap = g % 6;
      if (g >= P_BANK_SLOTS) mbar_wait(&bank_empty[bs], (uint32_t)(((g / P_BANK_SLOTS) - 1) & 1));
      if (elect_one()) {                                                // one tap = one contiguous block of the per-tap image
        mbar_arrive_expect_tx(&bank_full[bs], P_BANK);
        tma_load_1d(bank_ring + bs * P_BANK, wimg + ((size_t)blockIdx.y * 6 + tap) * (P_BANK / 2), P_BANK, &bank_full[bs]);
      }
      if (tap == 5) STRACE(g / 6, 7);
      __syncwarp();
    }
  } else {
    // ---- epilogue -----------------------------------------------------------------------------------------------------
    const int q = warp & 3;                                             // TMEM lane quarter of this warp
    float* tr = tr_all + (warp - P_STAGE_WARPS - 2) * (32 * 17);
    float gacc[GRAD ? SLU_NFILT : 1];                                   // GRAD: per-thread (= per frame row) sums over its tiles
#pragma unroll
    for (int i = 0; i < (GRAD ? SLU_NFILT : 1); ++i) gacc[i] = 0.f;
    for (int it = 0; it < n_mine; ++it) {
      const int tile = blockIdx.x + it * gridDim.x, slot = it & 1;
      const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TF;
      mbar_wait(&acc_full[slot], (uint32_t)((it >> 1) & 1));
      if (q == 0) STRACE(it, 5);
      fence_after_sync();
      const uint32_t acc = tmem + (uint32_t)slot * 128u + ((uint32_t)(q * 32) << 16);
      if (GRAD) {
        const int t = t0 + q * 32 + lane;
        const bool t_on = t < L0;
        const size_t o = ((size_t)b * L1 + (t_on ? (t >> 1) : 0)) * SLU_NFILT;
#pragma unroll
        for (int c0 = 0; c0 < SLU_NFILT; c0 += 16) {
          float v[16];
          tmem_ld16(acc + c0, v);
          tmem_ld_wait();
          uint4 rt = make_uint4(0, 0, 0, 0);
          float4 g4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t_on) {
            rt = __ldg(reinterpret_cast<const uint4*>(route + o + c0));
#pragma unroll
            for (int i = 0; i < 4; ++i) g4[i] = __ldg(reinterpret_cast<const float4*>(gy_in + o + c0) + i);
          }
          const uint32_t rw[4] = {rt.x, rt.y, rt.z, rt.w};
          const float gv[16] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w,
                                g4[2].x, g4[2].y, g4[2].z, g4[2].w, g4[3].x, g4[3].y, g4[3].z, g4[3].w};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t rb = (rw[i >> 2] >> (8 * (i & 3))) & 0xffu;
            const bool take = t_on && ((rb & 1u) == (uint32_t)(t & 1)) && !(rb & 4u);
            gacc[c0 + i] += take ? ((rb & 2u) ? -gv[i] : gv[i]) * v[i] : 0.f;
          }
        }
      } else {
        for (int c0 = 0; c0 < SLU_NFILT; c0 += 16) {
          float v[16];
          tmem_ld16(acc + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) tr[lane * 17 + i] = v[i];
          __syncwarp();
          const int jp = lane >> 1, cb = (lane & 1) * 8;                // lane -> (frame pair jp of this warp's 16 pairs, 8 of the 16 columns)
          const int t = t0 + q * 32 + 2 * jp;
          if (t < L0) {
            const bool has1 = t + 1 < L0;
            const size_t o = ((size_t)b * L1 + (t >> 1)) * SLU_NFILT + c0 + cb;
            float r_out[8]; uint8_t r_rt[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float v0 = tr[(2 * jp) * 17 + cb + i], v1 = tr[(2 * jp + 1) * 17 + cb + i];
              const float a0 = fabsf(v0), a1 = has1 ? fabsf(v1) : -1.f;
              const int sel = a1 > a0 ? 1 : 0;
              const float vs = sel ? v1 : v0;
              r_out[i] = sel ? a1 : a0;
              r_rt[i] = (uint8_t)(sel | ((vs < 0.f) ? 2 : 0) | ((vs == 0.f) ? 4 : 0));
            }
            *reinterpret_cast<float4*>(out + o) = make_float4(r_out[0], r_out[1], r_out[2], r_out[3]);
            *reinterpret_cast<float4*>(out + o + 4) = make_float4(r_out[4], r_out[5], r_out[6], r_out[7]);
            if (route) {
              uint2 pk;
              pk.x = r_rt[0] | (r_rt[1] << 8) | (r_rt[2] << 16) | ((uint32_t)r_rt[3] << 24);
              pk.y = r_rt[4] | (r_rt[5] << 8) | (r_rt[6] << 16) | ((uint32_t)r_rt[7] << 24);
              *reinterpret_cast<uint2*>(route + o) = pk;
            }
          }
          __syncwarp();
        }
      }
      fence_before_sync();                                              // this thread's tcgen05.ld are done (wait::ld above)
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
      if (q == 0) STRACE(it, 6);
    }
    if (GRAD) {
      // per-filter sums over this warp's 32 frame rows, then one fp64 atomic per filter per warp
      for (int c0 = 0; c0 < SLU_NFILT; c0 += 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) tr[lane * 17 + i] = gacc[c0 + i];
        __syncwarp();
        if (lane < 16) {
          double s2 = 0.0;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) s2 += (double)tr[r * 17 + lane];
          atomicAdd(dsum + blockIdx.y * SLU_NFILT + c0 + lane, s2);
        }
        __syncwarp();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// backward (filter gradient)
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t SBO_G = TF * 16 + 16;                   // 2064: stride between 8-filter chunks of the routed-gradient image
constexpr uint32_t G_PART = 16 * SBO_G;                    // 16 chunks = 128 filter rows (80 real), 33024 B
constexpr uint32_t BWD_STAGE = 2 * X_PART + 2 * G_PART;    // 109888 B
constexpr uint32_t BWD_SMEM = 2 * BWD_STAGE;               // two stages

__global__ void __launch_bounds__(THREADS, 1)
sincconv_bwd_tc_kernel(const float* __restrict__ x, const float* __restrict__ gy, const uint8_t* __restrict__ route, int B, int T,
                       int L0, int L1, int tiles_per_utt, float* __restrict__ dW) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t empty_bar[2], acc_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int n_tiles = B * tiles_per_utt;

  if (tid == 0) { mbar_init(&empty_bar[0], 1); mbar_init(&empty_bar[1], 1); mbar_init(&acc_bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  // filter chunks 10..15 (rows 80..127 of the M=128 tile) are never written: zero them once in both stages
  for (int s = 0; s < 2; ++s)
    for (int part = 0; part < 2; ++part) {
      uint8_t* g = smem + s * BWD_STAGE + 2 * X_PART + part * G_PART + KC * SBO_G;
      for (int i = tid * 16; i < (int)(6 * SBO_G); i += THREADS * 16) *reinterpret_cast<uint4*>(g + i) = make_uint4(0, 0, 0, 0);
    }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = idesc_bf16(128, SLU_NFILT, true, true);     // both operands MN-major: the reduction runs over frames

  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    if (it >= 2) mbar_wait(&empty_bar[s], (uint32_t)(((it >> 1) - 1) & 1));
    const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TF;
    uint8_t* st = smem + s * BWD_STAGE;
    uint8_t* x_hi = st; uint8_t* x_lo = st + X_PART; uint8_t* g_hi = st + 2 * X_PART; uint8_t* g_lo = g_hi + G_PART;
    stage_wave_image(x_hi, x_lo, x + (size_t)b * T, t0, T, tid);
    // routed gradient g0[t][c] (through max-pool and abs) for frames t0..t0+127, 8 filters per 16-byte chunk
    for (int base = 0; base < TF * KC; base += 5 * THREADS) {      // 1280 tasks = 5 per thread, loads batched
      float4 ga[5], gb[5]; uint2 rt[5]; bool okv[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int task = base + u * THREADS + tid;
        const int r = task / KC, cc = task - r * KC;            // frame row, filter chunk
        const int t = t0 + r;
        okv[u] = task < TF * KC && t < L0;
        const size_t o = okv[u] ? ((size_t)b * L1 + (t >> 1)) * SLU_NFILT + cc * 8 : 0;
        ga[u] = __ldg(reinterpret_cast<const float4*>(gy + o)); gb[u] = __ldg(reinterpret_cast<const float4*>(gy + o) + 1);
        rt[u] = __ldg(reinterpret_cast<const uint2*>(route + o));
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int task = base + u * THREADS + tid;
        if (task < TF * KC) {
          const int r = task / KC, cc = task - r * KC;
          const int t = t0 + r;
          const float gv[8] = {ga[u].x, ga[u].y, ga[u].z, ga[u].w, gb[u].x, gb[u].y, gb[u].z, gb[u].w};
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t rb = ((i < 4 ? rt[u].x : rt[u].y) >> (8 * (i & 3))) & 0xffu;
            const bool take = okv[u] && ((rb & 1u) == (uint32_t)(t & 1)) && !(rb & 4u);
            v[i] = take ? ((rb & 2u) ? -gv[i] : gv[i]) : 0.f;
          }
          uint4 h, l; split8(v, h, l);
          const uint32_t off = (uint32_t)cc * SBO_G + (uint32_t)r * 16;
          *reinterpret_cast<uint4*>(g_hi + off) = h;
          *reinterpret_cast<uint4*>(g_lo + off) = l;
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
        // MN-major no-swizzle descriptors: LBO = stride between 8-frame K groups (128 B), SBO = stride between 8-wide MN chunks
        const uint64_t ah0 = smem_desc(smem_u32(g_hi), 128, SBO_G), al0 = smem_desc(smem_u32(g_lo), 128, SBO_G);
        const uint64_t bh0 = smem_desc(smem_u32(x_hi), 128, LBO_X), bl0 = smem_desc(smem_u32(x_lo), 128, LBO_X);
        for (int kk = 0; kk < TF / 16; ++kk) {               // 16 frames per MMA
          const uint64_t ah = desc_advance(ah0, kk * 256), al = desc_advance(al0, kk * 256);
          const uint32_t acc = (it | kk) ? 1u : 0u;
#pragma unroll
          for (int tap = 0; tap < 6; ++tap) {                // frame shift = tap rows of 16 B
            const uint64_t bh = desc_advance(bh0, kk * 256 + tap * 16), bl = desc_advance(bl0, kk * 256 + tap * 16);
            const uint32_t d = tmem + tap * SLU_NFILT;
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
    // flush: D_tap[c][r] -> dW[c][80 tap + r]
    const int q = warp & 3, half = warp >> 2;
    const int c = q * 32 + lane;
    for (int tap = half * 3; tap < half * 3 + 3; ++tap) {
      for (int r0 = 0; r0 < SLU_NFILT; r0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + tap * SLU_NFILT + r0, v);
        tmem_ld_wait();
        if (c < SLU_NFILT) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = SLU_STRIDE * tap + r0 + i;
            if (k < SLU_NTAPS) atomicAdd(dW + c * SLU_NTAPS + k, v[i]);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

int slu_presplit_rows_cm(const float* W, long sn, long sk, long stap, int taps, int N, int K, int row_len, void* img, void* stream);

// 1 (default): the persistent warp-specialised kernel; 0: one CTA per tile (the round-1 kernel, kept for A/B measurements).
extern "C" int slu_debug_sinc_trace(long long* buf) {   // developer tool; buf = 128 zeroed int64 on the device, or NULL
  return (int)cudaMemcpyToSymbol(g_sinc_trace, &buf, sizeof(buf));
}
static int g_sinc_persistent = 1;
extern "C" int slu_set_sinc_persistent(int on) { g_sinc_persistent = on ? 1 : 0; return 0; }
static int sinc_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// Forward: out[B][L1][80] = maxpool2(|conv1d(x, W, stride 80, pad 200)|), route bits for the backward pass.
// `img` = scratch for the pre-split bank: 2*6*80*96 bf16 values.
extern "C" int slu_sincconv_fwd_tc(const float* x, const float* W, int B, int T, float* out, uint8_t* route, void* img, void* stream) {
  if (B <= 0 || T <= 0) return (int)cudaErrorInvalidValue;
  const int L0 = (T - 1) / SLU_STRIDE + 1, L1 = (L0 + 1) / 2;
  if (g_sinc_persistent) {
    sinc_bank_image_kernel<<<6, 256, 0, (cudaStream_t)stream>>>(W, (__nv_bfloat16*)img);
    SLU_SMEM_ONCE(sincconv_tc_persistent_kernel<false>, P_SMEM);
    const int tiles_per_utt = (L0 + TF - 1) / TF;
    const long n_tiles = (long)B * tiles_per_utt;
    if (n_tiles >= (1L << 31)) return SLU_ERR_TOO_LARGE;
    const int grid = (int)(n_tiles < sinc_sm_count() ? n_tiles : sinc_sm_count());
    sincconv_tc_persistent_kernel<false><<<grid, P_THREADS, P_SMEM, (cudaStream_t)stream>>>(x, (const __nv_bfloat16*)img, T, L0, L1, tiles_per_utt,
                                                                                          (int)n_tiles, out, route, nullptr, nullptr);
    SLU_CHECK_LAUNCH();
    return 0;
  }
  int e = slu_presplit_rows_cm(W, SLU_NTAPS, 1, SLU_STRIDE, 6, SLU_NFILT, SLU_STRIDE, SLU_NTAPS, img, stream);   // bank[c][80a + k], 0 beyond tap 400
  if (e) return e;
  SLU_SMEM_ONCE(sincconv_fwd_tc_kernel<false>, FWD_SMEM);
  dim3 grid((L0 + TF - 1) / TF, B);
  sincconv_fwd_tc_kernel<false><<<grid, THREADS, FWD_SMEM, (cudaStream_t)stream>>>(x, (const __nv_bfloat16*)img, T, L0, L1, out, route,
                                                                                   nullptr, nullptr);
  SLU_CHECK_LAUNCH();
  return 0;
}

// Cut-off gradients without a dW detour: d[0..79] += dL/d filt_b1, d[80..159] += dL/d filt_band (fp64, caller-zeroed) as two
// more strided convolutions of the waveform with the Jacobian banks J[2][80][401] (slu_sinc_filters_jac) dotted with the routed
// output gradient.  `img` = scratch for the pre-split banks: 2*6*160*96 bf16 values.
extern "C" int slu_sincconv_bwd_jac_tc(const float* x, const float* gy, const uint8_t* route, const float* J, int B, int T, double* d,
                                       void* img, void* stream) {
  if (B <= 0 || T <= 0) return (int)cudaErrorInvalidValue;
  const int L0 = (T - 1) / SLU_STRIDE + 1, L1 = (L0 + 1) / 2;
  if (g_sinc_persistent) {
    sinc_bank_image_kernel<<<12, 256, 0, (cudaStream_t)stream>>>(J, (__nv_bfloat16*)img);
    SLU_SMEM_ONCE(sincconv_tc_persistent_kernel<true>, P_SMEM);
    const int tiles_per_utt = (L0 + TF - 1) / TF;
    const long n_tiles = (long)B * tiles_per_utt;
    if (n_tiles >= (1L << 31)) return SLU_ERR_TOO_LARGE;
    const int half = sinc_sm_count() / 2;                                // the two Jacobian banks share the SMs
    const int gx = (int)(n_tiles < half ? n_tiles : half);
    sincconv_tc_persistent_kernel<true><<<dim3(gx, 2), P_THREADS, P_SMEM, (cudaStream_t)stream>>>(
        x, (const __nv_bfloat16*)img, T, L0, L1, tiles_per_utt, (int)n_tiles, nullptr, const_cast<uint8_t*>(route), gy, d);
    SLU_CHECK_LAUNCH();
    return 0;
  }
  int e = slu_presplit_rows_cm(J, SLU_NTAPS, 1, SLU_STRIDE, 6, 2 * SLU_NFILT, SLU_STRIDE, SLU_NTAPS, img, stream);   // both banks: 160 image rows
  if (e) return e;
  SLU_SMEM_ONCE(sincconv_fwd_tc_kernel<true>, FWD_SMEM);
  dim3 grid((L0 + TF - 1) / TF, B, 2);
  sincconv_fwd_tc_kernel<true><<<grid, THREADS, FWD_SMEM, (cudaStream_t)stream>>>(x, (const __nv_bfloat16*)img, T, L0, L1, nullptr,
                                                                                  const_cast<uint8_t*>(route), gy, d);
  SLU_CHECK_LAUNCH();
  return 0;
}

// Backward: dW[80][401] (zero-filled by the caller) += sum over frames of the routed gradient times the waveform.
extern "C" int slu_sincconv_bwd_tc(const float* x, const float* gy, const uint8_t* route, int B, int T, float* dW, void* stream) {
  if (B <= 0 || T <= 0) return (int)cudaErrorInvalidValue;
  const int L0 = (T - 1) / SLU_STRIDE + 1, L1 = (L0 + 1) / 2;
  const int tiles_per_utt = (L0 + TF - 1) / TF;
  SLU_SMEM_ONCE(sincconv_bwd_tc_kernel, BWD_SMEM);
  const long n_tiles = (long)B * tiles_per_utt;
  const int grid = (int)(n_tiles < 148 ? n_tiles : 148);
  sincconv_bwd_tc_kernel<<<grid, THREADS, BWD_SMEM, (cudaStream_t)stream>>>(x, gy, route, B, T, L0, L1, tiles_per_utt, dW);
  SLU_CHECK_LAUNCH();
  return 0;
}
