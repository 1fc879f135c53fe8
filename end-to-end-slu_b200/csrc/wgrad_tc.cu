// Weight-gradient ("reduction over frames") GEMM on tcgen05 with MN-major operands and TMEM-resident accumulators.
//
//   dW[tap][m][n] += sum over utterances b and frames t of  G[b][t][m] * X[b][t + shift0 + tap][n]      (0 outside [0, T))
//
// Both operands are stored frame-major in HBM (the reduction index is the slow one), which is exactly the MN-major
// operand form of tcgen05: a [8-wide column chunk][frame] image with 16 bytes per frame and chunk columns SBO apart.
// A persistent, warp-specialised CTA owns one [MT*128 x N] block of dW and walks 16-frame tiles of its share of the
// (utterance, tile) list:
//   warp 8     loader: ONE thread issues the tile's TMA tensor copies -- G and X are described to the copy engine as 3-D
//              (column, frame, utterance) fp32 tensors, a box is [16 (+ taps - 1) frames][<= 128 columns]; frames outside the
//              utterance (the tail tile, the +-1 / +-2 frame shifts) and columns beyond M / N are zero-filled by the hardware.
//              4-6 fp32 staging slots (as many as fit) keep ~100 KB per SM in flight; completion = mbarrier transaction bytes;
//   warps 0-7  converters (two groups of 4 warps taking alternate tiles): staging slot -> bf16 hi + lo MN-major images.  Lanes
//              walk the 8-column chunks of a frame row and read its two 16-byte halves in an order that depends on the chunk
//              ((c >> 2) & 1): conflict-free 16-byte shared-memory reads from the dense TMA boxes and conflict-free stores
//              (chunk columns are 16 bytes more than a multiple of 128 apart); 2-stage operand ring;
//   warp 9     MMA issuer: per tile hi*hi + hi*lo + lo*hi for every 128-row block of G and every tap (a tap is a frame
//              shift = +16 bytes on the B descriptor, so X is staged once for all taps); the MT * NTAPS accumulators
//              [128 x N] stay in tensor memory across ALL tiles of the CTA.
// One flush with fp32 atomics at the end.  G may come from two tensors (rows [0, m_split) from G0, the rest from G1):
// dW_hh = [dr, dz | dhn]^T . h_{t-+1} is ONE launch per direction.
// Used for: dW_ih = dgx^T.x, dW_hh (shift -1 / +1), Conv1d weight gradients (5 taps, shift -2).
// Replaces cuDNN's RNN / conv backward-weights (autograd of models.py:200, 232/262/686).
// Contract of the tensor copies: operands 16-byte aligned; ldg, ldx, M, N multiples of 4 floats; m_split a multiple of 128 (or = M).
// (Round 2 measured the previous 16-byte cp.async loader at 3.9-4.3 TB/s with everything else switched off: ~800 issue cycles and
// a 2 500-cycle latency per 32 KB tile with 3 slots in flight; profiles/r2_wgrad_*.)
#include "common.cuh"
#include "tc05.cuh"
#include "tmap.cuh"

namespace {
using namespace tc05;

constexpr int TF = 16;                         // frames per tile (= one K step of the MMA)
constexpr int NSTG_MAX = 6, NOPS = 2;          // fp32 staging slots (at most), bf16 operand stages
constexpr int CONV_GROUPS = 2;                 // converter groups of 4 warps taking alternate tiles (a tile's conversion is a latency chain)
constexpr int CONV_WARPS = 4 * CONV_GROUPS, LOAD_WARP = CONV_WARPS, MMA_WARP = CONV_WARPS + 1;
constexpr int THREADS = (MMA_WARP + 1) * 32;   // 320
constexpr int CONV_THREADS = 128;              // per group
constexpr int GB = 128;                        // G columns per staged block = one TMA box = one 128-row block of the MMA
constexpr uint32_t SMEM_BUDGET = 227 * 1024 - 512;       // per CTA, minus the static barriers and the alignment slack

struct WgradParams {
  CUtensorMap map_g0, map_g1, map_x;   // (column, frame, utterance) fp32 views: G rows [0, m_split), G rows [m_split, M), X
  int m_split;                          // first G column taken from map_g1 (>= m_valid: single source)
  int m_valid, n_valid;
  int B, T, shift0, tiles_per_utt;
  float* out; long s_m, s_n, s_tap;     // out[m*s_m + n*s_n + tap*s_tap] += D_tap[m][n]
  long long* trace;
  int zero;                              // 0 at run time, opaque to the compiler (see the converters' slot release)
  int dbg;                               // developer switch (slu_debug_wgrad_mode): 1 = no MMAs, 2 = no conversion, 4 = no flush
};
int g_dbg = 0;
// Ablation switches and the hand-off trace exist only in -DSLU_KERNEL_DEBUG builds (SLU_KERNEL_DEBUG=1 python __graft_entry__.py): even
// switched off they cost measurable time in the GEMM (gemm_tc.cu).  tools/wgrad_{only,rate,trace,check}.py need such a build.
#ifdef SLU_KERNEL_DEBUG
#define SLU_DBG(p) ((p).dbg)
#define SLU_TRACE(p) ((p).trace)
#else
#define SLU_DBG(p) 0
#define SLU_TRACE(p) ((long long*)nullptr)
#endif
long long* g_trace = nullptr;      // developer tool: CTA (0,0,0) records clock64() per tile: [tile][0] loader issue start, [1] issue end,
                                   // [2] converter saw the slot full, [3] converter done, [4] MMA warp saw operands, [5] after MMA issue

// fp32 x4 vector reduction to global memory (sm_90+): one L2 operation instead of four
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) { return idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16); }
__host__ __device__ constexpr uint32_t cmax(uint32_t a, uint32_t b) { return a > b ? a : b; }
__host__ __device__ constexpr uint32_t cmin(uint32_t a, uint32_t b) { return a < b ? a : b; }

template <int MT, int NCH, int NTAPS>
struct Cfg {
  static constexpr int MC = MT * 128, N = 8 * NCH;                   // staged G columns, X columns
  static constexpr int XR = TF + NTAPS - 1;                          // staged X rows
  static constexpr int XB = N <= 128 ? N : 128, NXB = N / XB;        // X columns per staged block (one TMA box), blocks
  static constexpr int XCH = XB / 8, XCHP = (XCH + 7) / 8 * 8;       // 8-column chunks per X block row; padded to whole quarter-warps
  static constexpr uint32_t G_BLK = TF * GB * 4, X_BLK = XR * XB * 4;             // dense fp32 boxes [frames][columns]
  static constexpr uint32_t STG_G = MT * G_BLK, STG_X = NXB * X_BLK, STG_SLOT = STG_G + STG_X;
  static constexpr uint32_t SBO_A = TF * 16 + 16, SBO_B = XR * 16 + 16;           // chunk-column strides of the bf16 images
  static constexpr uint32_t A_PART = (MC / 8) * SBO_A, B_PART = NCH * SBO_B;     // one of hi / lo
  static constexpr uint32_t OP_STAGE = 2 * A_PART + 2 * B_PART;
  static constexpr uint32_t PIPE = NOPS * OP_STAGE;
  static constexpr uint32_t FLUSH = 8 * 32 * 17 * 4;                              // transposers of the final flush (alias PIPE)
  static constexpr uint32_t STG_OFF = (cmax(PIPE, FLUSH) + 127u) & ~127u;
  static constexpr int NSTG = (int)cmin(NSTG_MAX, (SMEM_BUDGET - STG_OFF) / STG_SLOT) & ~1;    // even: a slot always meets the same converter group

  static constexpr uint32_t TOTAL = STG_OFF + NSTG * STG_SLOT + 128;              // + slack to align the base to 128 bytes
  static constexpr int ACC = MT * NTAPS * N;
  static constexpr uint32_t TCOLS = ACC <= 32 ? 32 : (ACC <= 64 ? 64 : (ACC <= 128 ? 128 : (ACC <= 256 ? 256 : 512)));
  static_assert(N % 16 == 0 && N <= 256 && ACC <= 512, "accumulators must fit the 512 TMEM columns");
  static_assert(N % XB == 0 && (X_BLK % 128) == 0 && (G_BLK % 128) == 0, "TMA boxes land on 128-byte boundaries");
  static_assert(NSTG >= 4, "at least four staging slots");
};

// two fp32 -> packed bf16 hi pair and lo (residual) pair; four of them = one 8-column chunk of one frame
struct Split8 { uint32_t h[4], l[4]; };
__device__ __forceinline__ Split8 split_chunk(const float4& a, const float4& b) {
  Split8 r;
  split2(a.x, a.y, r.h[0], r.l[0]); split2(a.z, a.w, r.h[1], r.l[1]);
  split2(b.x, b.y, r.h[2], r.l[2]); split2(b.z, b.w, r.h[3], r.l[3]);
  return r;
}
// One conversion task: the 32 bytes (8 columns) at `q` of a dense fp32 row -> 16-byte hi / lo chunks.  Lanes of a quarter-warp
// sit on 8 consecutive chunks of the same row; reading half `h` = (c >> 2) & 1 first spreads them over all eight 16-byte bank groups.
// Split in a load and a store half: a thread first pulls ALL its tasks of a tile into registers (the compiler will not move a
// shared-memory load above an earlier shared-memory store), hands the staging slot back, and only then converts.
struct Raw8 { float4 a, b; };
__device__ __forceinline__ Raw8 load_chunk(const uint8_t* q, int h) {
  Raw8 r;
  r.a = *reinterpret_cast<const float4*>(q + h * 16);
  r.b = *reinterpret_cast<const float4*>(q + (h ^ 1) * 16);
  return r;
}
__device__ __forceinline__ void store_chunk(const Raw8& raw, int h, uint8_t* hi, uint8_t* lo) {
  const Split8 r = split_chunk(raw.a, raw.b);                 // r.{h,l}[0..1] = the half read first
  *reinterpret_cast<uint4*>(hi) = h ? make_uint4(r.h[2], r.h[3], r.h[0], r.h[1]) : make_uint4(r.h[0], r.h[1], r.h[2], r.h[3]);
  *reinterpret_cast<uint4*>(lo) = h ? make_uint4(r.l[2], r.l[3], r.l[0], r.l[1]) : make_uint4(r.l[0], r.l[1], r.l[2], r.l[3]);
}

template <int MT, int NCH, int NTAPS>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  using C = Cfg<MT, NCH, NTAPS>;
  constexpr int N = C::N, NSTG = C::NSTG;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t stg_full[NSTG_MAX], stg_empty[NSTG_MAX], op_full[NOPS], op_empty[NOPS], acc_bar;
  __shared__ uint32_t tmem_base;
  uint8_t* const smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);      // TMA boxes want 128-byte aligned destinations
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.y * C::MC, n0 = blockIdx.z * N;
  const int mv = min(C::MC, p.m_valid - m0), nv = min(N, p.n_valid - n0);          // valid G / X columns of this block
  const int n_tiles = p.B * p.tiles_per_utt;
  uint8_t* const stg_base = smem + C::STG_OFF;

  if (tid == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(&stg_full[s], 1); mbar_init(&stg_empty[s], CONV_THREADS); }
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

  if (warp == LOAD_WARP) {
    // ================= loader: one elected thread, MT + NXB tensor copies per tile =================
    // (the whole warp walks the loop so that the copy instructions sit in warp-uniform control flow: their descriptor and
    // coordinate operands are uniform registers, and a divergent branch would make the compiler wrap each one in a vote loop)
    const bool two = p.m_split < p.m_valid;                  // a second G source exists
    if (lane == 0) {
      tma_prefetch_desc(&p.map_g0); tma_prefetch_desc(&p.map_x);
      if (two) tma_prefetch_desc(&p.map_g1);
    }
    const uint32_t stg0 = smem_u32(stg_base);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int slot = it % NSTG;
      if (it >= NSTG) mbar_wait(&stg_empty[slot], (uint32_t)(((it / NSTG) - 1) & 1));
      const int b = tile / p.tiles_per_utt, t0 = (tile - b * p.tiles_per_utt) * TF;
      const uint32_t dst = stg0 + (uint32_t)slot * C::STG_SLOT;
      if (elect_one()) {
        const bool tr = SLU_TRACE(p) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 64;
        if (tr) p.trace[it * 8 + 0] = clock64();
        mbar_arrive_expect_tx(&stg_full[slot], C::STG_SLOT);
#pragma unroll
        for (int k = 0; k < MT; ++k) {
          const int m = m0 + k * GB;                           // columns beyond the source's width read zeros
          if (two && m >= p.m_split) tma_load_3d(dst + (uint32_t)k * C::G_BLK, &p.map_g1, m - p.m_split, t0, b, &stg_full[slot]);
          else tma_load_3d(dst + (uint32_t)k * C::G_BLK, &p.map_g0, m, t0, b, &stg_full[slot]);
        }
#pragma unroll
        for (int j = 0; j < C::NXB; ++j)
          tma_load_3d(dst + C::STG_G + (uint32_t)j * C::X_BLK, &p.map_x, n0 + j * C::XB, t0 + p.shift0, b, &stg_full[slot]);
        if (tr) p.trace[it * 8 + 1] = clock64();
      }
      __syncwarp();
    }
  } else if (warp < CONV_WARPS) {
    // ================= converters: dense fp32 boxes -> bf16 hi/lo MN-major images =================
    const int grp = warp >> 2, ctid = tid & (CONV_THREADS - 1);
    for (int it = grp; it < my_tiles; it += CONV_GROUPS) {
      const int slot = it % NSTG, s = it % NOPS;
      mbar_wait(&stg_full[slot], (uint32_t)((it / NSTG) & 1));
      const bool tr = SLU_TRACE(p) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && ctid == 0 && it < 64;
      if (tr) p.trace[it * 8 + 2] = clock64();
      const uint8_t* src = stg_base + slot * C::STG_SLOT;
      if (SLU_DBG(p) & 2) {
        if (it >= NOPS) mbar_wait(&op_empty[s], (uint32_t)(((it / NOPS) - 1) & 1));
        mbar_arrive(&stg_empty[slot]); mbar_arrive(&op_full[s]);
        if (tr) p.trace[it * 8 + 3] = clock64();
        continue;
      }
      constexpr int GT = MT * 256 / CONV_THREADS;                                   // tasks per thread: G ...
      constexpr int XTOT = C::NXB * C::XR * C::XCHP, XT = (XTOT + CONV_THREADS - 1) / CONV_THREADS;   // ... and X
      Raw8 rg[GT], rx[XT];
#pragma unroll
      for (int u = 0; u < GT; ++u) {                                               // (block k, frame f, chunk c): 16 chunks per row
        const int i = ctid + u * CONV_THREADS, c = i & 15, f = (i >> 4) & 15, k = i >> 8;
        rg[u] = load_chunk(src + k * C::G_BLK + f * (GB * 4) + c * 32, (c >> 2) & 1);
      }
      const uint8_t* srcx = src + C::STG_G;
#pragma unroll
      for (int u = 0; u < XT; ++u) {
        const int i = ctid + u * CONV_THREADS, c = i % C::XCHP, r = i / C::XCHP, f = r % C::XR, j = r / C::XR;
        rx[u].a = rx[u].b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < XTOT && c < C::XCH) rx[u] = load_chunk(srcx + j * C::X_BLK + f * (C::XB * 4) + c * 32, (c >> 2) & 1);
      }
      // Hand the slot back once the tile is in registers.  "In registers" has to be enforced: an mbarrier arrive does not wait for
      // earlier shared-memory loads of the thread to return, and a TMA refill that hits in L2 can overtake them (measured: random
      // corrupted rows).  The arrive's address is therefore made data-dependent on every load (x & 0 with a run-time 0).
      uint32_t dep = 0;
#pragma unroll
      for (int u = 0; u < GT; ++u) dep ^= __float_as_uint(rg[u].a.x) ^ __float_as_uint(rg[u].b.x);
#pragma unroll
      for (int u = 0; u < XT; ++u) dep ^= __float_as_uint(rx[u].a.x) ^ __float_as_uint(rx[u].b.x);
      if (!(SLU_DBG(p) & 8)) mbar_arrive(reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(&stg_empty[slot]) + (dep & (uint32_t)p.zero)));
      if (it >= NOPS) mbar_wait(&op_empty[s], (uint32_t)(((it / NOPS) - 1) & 1));
      uint8_t* a_hi = smem + s * C::OP_STAGE;
      uint8_t* a_lo = a_hi + C::A_PART;
      uint8_t* b_hi = a_hi + 2 * C::A_PART;
      uint8_t* b_lo = b_hi + C::B_PART;
#pragma unroll
      for (int u = 0; u < GT; ++u) {
        const int i = ctid + u * CONV_THREADS, c = i & 15, f = (i >> 4) & 15, k = i >> 8;
        const uint32_t off = (uint32_t)(k * 16 + c) * C::SBO_A + (uint32_t)f * 16;
        store_chunk(rg[u], (c >> 2) & 1, a_hi + off, a_lo + off);
      }
#pragma unroll
      for (int u = 0; u < XT; ++u) {
        const int i = ctid + u * CONV_THREADS, c = i % C::XCHP, r = i / C::XCHP, f = r % C::XR, j = r / C::XR;
        if (i < XTOT && c < C::XCH) {
          const uint32_t off = (uint32_t)(j * C::XCH + c) * C::SBO_B + (uint32_t)f * 16;
          store_chunk(rx[u], (c >> 2) & 1, b_hi + off, b_lo + off);
        }
      }
      if (SLU_DBG(p) & 8) mbar_arrive(&stg_empty[slot]);
      fence_async_smem();                       // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&op_full[s]);
      if (tr) p.trace[it * 8 + 3] = clock64();
    }
  } else {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_mn(128, N);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % NOPS;
      mbar_wait(&op_full[s], (uint32_t)((it / NOPS) & 1));
      fence_after_sync();
      if (elect_one()) {
        const bool tr = SLU_TRACE(p) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 64;
        if (tr) p.trace[it * 8 + 4] = clock64();
        const uint32_t sa = smem_u32(smem + s * C::OP_STAGE);
        // MN-major, no swizzle: LBO = stride between 8-frame K groups (128 B), SBO = stride between 8-wide MN chunks
        const uint64_t bh0 = smem_desc(sa + 2 * C::A_PART, 128, C::SBO_B), bl0 = smem_desc(sa + 2 * C::A_PART + C::B_PART, 128, C::SBO_B);
        const uint32_t acc = it ? 1u : 0u;
#pragma unroll
        for (int mi = 0; mi < ((SLU_DBG(p) & 1) ? 0 : MT); ++mi) {
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
        if (tr) p.trace[it * 8 + 5] = clock64();
      }
      __syncwarp();
    }
  }

  // ================= flush: TMEM -> transposed through shared memory -> fp32 atomics (coalesced when s_n == 1) =================
  if (my_tiles > 0 && warp < 8 && !(SLU_DBG(p) & 4)) {
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

// (column, frame, utterance) view of base[(b*T + t)*ld + c], c < cols, with a [box_rows frames][box_cols columns] box; out-of-range
// coordinates (frames < 0 or >= T, columns >= cols) read zeros
int make_map(CUtensorMap* m, const float* base, long ld, int cols, int T, int B, int box_cols, int box_rows) {
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)T * (cuuint64_t)ld * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  return tmap::encode_f32(m, base, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

int x_box_cols(int N, int taps) {               // Cfg<>::XB of the instantiation run() picks
  if (taps == 5) return N <= 64 ? 64 : 80;
  return N <= 64 ? 64 : 128;
}

}  // namespace

// out[m*s_m + n*s_n + tap*s_tap] += sum_{b,t} G[(b*T+t)][m] * X[(b*T + t + shift0 + tap)*ldx + n]   (frames outside [0,T) read 0)
// with G rows [0, m_split) taken from G0 (pitch ldg0) and rows [m_split, M) from G1 (pitch ldg1); see include/slu_b200.h.
extern "C" int slu_wgrad2_tc(const float* G0, long ldg0, int m_split, const float* G1, long ldg1, int M, const float* X, long ldx, int N,
                             int B, int T, int taps, int shift0, float* out, long s_m, long s_n, long s_tap, void* stream) {
  if (M <= 0 || N <= 0 || B <= 0 || T <= 0 || (taps != 1 && taps != 5) || m_split < 0 || m_split > M) return (int)cudaErrorInvalidValue;
  if (taps > 1 && (shift0 < -4 || shift0 + taps - 1 > 4)) return (int)cudaErrorInvalidValue;
  if (m_split < M && (!G1 || (m_split % GB) != 0 || m_split == 0)) return (int)cudaErrorInvalidValue;
  if (!aligned16(G0) || !aligned16(G1) || !aligned16(X) || (ldg0 & 3) || (ldg1 & 3) || (ldx & 3) || (M & 3) || (N & 3))
    return (int)cudaErrorInvalidValue;
  if ((long)T * (ldg0 > ldx ? ldg0 : ldx) * 4 >= (1L << 40) || (long)B * T >= (1L << 31)) return SLU_ERR_TOO_LARGE;
  WgradParams p;
  const bool two = m_split < M;
  int err = make_map(&p.map_g0, G0, ldg0, two ? m_split : M, T, B, GB, TF);
  if (!err) err = two ? make_map(&p.map_g1, G1, ldg1, M - m_split, T, B, GB, TF) : make_map(&p.map_g1, G0, ldg0, M, T, B, GB, TF);
  if (!err) err = make_map(&p.map_x, X, ldx, N, T, B, x_box_cols(N, taps), TF + taps - 1);
  if (err) return err;
  p.m_split = two ? m_split : M;
  p.m_valid = M; p.n_valid = N; p.B = B; p.T = T; p.shift0 = shift0;
  p.tiles_per_utt = (T + TF - 1) / TF; p.out = out; p.s_m = s_m; p.s_n = s_n; p.s_tap = s_tap; p.dbg = g_dbg; p.trace = g_trace; p.zero = 0;
  return run(p, taps, (cudaStream_t)stream);
}

// Single-source form.
extern "C" int slu_wgrad_tc(const float* G, long ldg, int M, const float* X, long ldx, int N, int B, int T, int taps, int shift0,
                            float* out, long s_m, long s_n, long s_tap, void* stream) {
  return slu_wgrad2_tc(G, ldg, M, nullptr, 0, M, X, ldx, N, B, T, taps, shift0, out, s_m, s_n, s_tap, stream);
}

// Developer switch for tools/wgrad_only.py (which pipeline stage bounds the kernel): bit 0 skips the MMAs, bit 1 the conversion,
// bit 2 the flush.  Results are meaningless with any bit set.
extern "C" int slu_debug_wgrad_mode(int mode) {
  g_dbg = mode;
#ifdef SLU_KERNEL_DEBUG
  return 0;
#else
  return mode ? (int)cudaErrorNotSupported : 0;      // built without the switches
#endif
}
// CTA (0,0,0) records clock64() at its hand-off points into buf[64 tiles][8] (NULL: off).
extern "C" int slu_debug_wgrad_trace(long long* buf) { g_trace = buf; return 0; }
