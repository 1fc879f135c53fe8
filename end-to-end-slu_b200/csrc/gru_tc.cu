// Persistent bidirectional-GRU recurrence on tcgen05 tensor cores (sm_100a).
//
// One CTA = (direction d, NB batch rows).  Per time step it needs  Gh[384 x NB] = W_hh[384 x 128] . h[128 x NB]:
//   * W_hh is WEIGHTS-STATIONARY IN TENSOR MEMORY: the three gate blocks (r, z, n; 128 rows each) are split into
//     bf16 hi + bf16 lo and stored once as tcgen05 A-operands (lane = hidden unit j, 64 columns per gate per part,
//     384 of the 512 TMEM columns).  Nothing is re-read from shared memory or HBM during the T steps.
//   * h_{t-1} is the B operand: a tiny K-major bf16 hi/lo tile [NB x 128] in shared memory, rewritten by the
//     epilogue threads every step (h itself stays fp32 in registers; only the MMA operand copy is rounded).
//   * every product is hi*hi + hi*lo + lo*hi (fp32 accumulate in TMEM): 72 MMAs (M=128, N=NB, K=16) per step.
//   * epilogue: tcgen05.ld the three [128 x NB] accumulators -> gate sigmoid/tanh -> h_t, fused with the
//     Dropout-mask multiply and Downsample(avg, 2) of reference models.py:246-253, and the training stash.
// The x-projection gx = x.W_ih^T + b_ih is a dense GEMM done beforehand for both directions (see gemm_tc.cu).
// Same layouts / semantics as gru_simt.cu (slu_gru_fwd_simt); restates nn.GRU at models.py:232/262/686.
#include "common.cuh"
#include "tc05.cuh"

namespace {
using namespace tc05;

constexpr int TC_THREADS = 256;       // 8 warps: warp w owns TMEM lanes 32*(w%4).., batch columns (w/4)*NB/2 ..
constexpr uint32_t ACC_COL = 384;     // accumulators start after the 384 weight columns

__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float fast_tanh(float v) {
  // tanh(v) = 2*sigmoid(2v) - 1, ex2-based (rel err ~2^-22), saturates cleanly for |v| large
  return __fdividef(2.0f, 1.0f + __expf(-2.0f * v)) - 1.0f;
}

// Load W rows (this thread's lane) into TMEM as split bf16 A-operands.  src: 3 blocks of [128][128] fp32 with
// element (row j, k) at src[g*block_stride + j*row_stride + k*k_stride].
__device__ __forceinline__ void load_weights_to_tmem(uint32_t tmem, uint32_t lane_base, const float* src, size_t block_stride,
                                                     size_t row_stride, size_t k_stride, int j, int half) {
  // `half` (0/1) splits the work between the two warps that share a lane quarter: half 0 -> k in [0,64), 1 -> [64,128)
  for (int g = 0; g < 3; ++g) {
    float row[64];
    const float* p = src + g * block_stride + (size_t)j * row_stride + (size_t)(half * 64) * k_stride;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) row[k] = __ldg(p + (size_t)k * k_stride);
    const uint32_t t_hi = tmem + lane_base + (uint32_t)(g * 64 + half * 32);
    tmem_store_row_split(t_hi, t_hi + 192, row, 64);
  }
}

template <int NB, bool STASH>
__global__ void __launch_bounds__(TC_THREADS, 1)
gru_fwd_tc_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                  const float* __restrict__ mask, int B, int T, int ds, float* __restrict__ y_full,
                  float* __restrict__ y_out, float* __restrict__ stash) {
  constexpr int NC = NB / 2;                         // batch columns per thread
  constexpr uint32_t LBO = NB * 16 + 16;             // padded: conflict-free 2-byte operand stores
  __shared__ __align__(128) uint8_t h_tile[2 * 16 * LBO];   // [hi | lo] x 16 k-chunks x (NB rows x 16 B + pad)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int d = blockIdx.y, b0 = blockIdx.x * NB;
  const int j = (warp & 3) * 32 + lane;              // hidden unit == TMEM lane
  const int c0 = (warp >> 2) * NC;                   // first batch column of this thread
  uint8_t* h_hi = h_tile;
  uint8_t* h_lo = h_tile + 16 * LBO;

  if (tid == 0) { mbar_init(&bar, 3); fence_mbar_init(); }   // 3 gate-issuer warps commit per step
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  load_weights_to_tmem(tmem, lane_base, w_hh + (size_t)d * SLU_G3 * SLU_H, (size_t)128 * SLU_H, SLU_H, 1, j, warp >> 2);

  const float bhr = b_hh[d * SLU_G3 + j], bhz = b_hh[d * SLU_G3 + 128 + j], bhn = b_hh[d * SLU_G3 + 256 + j];
  const int T2 = (T + ds - 1) / ds;
  float hprev[NC], pend[NC];
  float gxr[NC], gxz[NC], gxn[NC], mk[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) { hprev[c] = 0.f; pend[c] = 0.f; gxr[c] = gxz[c] = gxn[c] = 0.f; mk[c] = 1.f; }

  auto load_step = [&](int t, float* r_, float* z_, float* n_, float* m_) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int b = b0 + c0 + c;
      if (b < B) {
        const float* p = gx + ((size_t)b * T + t) * 768 + d * SLU_G3 + j;
        r_[c] = __ldg(p); z_[c] = __ldg(p + 128); n_[c] = __ldg(p + 256);
        m_[c] = mask ? __ldg(mask + ((size_t)b * T + t) * 256 + d * SLU_H + j) : 1.f;
      }
    }
  };
  load_step(d ? T - 1 : 0, gxr, gxz, gxn, mk);
  const uint32_t idesc = idesc_bf16_f32(128, NB);
  const uint32_t acc_addr = tmem + lane_base + ACC_COL + c0;
  const uint64_t bdesc_hi = smem_desc(smem_u32(h_hi), LBO, 128), bdesc_lo = smem_desc(smem_u32(h_lo), LBO, 128);
  // byte offset of element (k = j) inside a k-chunk-major row b: (j/8)*LBO + b*16 + (j%8)*2
  const uint32_t h_off = (uint32_t)(j >> 3) * LBO + (uint32_t)(j & 7) * 2;

  for (int s = 0; s < T; ++s) {
    const int t = d ? T - 1 - s : s;
    float nr[NC], nz[NC], nn[NC], nm[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { nr[c] = nz[c] = nn[c] = 0.f; nm[c] = 1.f; }
    if (s + 1 < T) load_step(d ? t - 1 : t + 1, nr, nz, nn, nm);
    float ar[NC], az[NC], an[NC];
    if (s == 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ar[c] = az[c] = an[c] = 0.f;       // h_{-1} = 0
    } else {
      mbar_wait(&bar, (uint32_t)((s - 1) & 1));
      fence_after_sync();
      if (NC == 8) { tmem_ld8(acc_addr, ar); tmem_ld8(acc_addr + NB, az); tmem_ld8(acc_addr + 2 * NB, an); }
      else { tmem_ld16(acc_addr, ar); tmem_ld16(acc_addr + NB, az); tmem_ld16(acc_addr + 2 * NB, an); }
      tmem_ld_wait();
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int b = b0 + c0 + c;
      const float r = fast_sigmoid(gxr[c] + (ar[c] + bhr));
      const float z = fast_sigmoid(gxz[c] + (az[c] + bhz));
      const float hn = an[c] + bhn;
      const float n = fast_tanh(gxn[c] + r * hn);
      const float hnew = (1.f - z) * n + z * hprev[c];
      hprev[c] = hnew;
      // MMA operand copy of h_t: bf16 hi + lo
      const __nv_bfloat16 hh = __float2bfloat16_rn(hnew);
      const __nv_bfloat16 hl = __float2bfloat16_rn(hnew - __bfloat162float(hh));
      *reinterpret_cast<__nv_bfloat16*>(h_hi + h_off + (c0 + c) * 16) = hh;
      *reinterpret_cast<__nv_bfloat16*>(h_lo + h_off + (c0 + c) * 16) = hl;
      if (b < B) {
        const size_t bt = (size_t)b * T + t;
        y_full[bt * 256 + d * SLU_H + j] = hnew;
        if (STASH) {
          float* sp = stash + bt * 1024 + d * 512 + j;
          sp[0] = r; sp[128] = z; sp[256] = n; sp[384] = hn;
        }
        const float val = hnew * mk[c];
        if (ds == 1) {
          y_out[bt * 256 + d * SLU_H + j] = val;
        } else {
          float* yo = y_out + ((size_t)b * T2 + (t >> 1)) * 256 + d * SLU_H + j;
          if ((t & 1) == 0 && t == T - 1) *yo = val;
          else if ((t & 1) == (d ? 1 : 0)) pend[c] = val;
          else *yo = 0.5f * (pend[c] + val);
        }
      }
      gxr[c] = nr[c]; gxz[c] = nz[c]; gxn[c] = nn[c]; mk[c] = nm[c];
    }
    if (s + 1 < T) {
      fence_async_smem();          // h tile (generic-proxy stores) -> visible to the tensor core (async proxy)
      fence_before_sync();         // order this thread's tcgen05.ld before the barrier
      __syncthreads();
      if (warp < 3) {                 // warps 0,1,2 (three different SM sub-partitions) issue gate r, z, n concurrently
        if (elect_one()) {
          fence_after_sync();
          const uint32_t dacc = tmem + ACC_COL + warp * NB, a_hi = tmem + warp * 64, a_lo = a_hi + 192;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t bh = desc_advance(bdesc_hi, kk * 2 * LBO), bl = desc_advance(bdesc_lo, kk * 2 * LBO);
            mma_bf16_ts(dacc, a_hi + kk * 8, bh, idesc, kk > 0 ? 1u : 0u);
            mma_bf16_ts(dacc, a_hi + kk * 8, bl, idesc, 1u);
            mma_bf16_ts(dacc, a_lo + kk * 8, bh, idesc, 1u);
          }
          mma_commit(&bar);
        }
        __syncwarp();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Backward through time on tensor cores.  dh_{t-1}[k] += sum_row W_hh[row][k] * dG[row]:  M = 128 (k), K = 384 (gate rows),
// N = NB.  W_hh^T (hi/lo) is stationary in TMEM (2 x 192 columns); dG = (dr, dz, dhn) is the shared-memory B tile.
template <int NB>
__global__ void __launch_bounds__(TC_THREADS, 1)
gru_bwd_tc_kernel(const float* __restrict__ dy_out, const float* __restrict__ mask, const float* __restrict__ y_full,
                  const float* __restrict__ stash, const float* __restrict__ w_hh, int B, int T, int ds,
                  float* __restrict__ dgx, float* __restrict__ dhn_out) {
  constexpr int NC = NB / 2;
  constexpr uint32_t LBO = NB * 16 + 16;
  __shared__ __align__(128) uint8_t g_tile[2 * 48 * LBO];   // [hi | lo] x 48 k-chunks (384 gate rows)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int d = blockIdx.y, b0 = blockIdx.x * NB;
  const int j = (warp & 3) * 32 + lane;
  const int c0 = (warp >> 2) * NC;
  uint8_t* g_hi = g_tile;
  uint8_t* g_lo = g_tile + 48 * LBO;
  if (tid == 0) { mbar_init(&bar, 3); fence_mbar_init(); }   // the K=384 reduction is issued as 3 chunks by 3 warps
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  // A[k=j][K index = row]: element (lane j, kk = g*128 + i) = W_hh[d][g*128 + i][j]  -> block stride 128*128, "row" stride 1, k stride 128
  load_weights_to_tmem(tmem, lane_base, w_hh + (size_t)d * SLU_G3 * SLU_H, (size_t)128 * SLU_H, 1, SLU_H, j, warp >> 2);

  const int T2 = (T + ds - 1) / ds;
  struct In { float r, z, n, hn, hp, dy; };
  In cur[NC], nxt[NC];
  auto load_step = [&](int t, In* v) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      v[c].r = v[c].z = v[c].n = v[c].hn = v[c].hp = v[c].dy = 0.f;
      const int b = b0 + c0 + c;
      if (b >= B) continue;
      const size_t bt = (size_t)b * T + t;
      const float* sp = stash + bt * 1024 + d * 512 + j;
      v[c].r = __ldg(sp); v[c].z = __ldg(sp + 128); v[c].n = __ldg(sp + 256); v[c].hn = __ldg(sp + 384);
      const int tp = d ? t + 1 : t - 1;
      if (tp >= 0 && tp < T) v[c].hp = __ldg(y_full + ((size_t)b * T + tp) * 256 + d * SLU_H + j);
      float g;
      if (ds == 1) g = __ldg(dy_out + bt * 256 + d * SLU_H + j);
      else {
        g = __ldg(dy_out + ((size_t)b * T2 + (t >> 1)) * 256 + d * SLU_H + j);
        if (!((t & 1) == 0 && t == T - 1)) g *= 0.5f;
      }
      if (mask) g *= __ldg(mask + bt * 256 + d * SLU_H + j);
      v[c].dy = g;
    }
  };
  load_step(d ? 0 : T - 1, cur);
  const uint32_t idesc = idesc_bf16_f32(128, NB);
  const uint32_t acc_addr = tmem + lane_base + ACC_COL + c0;
  const uint64_t bdesc_hi = smem_desc(smem_u32(g_hi), LBO, 128), bdesc_lo = smem_desc(smem_u32(g_lo), LBO, 128);
  const uint32_t g_off = (uint32_t)(j >> 3) * LBO + (uint32_t)(j & 7) * 2;     // + gate*16*LBO + b*16
  float dh_direct[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) dh_direct[c] = 0.f;

  for (int s = 0; s < T; ++s) {
    const int t = d ? s : T - 1 - s;
    if (s + 1 < T) load_step(d ? t + 1 : t - 1, nxt);
    float rec[NC];
    if (s == 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) rec[c] = 0.f;
    } else {
      mbar_wait(&bar, (uint32_t)((s - 1) & 1));
      fence_after_sync();
      float r1[NC], r2[NC];                 // three partial accumulators (one per gate-row chunk / issuing warp)
      if (NC == 8) { tmem_ld8(acc_addr, rec); tmem_ld8(acc_addr + NB, r1); tmem_ld8(acc_addr + 2 * NB, r2); }
      else { tmem_ld16(acc_addr, rec); tmem_ld16(acc_addr + NB, r1); tmem_ld16(acc_addr + 2 * NB, r2); }
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NC; ++c) rec[c] += r1[c] + r2[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int b = b0 + c0 + c;
      const float dh = rec[c] + dh_direct[c] + cur[c].dy;
      const float r = cur[c].r, z = cur[c].z, n = cur[c].n;
      const float dn_pre = dh * (1.f - z) * (1.f - n * n);
      const float dz_pre = dh * (cur[c].hp - n) * z * (1.f - z);
      const float dhn = dn_pre * r;
      const float dr_pre = dn_pre * cur[c].hn * r * (1.f - r);
      dh_direct[c] = dh * z;
      const float gv[3] = {dr_pre, dz_pre, dhn};
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const __nv_bfloat16 hh = __float2bfloat16_rn(gv[g]);
        const __nv_bfloat16 hl = __float2bfloat16_rn(gv[g] - __bfloat162float(hh));
        const uint32_t off = g_off + (uint32_t)g * 16 * LBO + (c0 + c) * 16;
        *reinterpret_cast<__nv_bfloat16*>(g_hi + off) = hh;
        *reinterpret_cast<__nv_bfloat16*>(g_lo + off) = hl;
      }
      if (b < B) {
        const size_t bt = (size_t)b * T + t;
        float* p = dgx + bt * 768 + d * SLU_G3 + j;
        p[0] = dr_pre; p[128] = dz_pre; p[256] = dn_pre;
        dhn_out[bt * 256 + d * SLU_H + j] = dhn;
      }
      cur[c] = nxt[c];
    }
    if (s + 1 < T) {
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (warp < 3) {                 // warp g reduces gate-row chunk g (K steps 8g..8g+7) into its own accumulator
        if (elect_one()) {
          fence_after_sync();
          const uint32_t dacc = tmem + ACC_COL + warp * NB, a_hi = tmem + warp * 64, a_lo = a_hi + 192;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t koff = (uint32_t)(warp * 8 + kk) * 2 * LBO;
            const uint64_t bh = desc_advance(bdesc_hi, koff), bl = desc_advance(bdesc_lo, koff);
            mma_bf16_ts(dacc, a_hi + kk * 8, bh, idesc, kk > 0 ? 1u : 0u);
            mma_bf16_ts(dacc, a_hi + kk * 8, bl, idesc, 1u);
            mma_bf16_ts(dacc, a_lo + kk * 8, bh, idesc, 1u);
          }
          mma_commit(&bar);
        }
        __syncwarp();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int slu_gru_fwd_tc(const float* gx, const float* w_hh, const float* b_hh, const float* drop_mask, int B, int T,
                              int ds, float* y_full, float* y_out, float* stash, void* stream) {
  if (B <= 0 || T <= 0 || (ds != 1 && ds != 2)) return (int)cudaErrorInvalidValue;
  constexpr int NB = 16;
  dim3 grid((B + NB - 1) / NB, 2);
  if (stash) gru_fwd_tc_kernel<NB, true><<<grid, TC_THREADS, 0, (cudaStream_t)stream>>>(gx, w_hh, b_hh, drop_mask, B, T, ds, y_full, y_out, stash);
  else gru_fwd_tc_kernel<NB, false><<<grid, TC_THREADS, 0, (cudaStream_t)stream>>>(gx, w_hh, b_hh, drop_mask, B, T, ds, y_full, y_out, nullptr);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_gru_bwd_tc(const float* dy_out, const float* drop_mask, const float* y_full, const float* stash,
                              const float* w_hh, int B, int T, int ds, float* dgx, float* dhn, void* stream) {
  if (B <= 0 || T <= 0 || (ds != 1 && ds != 2)) return (int)cudaErrorInvalidValue;
  constexpr int NB = 16;
  dim3 grid((B + NB - 1) / NB, 2);
  gru_bwd_tc_kernel<NB><<<grid, TC_THREADS, 0, (cudaStream_t)stream>>>(dy_out, drop_mask, y_full, stash, w_hh, B, T, ds, dgx, dhn);
  SLU_CHECK_LAUNCH();
  return 0;
}
