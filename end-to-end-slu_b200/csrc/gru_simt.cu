// Persistent bidirectional-GRU recurrence, fp32 CUDA-core variant (exact-precision path; the
// tcgen05 variant lives in gru_tc.cu).  One CTA = (direction, 4 batch rows); W_hh of that
// direction stays resident in shared memory for all T steps; h stays in shared memory/registers.
// The x-projection  gx = x.W_ih^T + b_ih  is a dense GEMM done beforehand (both directions, N=768).
// Fused into the step epilogue: gate sigmoid/tanh, Dropout mask multiply and Downsample(avg,2).
// Restates nn.GRU as used at reference models.py:232/262/686 (+ :246-253 dropout/downsample).
//
// Layouts (all fp32, row-major):
//   gx     [B][T][768]   col = d*384 + g*128 + j      (g: 0=r 1=z 2=n)
//   y_full [B][T][256]   col = d*128 + j              raw h_t  (kept for backward)
//   y_out  [B][ceil(T/ds)][256]                       after dropout-mask and avg-downsample
//   stash  [B][T][1024]  col = d*512 + 4*j + s        s: 0=r 1=z 2=n 3=hn (=W_hn h + b_hn), training only
//   mask   [B][T][256]   dropout keep-mask pre-scaled by 1/(1-p), or NULL
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int BT = 4;                       // batch rows per CTA
constexpr size_t W_SMEM = 3 * 32 * 128 * sizeof(float4);   // 196608 B

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// lane q (0..3) ends up with sum over the 4 q-lanes of v[q]   (q = lane>>3; partners lane^16, lane^8)
__device__ __forceinline__ float reduce_scatter4(const float v[4], int q) {
  const int hi = q >> 1, lo = q & 1;
  float k0 = hi ? v[2] : v[0], k1 = hi ? v[3] : v[1];
  float s0 = hi ? v[0] : v[2], s1 = hi ? v[1] : v[3];
  float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
  float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
  float keep = lo ? r1 : r0, send = lo ? r0 : r1;
  return keep + __shfl_xor_sync(0xffffffffu, send, 8);
}

template <bool STASH>
__global__ void __launch_bounds__(512, 1)
gru_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
               const float* __restrict__ mask, uint32_t drop_thr, float drop_scale, uint64_t drop_seed_in,
               const unsigned long long* __restrict__ drop_seed_dev, int B, int T, int ds, float* __restrict__ y_full,
               float* __restrict__ y_out, float* __restrict__ stash) {
  const uint64_t drop_seed = drop_seed_dev ? (drop_seed_in ^ (uint64_t)__ldg(drop_seed_dev)) : drop_seed_in;
  extern __shared__ float4 smem4[];
  float4* Ws = smem4;                                        // [(g*32 + c)*128 + j] = W[g*128+j][4c..4c+3]
  float* hs = reinterpret_cast<float*>(smem4 + 3 * 32 * 128);  // [2][BT][128]
  const int d = blockIdx.y, b0 = blockIdx.x * BT, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, q = lane >> 3, j = warp * 8 + (lane & 7);
  const float* Wd = w_hh + (size_t)d * SLU_G3 * SLU_H;
  for (int idx = tid; idx < 3 * 32 * 128; idx += 512) {
    const int jj = idx & 127, c = (idx >> 7) & 31, g = idx >> 12;
    Ws[idx] = *reinterpret_cast<const float4*>(Wd + (size_t)(g * 128 + jj) * SLU_H + 4 * c);
  }
  for (int idx = tid; idx < 2 * BT * SLU_H; idx += 512) hs[idx] = 0.f;
  const float bhr = b_hh[d * SLU_G3 + j], bhz = b_hh[d * SLU_G3 + 128 + j], bhn = b_hh[d * SLU_G3 + 256 + j];
  const int b = b0 + q;
  const bool valid = b < B;
  const int T2 = (T + ds - 1) / ds;
  float hprev = 0.f, pend = 0.f;
  // software prefetch of step inputs
  float gxr = 0.f, gxz = 0.f, gxn = 0.f, mk = 1.f;
  auto load_step = [&](int t, float& r_, float& z_, float& n_, float& m_) {
    if (valid) {
      const float* p = gx + ((size_t)b * T + t) * 768 + d * SLU_G3 + j;
      r_ = __ldg(p); z_ = __ldg(p + 128); n_ = __ldg(p + 256);
      if (mask) m_ = __ldg(mask + ((size_t)b * T + t) * 256 + d * SLU_H + j);
      else if (drop_thr != 0u) {          // the canonical Philox mask (philox.cuh); this variant simply draws per element
        uint32_t w[4];
        slu_gru_mask_draws(b, d * SLU_H + j, t >> 3, drop_seed, w);
        m_ = slu_gru_mask_draw16(w, t) < drop_thr ? drop_scale : 0.f;
      } else m_ = 1.f;
    }
  };
  if (T > 0) load_step(d ? T - 1 : 0, gxr, gxz, gxn, mk);
  __syncthreads();
  int cur = 0;
  for (int s = 0; s < T; ++s) {
    const int t = d ? T - 1 - s : s;
    float nr = 0.f, nz = 0.f, nn = 0.f, nm = 1.f;
    if (s + 1 < T) load_step(d ? t - 1 : t + 1, nr, nz, nn, nm);
    const float4* h4 = reinterpret_cast<const float4*>(hs + cur * BT * SLU_H);
    float acc[3][BT];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int bb = 0; bb < BT; ++bb) acc[g][bb] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 hv[BT];
#pragma unroll
      for (int bb = 0; bb < BT; ++bb) hv[bb] = h4[bb * 32 + 8 * q + i];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const float4 w = Ws[(g * 32 + 8 * q + i) * 128 + j];
#pragma unroll
        for (int bb = 0; bb < BT; ++bb)
          acc[g][bb] = fmaf(w.x, hv[bb].x, fmaf(w.y, hv[bb].y, fmaf(w.z, hv[bb].z, fmaf(w.w, hv[bb].w, acc[g][bb]))));
      }
    }
    const float ghr = reduce_scatter4(acc[0], q), ghz = reduce_scatter4(acc[1], q), ghn = reduce_scatter4(acc[2], q);
    const float r = sigmoidf_(gxr + (ghr + bhr));
    const float z = sigmoidf_(gxz + (ghz + bhz));
    const float hn = ghn + bhn;
    const float n = tanhf(gxn + r * hn);
    const float hnew = (1.f - z) * n + z * hprev;
    hs[(cur ^ 1) * BT * SLU_H + q * SLU_H + j] = hnew;
    hprev = hnew;
    if (valid) {
      const size_t bt = (size_t)b * T + t;
      y_full[bt * 256 + d * SLU_H + j] = hnew;
      if (STASH) {
        *reinterpret_cast<float4*>(stash + bt * 1024 + d * 512 + 4 * j) = make_float4(r, z, n, hn);
      }
      const float val = hnew * mk;
      if (ds == 1) {
        y_out[bt * 256 + d * SLU_H + j] = val;
      } else {
        float* yo = y_out + ((size_t)b * T2 + (t >> 1)) * 256 + d * SLU_H + j;
        if ((t & 1) == 0 && t == T - 1) *yo = val;                   // odd tail frame: divisor 1 (ceil_mode)
        else if ((t & 1) == (d ? 1 : 0)) pend = val;                 // first visited of the pair
        else *yo = 0.5f * (pend + val);
      }
    }
    gxr = nr; gxz = nz; gxn = nn; mk = nm;
    __syncthreads();
    cur ^= 1;
  }
}

// Backward-through-time.  Consumes the stashed gates; emits the pre-activation gradients
//   dgx [B][T][768]  (dr, dz, dn wrt the x-projection -> dW_ih, db_ih, dX by GEMM)
//   dhn [B][T][256]  (dn*r: the n-gate gradient wrt W_hn h + b_hn -> dW_hh / db_hh n-rows)
// and carries dh through  dh_{t-1} += W_hh^T [dr, dz, dhn].
__global__ void __launch_bounds__(512, 1)
gru_bwd_kernel(const float* __restrict__ dy_out, const float* __restrict__ mask, uint32_t drop_thr, float drop_scale,
               uint64_t drop_seed_in, const unsigned long long* __restrict__ drop_seed_dev, const float* __restrict__ y_full,
               const float* __restrict__ stash, const float* __restrict__ w_hh, int B, int T, int ds,
               float* __restrict__ dgx, float* __restrict__ dhn_out, float* __restrict__ db_ih, float* __restrict__ db_hh) {
  const uint64_t drop_seed = drop_seed_dev ? (drop_seed_in ^ (uint64_t)__ldg(drop_seed_dev)) : drop_seed_in;
  extern __shared__ float4 smem4[];
  float4* Wt = smem4;                                          // [c*128 + k] = W[4c..4c+3][k], c < 96
  float* gs = reinterpret_cast<float*>(smem4 + 96 * 128);      // [2][BT][384]
  const int d = blockIdx.y, b0 = blockIdx.x * BT, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, q = lane >> 3, j = warp * 8 + (lane & 7);
  const float* Wd = w_hh + (size_t)d * SLU_G3 * SLU_H;
  for (int idx = tid; idx < 96 * 128; idx += 512) {
    const int k = idx & 127, c = idx >> 7;
    Wt[idx] = make_float4(Wd[(size_t)(4 * c + 0) * SLU_H + k], Wd[(size_t)(4 * c + 1) * SLU_H + k],
                          Wd[(size_t)(4 * c + 2) * SLU_H + k], Wd[(size_t)(4 * c + 3) * SLU_H + k]);
  }
  const int b = b0 + q;
  const bool valid = b < B;
  const int T2 = (T + ds - 1) / ds;
  struct In { float r, z, n, hn, hp, dy; };
  auto load_step = [&](int t, In& v) {
    v.r = v.z = v.n = v.hn = v.hp = v.dy = 0.f;
    if (!valid) return;
    const size_t bt = (size_t)b * T + t;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(stash + bt * 1024 + d * 512 + 4 * j));
    v.r = g4.x; v.z = g4.y; v.n = g4.z; v.hn = g4.w;
    const int tp = d ? t + 1 : t - 1;
    if (tp >= 0 && tp < T) v.hp = __ldg(y_full + ((size_t)b * T + tp) * 256 + d * SLU_H + j);
    float g;
    if (ds == 1) g = __ldg(dy_out + bt * 256 + d * SLU_H + j);
    else {
      g = __ldg(dy_out + ((size_t)b * T2 + (t >> 1)) * 256 + d * SLU_H + j);
      if (!((t & 1) == 0 && t == T - 1)) g *= 0.5f;
    }
    if (mask) g *= __ldg(mask + bt * 256 + d * SLU_H + j);
    else if (drop_thr != 0u) {
      uint32_t w[4];
      slu_gru_mask_draws((int)(bt / T), d * SLU_H + j, (int)(bt % T) >> 3, drop_seed, w);
      g *= slu_gru_mask_draw16(w, (int)(bt % T)) < drop_thr ? drop_scale : 0.f;
    }
    v.dy = g;
  };
  In cur_in, nxt_in;
  if (T > 0) load_step(d ? 0 : T - 1, cur_in);
  __syncthreads();
  float dh_rec = 0.f;
  float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_hn = 0.f;
  int cur = 0;
  for (int s = 0; s < T; ++s) {
    const int t = d ? s : T - 1 - s;
    if (s + 1 < T) load_step(d ? t + 1 : t - 1, nxt_in);
    const float dh = dh_rec + cur_in.dy;
    const float r = cur_in.r, z = cur_in.z, n = cur_in.n;
    const float dn_pre = dh * (1.f - z) * (1.f - n * n);
    const float dz_pre = dh * (cur_in.hp - n) * z * (1.f - z);
    const float dhn = dn_pre * r;
    const float dr_pre = dn_pre * cur_in.hn * r * (1.f - r);
    float* g = gs + cur * BT * SLU_G3 + q * SLU_G3;
    g[j] = dr_pre; g[128 + j] = dz_pre; g[256 + j] = dhn;
    if (valid) {
      const size_t bt = (size_t)b * T + t;
      float* p = dgx + bt * 768 + d * SLU_G3 + j;
      p[0] = dr_pre; p[128] = dz_pre; p[256] = dn_pre;
      dhn_out[bt * 256 + d * SLU_H + j] = dhn;
      sb_r += dr_pre; sb_z += dz_pre; sb_n += dn_pre; sb_hn += dhn;
    }
    __syncthreads();
    const float4* g4 = reinterpret_cast<const float4*>(gs + cur * BT * SLU_G3);
    float acc[BT];
#pragma unroll
    for (int bb = 0; bb < BT; ++bb) acc[bb] = 0.f;
#pragma unroll 6
    for (int i = 0; i < 24; ++i) {
      const float4 w = Wt[(24 * q + i) * 128 + j];
#pragma unroll
      for (int bb = 0; bb < BT; ++bb) {
        const float4 gv = g4[bb * 96 + 24 * q + i];
        acc[bb] = fmaf(w.x, gv.x, fmaf(w.y, gv.y, fmaf(w.z, gv.z, fmaf(w.w, gv.w, acc[bb]))));
      }
    }
    dh_rec = reduce_scatter4(acc, q) + dh * z;
    cur_in = nxt_in;
    cur ^= 1;
  }
  if (db_ih) {                   // parameter layout [2][384]: b_ih <- (dr, dz, dn), b_hh <- (dr, dz, dhn)
    float* pa = db_ih + d * SLU_G3 + j;
    float* pb = db_hh + d * SLU_G3 + j;
    atomicAdd(pa, sb_r); atomicAdd(pa + 128, sb_z); atomicAdd(pa + 256, sb_n);
    atomicAdd(pb, sb_r); atomicAdd(pb + 128, sb_z); atomicAdd(pb + 256, sb_hn);
  }
}

}  // namespace

extern "C" int slu_gru_fwd_simt(const float* gx, const float* w_hh, const float* b_hh, const float* drop_mask, float drop_p,
                                unsigned long long drop_seed, const unsigned long long* drop_seed_dev, int B, int T, int ds,
                                float* y_full, float* y_out, float* stash, void* stream) {
  if (B <= 0 || T <= 0 || (ds != 1 && ds != 2) || !(drop_p >= 0.f && drop_p < 1.f)) return (int)cudaErrorInvalidValue;
  const uint32_t thr = (!drop_mask && drop_p > 0.f) ? slu_keep_threshold16(drop_p) : 0u;
  const float dscale = (float)(1.0 / (1.0 - (double)drop_p));
  const size_t smem = W_SMEM + 2 * BT * SLU_H * sizeof(float);
  SLU_SMEM_ONCE(gru_fwd_kernel<true>, smem);
  SLU_SMEM_ONCE(gru_fwd_kernel<false>, smem);
  dim3 grid((B + BT - 1) / BT, 2);
  if (stash) gru_fwd_kernel<true><<<grid, 512, smem, (cudaStream_t)stream>>>(gx, w_hh, b_hh, drop_mask, thr, dscale, drop_seed, drop_seed_dev, B, T, ds, y_full, y_out, stash);
  else gru_fwd_kernel<false><<<grid, 512, smem, (cudaStream_t)stream>>>(gx, w_hh, b_hh, drop_mask, thr, dscale, drop_seed, drop_seed_dev, B, T, ds, y_full, y_out, nullptr);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_gru_bwd_simt(const float* dy_out, const float* drop_mask, float drop_p, unsigned long long drop_seed,
                                const unsigned long long* drop_seed_dev, const float* y_full, const float* stash, const float* w_hh,
                                int B, int T, int ds, float* dgx, float* dhn, float* db_ih, float* db_hh, void* stream) {
  if (B <= 0 || T <= 0 || (ds != 1 && ds != 2) || (db_ih == nullptr) != (db_hh == nullptr) || !(drop_p >= 0.f && drop_p < 1.f))
    return (int)cudaErrorInvalidValue;
  const uint32_t thr = (!drop_mask && drop_p > 0.f) ? slu_keep_threshold16(drop_p) : 0u;
  const float dscale = (float)(1.0 / (1.0 - (double)drop_p));
  const size_t smem = W_SMEM + 2 * BT * SLU_G3 * sizeof(float);
  SLU_SMEM_ONCE(gru_bwd_kernel, smem);
  dim3 grid((B + BT - 1) / BT, 2);
  gru_bwd_kernel<<<grid, 512, smem, (cudaStream_t)stream>>>(dy_out, drop_mask, thr, dscale, drop_seed, drop_seed_dev, y_full, stash, w_hh, B, T, ds, dgx, dhn, db_ih, db_hh);
  SLU_CHECK_LAUNCH();
  return 0;
}
