// Intent head of the SLU model as two kernels (reference models.py:709 Linear, :112-123 FinalPool = max over time,
// :811-823 per-slot cross-entropy summed over slots + all-slots-right accuracy):
//   logits[b][c] = max_t ( feats[b][t][:] . W[c][:] ) + bias[c]                     c < C = sum(values_per_slot)
//   loss = sum_slots mean_b CE(logits[b][slot], y[b][slot]);   acc = mean_b [ argmax of every slot == y ]
// The arithmetic is tiny (B*T*C*256 MACs, C = 24 for FSC): the point is ONE launch per direction instead of ~35 library
// launches.  fp32 CUDA-core math; the batch means are reduced in a fixed order (last-CTA pattern), so results are
// bit-reproducible run to run.
#include "common.cuh"

namespace {

constexpr int HEAD_F = 256;        // feature width (2 * SLU_H)
constexpr int HEAD_MAXC = 128;     // max total values
constexpr int HEAD_MAXS = 16;      // max slots
constexpr int HEAD_TT = 32;        // frames per staged tile

struct SlotTable {
  int n_slots;
  int start[HEAD_MAXS + 1];
};

// ---- forward: one CTA (256 threads) per utterance ------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ W,
                                                       const float* __restrict__ bias, const long long* __restrict__ y, int B, int T,
                                                       int C, SlotTable slots, float* __restrict__ logits, int* __restrict__ tstar,
                                                       float* __restrict__ row_loss, float* __restrict__ row_ok,
                                                       float* __restrict__ loss_acc, unsigned int* __restrict__ ticket) {
  extern __shared__ float sm[];
  float* w_s = sm;                           // [C][257]   padded rows: a warp walking c reads distinct banks
  float* f_s = w_s + ((C * 257 + 3) & ~3);   // [TT][256]  one time tile (row t is broadcast within a warp), 16 B aligned
  float* sc = f_s + HEAD_TT * HEAD_F;        // [TT][C]    scores of the tile
  float* lg = sc + HEAD_TT * C;              // [C]        final logits of this utterance
  __shared__ float slot_loss[HEAD_MAXS];
  __shared__ int slot_ok[HEAD_MAXS];
  __shared__ float red[2][256];
  __shared__ bool is_last;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < C * HEAD_F; i += 256) w_s[(i >> 8) * 257 + (i & 255)] = W[i];
  float best = -INFINITY;                    // thread c < C owns class c
  int best_t = 0;
  for (int t0 = 0; t0 < T; t0 += HEAD_TT) {
    const int tt = min(HEAD_TT, T - t0);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(feats + ((long)b * T + t0) * HEAD_F);
    for (int i = tid; i < tt * (HEAD_F / 4); i += 256) reinterpret_cast<float4*>(f_s)[i] = src[i];
    __syncthreads();
    for (int i = tid; i < tt * C; i += 256) {
      const int t = i / C, c = i - t * C;
      const float* fr = f_s + t * HEAD_F;
      const float* wr = w_s + c * 257;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int k = 0; k < HEAD_F; k += 4) {
        a0 = fmaf(fr[k], wr[k], a0);
        a1 = fmaf(fr[k + 1], wr[k + 1], a1);
        a2 = fmaf(fr[k + 2], wr[k + 2], a2);
        a3 = fmaf(fr[k + 3], wr[k + 3], a3);
      }
      sc[i] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < C)
      for (int t = 0; t < tt; ++t) {
        const float v = sc[t * C + tid];
        if (v > best) { best = v; best_t = t0 + t; }          // strict: the first maximum wins
      }
  }
  if (tid < C) {
    const float v = best + bias[tid];
    lg[tid] = v;
    logits[(long)b * C + tid] = v;
    tstar[(long)b * C + tid] = best_t;
  }
  if (y == nullptr) return;
  __syncthreads();
  if (tid < slots.n_slots) {
    const int s0 = slots.start[tid], s1 = slots.start[tid + 1];
    float m = lg[s0];
    int am = s0;
    for (int c = s0 + 1; c < s1; ++c)
      if (lg[c] > m) { m = lg[c]; am = c; }
    float se = 0.f;
    for (int c = s0; c < s1; ++c) se += expf(lg[c] - m);
    const long long yv = y[(long)b * slots.n_slots + tid];
    const bool in_range = yv >= 0 && yv < (long long)(s1 - s0);      // F.cross_entropy would device-assert; here the loss turns NaN
    const int target = s0 + (in_range ? (int)yv : 0);
    slot_loss[tid] = in_range ? (m + logf(se)) - lg[target] : __int_as_float(0x7fc00000);
    slot_ok[tid] = in_range && am == target;
  }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f;
    int ok = 1;
    for (int s = 0; s < slots.n_slots; ++s) { l += slot_loss[s]; ok &= slot_ok[s]; }
    row_loss[b] = l;
    row_ok[b] = (float)ok;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float l = 0.f, a = 0.f;                                       // fixed order: thread-strided partials, then a tree
  for (int i = tid; i < B; i += 256) { l += __ldcg(row_loss + i); a += __ldcg(row_ok + i); }
  red[0][tid] = l;
  red[1][tid] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) { red[0][tid] += red[0][tid + s]; red[1][tid] += red[1][tid + s]; }
    __syncthreads();
  }
  if (tid == 0) {
    loss_acc[0] = red[0][0] / (float)B;
    loss_acc[1] = red[1][0] / (float)B;
    *ticket = 0u;                                               // ready for the next launch
  }
}

// ---- backward: thread j owns feature column j; a CTA walks utterances b, b + grid, ... -----------------------------
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ gloss, const float* __restrict__ feats,
                                                       const float* __restrict__ W, const long long* __restrict__ y,
                                                       const float* __restrict__ logits, const int* __restrict__ tstar, int B, int T,
                                                       int C, SlotTable slots, float* __restrict__ dfeats, float* __restrict__ dW,
                                                       float* __restrict__ dbias) {
  extern __shared__ float sm[];
  float* dw_s = sm;                          // [C][256] this CTA's partial dW
  float* dl = dw_s + C * HEAD_F;             // [C] dL/dlogit of the current utterance
  int* ts = (int*)(dl + HEAD_MAXC);          // [C]
  const int tid = threadIdx.x;
  for (int i = tid; i < C * HEAD_F; i += 256) dw_s[i] = 0.f;
  float db = 0.f;
  const float g = y ? gloss[0] / (float)B : 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    if (y == nullptr) {                      // logits-only head (predict path): `gloss` is dL/dlogits [B][C]
      if (tid < C) dl[tid] = gloss[(long)b * C + tid];
    } else if (tid < slots.n_slots) {
      const int s0 = slots.start[tid], s1 = slots.start[tid + 1];
      const float* lg = logits + (long)b * C;
      float m = lg[s0];
      for (int c = s0 + 1; c < s1; ++c) m = fmaxf(m, lg[c]);
      float se = 0.f;
      for (int c = s0; c < s1; ++c) se += expf(lg[c] - m);
      const float inv = 1.f / se;
      const long long yv = y[(long)b * slots.n_slots + tid];
      const float poison = (yv >= 0 && yv < (long long)(s1 - s0)) ? 0.f : __int_as_float(0x7fc00000);   // out-of-range label -> NaN grads
      const int target = s0 + (int)yv;
      for (int c = s0; c < s1; ++c) dl[c] = g * (expf(lg[c] - m) * inv - (c == target ? 1.f : 0.f)) + poison;
    }
    if (tid < C) ts[tid] = tstar[(long)b * C + tid];
    __syncthreads();
    if (tid < C) db += dl[tid];
    float* drow = dfeats + (long)b * T * HEAD_F + tid;
    const float* frow = feats + (long)b * T * HEAD_F + tid;
    for (int t = 0; t < T; ++t) drow[(long)t * HEAD_F] = 0.f;
    for (int c = 0; c < C; ++c) {            // only the arg-max frame of each class receives gradient
      const int t = ts[c];
      const float d = dl[c];
      drow[(long)t * HEAD_F] += d * __ldg(W + c * HEAD_F + tid);
      dw_s[c * HEAD_F + tid] += d * __ldg(frow + (long)t * HEAD_F);
    }
  }
  for (int c = 0; c < C; ++c) atomicAdd(dW + c * HEAD_F + tid, dw_s[c * HEAD_F + tid]);
  if (tid < C) atomicAdd(dbias + tid, db);
}

int make_slots(const int* values_per_slot, int n_slots, int C, SlotTable* st) {
  if (n_slots < 1 || n_slots > HEAD_MAXS || C < 1 || C > HEAD_MAXC) return (int)cudaErrorInvalidValue;
  st->n_slots = n_slots;
  st->start[0] = 0;
  for (int s = 0; s < n_slots; ++s) {
    if (values_per_slot[s] < 1) return (int)cudaErrorInvalidValue;
    st->start[s + 1] = st->start[s] + values_per_slot[s];
  }
  return st->start[n_slots] == C ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace

extern "C" int slu_intent_head_fwd(const float* feats, const float* W, const float* bias, const long long* y, int B, int T, int C,
                                   const int* values_per_slot, int n_slots, float* logits, int* tstar, float* row_loss,
                                   float* row_ok, float* loss_acc, unsigned int* ticket, void* stream) {
  if (B <= 0 || T <= 0) return (int)cudaErrorInvalidValue;
  SlotTable st;
  if (int e = make_slots(values_per_slot, n_slots, C, &st)) return e;
  const size_t smem = sizeof(float) * ((size_t)((C * 257 + 3) & ~3) + HEAD_TT * HEAD_F + (size_t)HEAD_TT * C + HEAD_MAXC);
  if (smem > 48 * 1024)
    if (int e = slu_set_smem((const void*)head_fwd_kernel, smem)) return e;
  head_fwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(feats, W, bias, y, B, T, C, st, logits, tstar, row_loss, row_ok, loss_acc,
                                                          ticket);
  SLU_CHECK_LAUNCH();
  return 0;
}

extern "C" int slu_intent_head_bwd(const float* gloss, const float* feats, const float* W, const long long* y, const float* logits,
                                   const int* tstar, int B, int T, int C, const int* values_per_slot, int n_slots, float* dfeats,
                                   float* dW, float* dbias, void* stream) {
  if (B <= 0 || T <= 0) return (int)cudaErrorInvalidValue;
  SlotTable st;
  if (int e = make_slots(values_per_slot, n_slots, C, &st)) return e;
  const size_t smem = sizeof(float) * ((size_t)C * HEAD_F + HEAD_MAXC) + sizeof(int) * HEAD_MAXC;
  if (smem > 48 * 1024)
    if (int e = slu_set_smem((const void*)head_bwd_kernel, smem)) return e;
  const int grid = B < 148 ? B : 148;
  head_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(gloss, feats, W, y, logits, tstar, B, T, C, st, dfeats, dW, dbias);
  SLU_CHECK_LAUNCH();
  return 0;
}
