// Backward of one bidirectional GRU layer as ONE host call: the launch sequence that ops.BiGRU.backward otherwise drives from
// Python (autograd of nn.GRU at reference models.py:232/262/686).  Nothing new runs on the device -- it is the same kernels in
// the same order -- but the host pays one C-ABI call per layer instead of eight, which matters because a train step is only a
// few hundred microseconds away from being host-bound (multi-GPU runs, per-step result reads).
//   1. slu_gru_bwd_tc            recurrence backward -> dgx [B][T][768], dhn [B][T][256], db_ih / db_hh
//   2. fork: slu_wgrad_tc        dW_ih = dgx^T . x                      (side stream 0)
//            slu_wgrad2_tc x2    dW_hh[d] = [dr,dz | dhn]^T . h_{t-+1}  (side streams 1, 2)
//   3. slu_gemm_tc               dX = dgx . W_ih (pre-split operand image), on the caller's stream
//   4. join
#include "common.cuh"
#include "../../include/slu_b200.h"

extern "C" int slu_bigru_bwd_tc(const float* gy, const float* drop_mask, float drop_p, unsigned long long drop_seed,
                                const unsigned long long* drop_seed_dev, const float* y_full, const float* stash, const float* w_hh,
                                const float* x, int I, const void* w_ih_nn_img, int B, int T, int ds, float* dgx, float* dhn,
                                float* db_ih, float* db_hh, float* dw_ih, float* dw_hh, float* dx, int overlap, void* stream) {
  if (B <= 0 || T <= 0 || I <= 0 || !dgx || !dhn) return (int)cudaErrorInvalidValue;
  if (dx && !w_ih_nn_img) return (int)cudaErrorInvalidValue;
  if ((dw_ih == nullptr) != (dw_hh == nullptr)) return (int)cudaErrorInvalidValue;
  int e = slu_gru_bwd_tc(gy, drop_mask, drop_p, drop_seed, drop_seed_dev, y_full, stash, w_hh, B, T, ds, dgx, dhn, db_ih, db_hh, stream);
  if (e) return e;
  void* side[3] = {stream, stream, stream};
  const int ns = (dw_ih && overlap) ? 3 : 0;
  if (ns && (e = slu_stream_fork(stream, ns, side))) return e;
  if (dw_ih) {
    if ((e = slu_wgrad_tc(dgx, 768, 768, x, I, I, B, T, 1, 0, dw_ih, I, 1, 0, side[0]))) return e;
    for (int d = 0; d < 2; ++d) {       // rows [dr, dz] of direction d from dgx, rows dhn from dhn; h shifted against the direction
      e = slu_wgrad2_tc(dgx + d * SLU_G3, 768, 256, dhn + d * SLU_H, 256, SLU_G3, y_full + d * SLU_H, 256, SLU_H, B, T, 1, d ? 1 : -1,
                        dw_hh + (size_t)d * SLU_G3 * SLU_H, SLU_H, 1, 0, side[1 + d]);
      if (e) return e;
    }
  }
  if (dx && (e = slu_gemm_tc(dgx, 768, w_ih_nn_img, nullptr, dx, I, B * T, I, 768, 1, 0, 0, 0, 0.f, stream))) return e;
  if (ns && (e = slu_stream_join(stream, ns))) return e;
  return 0;
}
