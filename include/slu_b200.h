/* slu_b200 -- C ABI of the Blackwell-native (sm_100a) speech-encoder hot path.
 *
 * Drop-in boundary: the reference (lorenlugosch/end-to-end-SLU) has no FFI; its hot path is a chain
 * of PyTorch library calls inside models.py.  Each entry point below replaces the library calls cited
 * beside it (reference models.py:<line>); the repo-root models.py mirrors the reference class surface
 * and reaches these through ctypes (end-to-end-slu_b200/_lib.py).  See INTEGRATION.md.
 *
 * Conventions: plain device pointers + sizes, no torch types; every function enqueues work on `stream`
 * (a cudaStream_t passed as void*), returns 0, a cudaError_t or SLU_ERR_TOO_LARGE, never synchronises, never allocates.
 * All tensors are contiguous row-major fp32 unless stated.  H = 128 hidden units, gate order (r, z, n).
 */
#ifndef SLU_B200_H
#define SLU_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Returned (instead of a cudaError_t) when a size exceeds a kernel's 32-bit index range, e.g. B*T >= 2^21 frames in
 * slu_gru_{fwd,bwd}_tc: split the batch. */
#define SLU_ERR_TOO_LARGE 100001

/* Filter-bank synthesis W[80][401] from the fp64 cut-offs -- replaces the 80-iteration python loop
 * models.py:82-106 (sinc() :17-24, flip() :7-14). */
int slu_sinc_filters_fwd(const double* filt_b1, const double* filt_band, float* W, void* stream);
/* Its analytic backward: dL/dW[80][401] -> dL/dfilt_b1[80], dL/dfilt_band[80] (fp64), including the
 * max-normalisation (:103) and |.| (:88-89) -- replaces autograd over those ops. */
int slu_sinc_filters_bwd(const double* filt_b1, const double* filt_band, const float* dW, double* d_b1, double* d_band,
                         void* stream);

/* Jacobian banks J[2][80][401] of the synthesis: J[0] = dW/d filt_b1, J[1] = dW/d filt_band (per filter row; the
 * max-normalisation and |.| terms included, evaluated in fp64) -- the operand of slu_sincconv_bwd_jac_tc. */
int slu_sinc_filters_jac(const double* filt_b1, const double* filt_band, float* J, void* stream);

/* conv1d(x[B][1][T], W, stride 80, pad 200) + Abs + MaxPool1d(2, ceil) -- replaces models.py:108, :163-168, :205
 * (LeakyReLU :211 is the identity on the non-negative result, Dropout(0) :218 likewise).
 * out[B][L1][80] (time-major), L0=(T-1)/80+1, L1=(L0+1)/2.  route[B][L1][80] (u8, may be NULL): bit0 = which
 * frame of the pair won, bit1 = sign of the winner, bit2 = winner was exactly 0 (abs has zero gradient). */
int slu_sincconv_fwd_simt(const float* x, const float* W, int B, int T, float* out, uint8_t* route, void* stream);
/* dL/dW[80][401] from dL/dout -- replaces cuDNN's conv backward-filter + abs/max-pool autograd.  The waveform
 * needs no gradient. */
int slu_sincconv_bwd_simt(const float* x, const float* gy, const uint8_t* route, int B, int T, float* dW, void* stream);

/* Same contracts on tcgen05 tensor cores: the waveform is viewed as a [frames][80] matrix and the 401-tap filter as 6
 * accumulating K=80 taps (no im2col), bf16 hi/lo 3-pass split, abs + max-pool + route bits in the GEMM epilogue.
 * `img` = scratch for the pre-split bank (2*6*80*96 bf16 values).  slu_sincconv_bwd_tc needs dW zero-filled (split-K atomics). */
int slu_sincconv_fwd_tc(const float* x, const float* W, int B, int T, float* out, uint8_t* route, void* img, void* stream);
int slu_sincconv_bwd_tc(const float* x, const float* gy, const uint8_t* route, int B, int T, float* dW, void* stream);
/* The cut-off gradients directly (no dW): d[0..79] += dL/d filt_b1, d[80..159] += dL/d filt_band (fp64, caller-zeroed) =
 * sum over frames of the routed output gradient times the convolution of the waveform with the Jacobian banks J
 * (slu_sinc_filters_jac) -- same tcgen05 kernel as the forward, two stacked banks, reducing epilogue.  The cancellation between
 * the direct and the max-normalisation term happens analytically inside J, so bf16 hi/lo operands keep fp32-class accuracy
 * (the dW route loses ~3 digits there).  `img` = scratch for the pre-split banks (2*6*160*96 bf16 values). */
/* Both run as ONE persistent CTA per SM walking the 128-frame tiles with stager / MMA / filter-bank TMA / epilogue warps and two
 * TMEM accumulators (staging of tile i+1 and the epilogue of tile i-1 overlap the MMAs of tile i); slu_set_sinc_persistent(0)
 * selects the one-CTA-per-tile kernel instead (A/B measurements). */
int slu_set_sinc_persistent(int on);
/* Developer tool: CTA (0,0) of the persistent kernel records clock64() at its hand-off points into buf[16 tiles][8] (or NULL: off). */
int slu_debug_sinc_trace(long long* buf);
int slu_sincconv_bwd_jac_tc(const float* x, const float* gy, const uint8_t* route, const float* J, int B, int T, double* d,
                            void* img, void* stream);

/* Persistent bidirectional GRU recurrence (h0 = 0) with fused gate non-linearities, Dropout mask multiply and
 * Downsample -- replaces nn.GRU (_VF.gru / cuDNN RNN) at models.py:232/262/686 plus RNNSelect :138-149,
 * Dropout :246/276/700 and Downsample :26-46.
 *   gx    [B][T][768]  x.W_ih^T + b_ih for both directions, col = d*384 + g*128 + j
 *   w_hh  [2][384][128], b_hh [2][384]
 *   drop_mask [B][T][256] keep-mask scaled by 1/(1-p), or NULL.  With drop_mask == NULL and drop_p > 0 the kernels generate the
 *             canonical Philox mask of (drop_p, drop_seed) in registers (csrc/philox.cuh; slu_dropout_mask_gru writes the same
 *             mask out as a tensor): no mask tensor exists in HBM and the backward kernel regenerates it.  drop_p = 0: no dropout.
 *             drop_seed_dev (may be NULL): one device word XOR-ed into drop_seed at run time -- a train step captured as a CUDA
 *             graph freezes its by-value arguments, slu_seed_advance(word) as the graph's first node gives every replay new masks.
 *   ds    1 = Downsample("none",1), 2 = Downsample("avg",2) (ceil mode: an odd tail frame is kept as is)
 *   y_full [B][T][256] raw hidden states (col = d*128 + j);  y_out [B][ceil(T/ds)][256]
 *   stash [B][T][1024] (col = d*512 + 4*j + s: r, z, n, W_hn h + b_hn of unit j, direction d, adjacent) for the backward pass, or NULL
 *         for inference. */
int slu_gru_fwd_simt(const float* gx, const float* w_hh, const float* b_hh, const float* drop_mask, float drop_p,
                     unsigned long long drop_seed, const unsigned long long* drop_seed_dev, int B, int T, int ds, float* y_full,
                     float* y_out, float* stash, void* stream);
/* Backward through time -- replaces _cudnn_rnn_backward.  Emits dgx[B][T][768] (gradient wrt gx) and
 * dhn[B][T][256] (gradient wrt the n-gate's recurrent pre-activation); the weight/input gradients are dense
 * GEMMs over these.  db_ih[2][384] and db_hh[2][384] (the bias parameters' own layout; both NULL or both caller-zeroed)
 * ACCUMULATE the bias gradients: b_ih <- sums over (b,t) of (dr, dz, dn), b_hh <- (dr, dz, dhn). */
int slu_gru_bwd_simt(const float* dy_out, const float* drop_mask, float drop_p, unsigned long long drop_seed,
                     const unsigned long long* drop_seed_dev, const float* y_full, const float* stash, const float* w_hh, int B, int T, int ds, float* dgx, float* dhn, float* db_ih, float* db_hh,
                     void* stream);

/* Same contracts as slu_gru_fwd_simt / slu_gru_bwd_simt, executed on tcgen05 tensor cores: W_hh (bf16 hi+lo) stationary
 * in tensor memory, h / dG as the shared-memory B operand, 3-pass bf16 split with fp32 accumulation in TMEM. */
int slu_gru_fwd_tc(const float* gx, const float* w_hh, const float* b_hh, const float* drop_mask, float drop_p,
                   unsigned long long drop_seed, const unsigned long long* drop_seed_dev, int B, int T, int ds, float* y_full,
                   float* y_out, float* stash, void* stream);
int slu_gru_bwd_tc(const float* dy_out, const float* drop_mask, float drop_p, unsigned long long drop_seed,
                   const unsigned long long* drop_seed_dev, const float* y_full, const float* stash, const float* w_hh, int B, int T, int ds, float* dgx, float* dhn, float* db_ih, float* db_hh,
                   void* stream);

/* Operand format of the tcgen05 recurrence:
 *   0 = bf16 hi/lo split, fp32-class accuracy (default).  On the 4- and 8-row batch tiles the hi and lo rows of the activation
 *       tile are stacked along the MMA's N dimension, so 2 MMAs per K step (W_hi, W_lo) accumulate all four partial products;
 *       the 16-row tile issues hi*hi + hi*lo + lo*hi as three passes;
 *   1 = one fp16 pass (11-bit operands; intent logits stay within the 1e-3 parity tolerance);
 *   2 = three separate bf16 passes on every tile (un-stacked form, for A/B checks). */
int slu_set_gru_precision(int mode);
/* Batch rows one CTA of slu_gru_{fwd,bwd}_tc carries for a batch of B utterances (4, 8 or 16; host-side query, no launch). */
int slu_gru_rows_per_cta(int B);

/* Developer tool: accumulate clock64() per step phase of slu_gru_fwd_tc (CTA 0, threads 0 and 128) into buf[2][8]. */
int slu_debug_gru_phase_clocks(long long* buf);

/* Intent head (models.py:709 Linear(256 -> C), :112-123 FinalPool max over time, :811-823 summed per-slot cross-entropy and
 * all-slots-right accuracy), one launch per direction.  feats [B][T][256]; W [C][256]; y [B][n_slots] int64 class indices
 * (NULL: logits only); values_per_slot = n_slots HOST ints summing to C (C <= 128, n_slots <= 16).
 * fwd writes logits [B][C], tstar [B][C] (arg-max frame), row_loss/row_ok [B] scratch and loss_acc[2] = {loss, accuracy};
 * `ticket` is one zero-initialised device word the kernel uses and resets.
 * bwd: gloss = dL/dloss (one device float); writes dfeats [B][T][256], ACCUMULATES into dW [C][256] and dbias [C].
 * With y == NULL (the logits-only head of predict_intents) gloss is dL/dlogits [B][C] instead.
 * A label outside [0, values_per_slot[slot]) makes the loss and the gradients NaN (F.cross_entropy would device-assert). */
int slu_intent_head_fwd(const float* feats, const float* W, const float* bias, const long long* y, int B, int T, int C,
                        const int* values_per_slot, int n_slots, float* logits, int* tstar, float* row_loss, float* row_ok,
                        float* loss_acc, unsigned int* ticket, void* stream);
int slu_intent_head_bwd(const float* gloss, const float* feats, const float* W, const long long* y, const float* logits,
                        const int* tstar, int B, int T, int C, const int* values_per_slot, int n_slots, float* dfeats, float* dW,
                        float* dbias, void* stream);

/* Backward of one bidirectional GRU layer in one host call (the launch sequence of slu_gru_bwd_tc, slu_wgrad_tc (dW_ih),
 * 2 x slu_wgrad2_tc (dW_hh per direction), slu_gemm_tc (dX) with the weight-gradient launches forked to side streams when
 * overlap != 0).  x [B][T][I] = the layer input; w_ih_nn_img = slu_presplit_bf16 image of W_ih [768][I] read as the [K=768][N=I]
 * operand (NULL with dx == NULL: no input gradient); dw_ih [768][I] and dw_hh [2][384][128] accumulate (NULL, NULL: no weight
 * gradients); dgx [B][T][768] and dhn [B][T][256] are caller-provided scratch that holds the pre-activation gradients. */
int slu_bigru_bwd_tc(const float* gy, const float* drop_mask, float drop_p, unsigned long long drop_seed,
                     const unsigned long long* drop_seed_dev, const float* y_full,
                     const float* stash, const float* w_hh, const float* x,
                     int I, const void* w_ih_nn_img, int B, int T, int ds, float* dgx, float* dhn, float* db_ih, float* db_hh,
                     float* dw_ih, float* dw_hh, float* dx, int overlap, void* stream);

/* Fork / join of independent launches (host-side stream plumbing, no kernels): after slu_stream_fork the n (<= 8)
 * streams returned in side_streams[] wait for everything queued on main_stream so far; after slu_stream_join work queued
 * on main_stream waits for everything queued on those n side streams.  Used to run one layer's weight-gradient GEMMs
 * next to its input-gradient GEMM.  Buffers must stay alive until the join has been passed on main_stream. */
int slu_stream_fork(void* main_stream, int n, void** side_streams);
int slu_stream_join(void* main_stream, int n);

/* Batch staging (replaces the synchronous `x = x.cuda()` inside forward, models.py:300-303): slu_h2d_async queues a host -> device
 * copy on the library's copy stream (order_after != 0: only after everything queued on after_stream so far -- for recycled
 * destination buffers); slu_h2d_ready makes consumer_stream wait for all copies queued so far.  src should be pinned. */
int slu_h2d_async(void* dst, const void* src, size_t bytes, void* after_stream, int order_after);
int slu_h2d_ready(void* consumer_stream);
/* Lifetime of the pinned sources: *pending = 1 while copies queued with slu_h2d_async are still in flight (non-blocking query);
 * slu_h2d_wait blocks the host until all of them have completed. */
int slu_h2d_pending(int* pending);
int slu_h2d_wait(void);

/* Dropout keep-mask (nn.Dropout, models.py:246/276/700, training mode): mask[i] = Bernoulli(1-p) / (1-p), i < n, from
 * Philox4x32-10 keyed by `seed` (counter = i/4).  `mask` must be 16-byte aligned.  The GRU kernels multiply by it. */
int slu_dropout_mask(float* mask, long n, float p, unsigned long long seed, void* stream);
/* The canonical GRU-layer mask [B][T][256] of (p, seed): exactly what slu_gru_{fwd,bwd}_* generate in registers when they are
 * given (drop_p, drop_seed) and no mask tensor. */
int slu_dropout_mask_gru(float* mask, int B, int T, float p, unsigned long long seed, void* stream);

/* Backward of the conv blocks' LeakyReLU plus the Conv1d bias gradient in one pass (autograd of models.py:200/211):
 *   dpre[r][c] = y[r][c] > 0 ? gy[r][c] : slope*gy[r][c];  db[c] += sum_r dpre[r][c]     (R rows of C floats, C % 4 == 0) */
int slu_leaky_bwd_bias(const float* y, const float* gy, float slope, float* dpre, float* db, long R, int C, void* stream);

/* Dense "tap-GEMM" on tcgen05 (fp32 in/out, 3-pass bf16 split, fp32 accumulate in TMEM) -- replaces the cuBLAS / cuDNN
 * calls behind nn.GRU's input projection (models.py:232/262/686), nn.Conv1d (models.py:200) and their input gradients:
 *   C[m][n] = sum_tap sum_k A[(m + tap - tap_pad)*lda + k] * W(n, tap, k) (+ bias[n]) (LeakyReLU(slope) if act == 1)
 * Rows are frames of utterances of T frames (T = 0: no boundary); rows shifted out of their utterance read as 0 (conv padding).
 * w_img = the weight operand pre-split by slu_presplit_bf16. */
int slu_gemm_tc(const float* A, long lda, const void* w_img, const float* bias, float* C, long ldc, int M, int N, int K, int taps,
                int tap_pad, int T, int act, float slope, void* stream);
/* Weights W(n, tap, k) = W[n*sn + k*sk + tap*stap] (any strides, so transposed / reversed-tap views cost nothing) -> bf16 hi/lo
 * image [2][taps][N][Kp], Kp = K rounded up to 32 (img: 2*taps*N*Kp bf16 values). */
int slu_presplit_bf16(const float* W, long sn, long sk, long stap, int taps, int N, int K, void* img, void* stream);
/* n <= 16 such jobs in ONE launch; `jobs` = host array of struct { const float* W; long sn, sk, stap; int taps, N, K, pad; void* img; }. */
int slu_presplit_multi(const void* jobs, int n, void* stream);

/* Weight-gradient GEMM (reduction over frames) with MN-major tcgen05 operands and TMEM-resident accumulators:
 *   out[m*s_m + n*s_n + tap*s_tap] += sum_{b<B, t<T} G[(b*T+t)][m] * X[(b*T + t + shift0 + tap)*ldx + n]
 * (frames outside [0,T) read 0; taps in {1, 5}).  Replaces the backward-weights kernels behind autograd of nn.GRU
 * (dW_ih, dW_hh with shift0 = -1/+1) and nn.Conv1d (5 taps, shift0 = -2, out in the [Cout][Cin][5] weight layout).
 * slu_wgrad2_tc takes the rows of G from two tensors: G[..][m] = G0[(b*T+t)*ldg0 + m] for m < m_split, else
 * G1[(b*T+t)*ldg1 + m - m_split] (dW_hh = [dr, dz | dhn]^T . h in one launch).  Operands 16-byte aligned; ldg*, ldx, M, N,
 * m_split multiples of 4. */
int slu_wgrad2_tc(const float* G0, long ldg0, int m_split, const float* G1, long ldg1, int M, const float* X, long ldx, int N, int B,
                  int T, int taps, int shift0, float* out, long s_m, long s_n, long s_tap, void* stream);
int slu_wgrad_tc(const float* G, long ldg, int M, const float* X, long ldx, int N, int B, int T, int taps, int shift0, float* out,
                 long s_m, long s_n, long s_tap, void* stream);
/* Developer tool (tools/wgrad_only.py): bit 0 skips the MMAs, bit 1 the operand conversion, bit 2 the flush; 0 = normal. */
int slu_debug_wgrad_mode(int mode);
int slu_debug_gemm_mode(int mode);      /* same for slu_gemm_tc: 1 no MMAs, 2 no conversion, 4 no global stores, 8 no epilogue */
/* Developer tool: CTA (0,0,0) records clock64() at its hand-off points into buf[64 tiles][8] (NULL: off). */
int slu_debug_wgrad_trace(long long* buf);

/* tcgen05 self-test: C[128][N] = A[128][K] . B[N][K]^T (3-pass bf16 split, fp32 accumulate in TMEM). */
/* ---- ASR heads: frame-wise cross-entropy without materialising [B*T', V] logits (reference models.py:308-314, 321-329:
 * Linear(256 -> V) -> view(B*T', V) -> F.cross_entropy(ignore_index=-1) + masked arg-max accuracy).  The caller walks the frames
 * in row chunks: slu_gemm_tc writes a chunk's logits tile, slu_ce_rows turns it IN PLACE into dL/dlogits (softmax - one-hot,
 * scaled by 1/n_valid; rows with y == -1 become 0) and emits per-row loss / hit flags, the tile then feeds slu_gemm_tc (input
 * gradient), slu_wgrad_tc (weight gradient) and slu_colsum_acc (bias gradient).
 *   slu_ce_count : nvalid[0] = #(y != -1), nvalid[1] = 1 / nvalid[0]                       (y: M int64 labels)
 *   slu_ce_rows  : logits [R][ld] (V <= ld, V <= 12288), y [R]; write_grad = 0 leaves the tile untouched (loss only).
 *                  A label outside [0, V) other than -1 makes that row's loss NaN.
 *   slu_ce_finish: loss_acc[0] = sum(row_loss) / n_valid, loss_acc[1] = sum(row_ok) / n_valid, fixed summation order.
 *   slu_colsum_acc: out[c] += sum_r A[r*ld + c], c < C.      slu_scale: dst[i] = src[i] * g[0] (g on the device). */
int slu_ce_count(const long long* y, long M, float* nvalid, void* stream);
int slu_ce_rows(float* logits, long ld, int V, const long long* y, long R, const float* nvalid, int write_grad, float* row_loss,
                float* row_ok, void* stream);
int slu_ce_finish(const float* row_loss, const float* row_ok, long M, const float* nvalid, float* loss_acc, void* stream);
int slu_colsum_acc(const float* A, long ld, long R, int C, float* out, void* stream);
int slu_scale(const float* src, float* dst, long n, const float* g, void* stream);

/* ---- seq2seq intent decoder, teacher-forced training path (reference models.py:413-436 Attention, 438-484 DecoderRNN,
 * 500-556 Seq2SeqDecoder.forward).  Everything outside the recurrence runs as dense slu_gemm_tc / slu_wgrad_tc calls over all
 * output symbols at once; per symbol the sequential part is these kernels between four small GEMMs (see csrc/decoder.cu).
 *   slu_attn_step_fwd: q [B][ldq] (K used), keys [B][T][K], values [B][T][V] -> w [B][T] = softmax_t(keys.q * inv_scale),
 *                      ctx [B][V] = sum_t w[t] values[t]           (T <= 256, K, V <= 512)
 *   slu_attn_step_bwd: dctx [B][V] -> dq [B][lddq]; dkeys, dvalues ACCUMULATE (one launch per symbol, walked backwards)
 *   slu_grucell_fwd  : torch.nn.GRUCell gate math on precomputed gi = gi_a (+ gi_b) and gh (rows of 3D: r | z | n, biases
 *                      included): h = (1-z) n + z hprev; hprev == NULL: the row h0[D] for every utterance (initial state);
 *                      stash [B][4D] = r | z | n | gh_n; dropped [B][D] (may be NULL) = h * Philox keep-mask(drop_p, drop_seed, step)
 *   slu_grucell_bwd  : dh = da * mask(step) + db + dc (db, dc may be NULL) -> dgi [B][ldgi], dgh [B][ldgh], dh_direct = dh * z
 *   slu_skinny_gemm  : C[m][n] = sum_k A[m*lda + k] * W[n*sn + k*sk] (+ bias[n]) for M <= 64 rows: the per-symbol projections are
 *                      latency-bound (a persistent tensor-core pipeline costs more to start than they take): exact-fp32 CUDA cores. */
int slu_skinny_gemm(const float* A, long lda, const float* W, long sn, long sk, const float* bias, float* C, long ldc, int M, int N, int K,
                    void* stream);
int slu_attn_step_fwd(const float* q, long ldq, const float* keys, const float* values, int B, int T, int K, int V, float inv_scale,
                      float* w, float* ctx, void* stream);
int slu_attn_step_bwd(const float* dctx, const float* w, const float* q, long ldq, const float* keys, const float* values, int B, int T,
                      int K, int V, float inv_scale, float* dq, long lddq, float* dkeys, float* dvalues, void* stream);
int slu_grucell_fwd(const float* gi_a, long lda, const float* gi_b, long ldb, const float* gh, long ldh, const float* hprev,
                    const float* h0, int B, int D, float drop_p, unsigned long long drop_seed, const unsigned long long* drop_seed_dev,
                    int step, float* h, float* stash, float* dropped, void* stream);
int slu_grucell_bwd(const float* da, const float* db, const float* dc, const float* stash, const float* hprev, const float* h0, int B,
                    int D, float drop_p, unsigned long long drop_seed, const unsigned long long* drop_seed_dev, int step, float* dgi,
                    long ldgi, float* dgh, long ldgh, float* dh_direct, void* stream);

/* ---- optimizer step / gradient bucket (reference training.py:19, 64-66, 96-98: torch.optim.Adam, zero_grad/backward/step) ----
 * slu_adam_multi: one Adam step over `n` parameter tensors (fp32 or fp64, any sizes) in ceil(n/64) launches; torch.optim.Adam's
 * single-tensor arithmetic with PER-TENSOR step counts (parameters un-frozen later keep their own bias corrections):
 *   g += weight_decay*p;  m += (g - m)(1 - beta1);  v = v*beta2 + (1 - beta2) g^2;  p -= step_size * m / (sqrt(v)/bc2_sqrt + eps)
 * `tensors` is a HOST array (it travels in the kernel parameters). */
struct SluAdamTensor {
  void* p; const void* g; void* m; void* v;   /* parameter, gradient, exp_avg, exp_avg_sq (device pointers, same dtype) */
  long n;                                     /* elements */
  float step_size;                            /* lr / (1 - beta1^t) */
  float bc2_sqrt;                             /* sqrt(1 - beta2^t) */
  int is_f64;                                 /* 0: float, 1: double */
  int pad;
};
int slu_seed_advance(unsigned long long* state, void* stream);   /* state[0] <- LCG(state[0]): see drop_seed_dev above */
int slu_adam_multi(const void* tensors, int n, double beta1, double beta2, float eps, float weight_decay, void* stream);
/* fp64 gradients inside an fp32 all-reduce bucket: split into (hi, lo) floats before the collective, merge after it. */
int slu_f64_hilo_split(const double* src, float* hi, float* lo, int n, void* stream);
int slu_f64_hilo_merge(double* dst, const float* hi, const float* lo, int n, void* stream);

int slu_tc_selftest(const float* A, const float* B, float* C, int N, int K, void* stream);
/* Same, A operand resident in tensor memory (K <= 128), B tile with a padded leading-byte-offset. */
int slu_tc_selftest_ts(const float* A, const float* B, float* C, int N, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif
