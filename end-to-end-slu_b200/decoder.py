"""Teacher-forced seq2seq intent decoder on the library's kernels.

Reference: models.py:413-436 (Attention), 438-484 (DecoderRNN = GRUCell stack + Dropout), 500-556 (Seq2SeqDecoder.forward).
The reference runs ~25 library launches per output symbol forward (re-projecting the encoder states to keys / values inside
every step) and as many again in autograd.  Here
  * everything outside the recurrence is one dense tcgen05 GEMM over ALL symbols: keys, values, the embeddings of the previous
    symbols, the embedding half of the first cell's input projection, and the output projection fused with log-softmax + NLL
    (ops.LinearNT / ops.LinearCE -- the output logits [B*U, |S|] are never needed as a tensor);
  * the recurrence itself is `DecoderStates`: per symbol 4 small projections (slu_skinny_gemm up to 64 utterances, slu_gemm_tc beyond) + 3 fused kernels (csrc/decoder.cu:
    attention step, two GRUCell gate kernels with the inter-cell Dropout), with a hand-written backward through time whose weight
    gradients are 4 weight-gradient GEMMs over all symbols at the end (slu_wgrad_tc) instead of per-step accumulations.
Two cells (num_intent_decoder_layers = 2) of equal width, as in every reference cfg that enables seq2seq; other depths raise.
"""
import math

import torch

from . import _lib, ops

_f32 = ops._f32


SKINNY_MAX_ROWS = 64        # csrc/decoder.cu skinny_gemm_kernel: the per-symbol projections of up to 64 utterances


class _W:
    """One recurrent weight in both operand forms of the per-symbol projections: `nt` (x @ W^T) and `nn` (g @ W).  Up to 64
    utterances take the exact-fp32 skinny kernel straight on the weight (no operand image); larger batches the tcgen05 GEMM."""

    def __init__(self, w, need_nn, B):
        self.w = w
        self.skinny = B <= SKINNY_MAX_ROWS
        self.img_nt = self.img_nn = None
        if not self.skinny:
            self.img_nt = ops.presplit(w, *ops._form_nt(w))
            self.img_nn = ops.presplit(w, *ops._form_nn(w)) if need_nn else None

    def nt(self, x2, M, out, bias=None):          # out [M, N] = x2 [M, K] @ W[N, K]^T
        N, K = self.w.shape
        if self.skinny:
            _lib.call("slu_skinny_gemm", x2.data_ptr(), K, self.w.data_ptr(), K, 1, _lib.ptr(bias), out.data_ptr(), N, M, N, K, _lib.stream())
        else:
            ops.gemm_tc(x2, K, self.img_nt, M, N, K, out, bias=bias)

    def nn(self, g2, M, out):                     # out [M, K] = g2 [M, N] @ W[N, K]
        N, K = self.w.shape
        if self.skinny:
            _lib.call("slu_skinny_gemm", g2.data_ptr(), N, self.w.data_ptr(), 1, K, None, out.data_ptr(), K, M, K, N, _lib.stream())
        else:
            ops.gemm_tc(g2, N, self.img_nn, M, K, N, out)


class DecoderStates(torch.autograd.Function):
    """keys [B,T,K], values [B,T,V], ge_all [U,B,3D] (= embed(y_prev) W_ie^T + b_ih0), initial_state [2,D] and the recurrent
    weights -> top-layer states after every symbol [U,B,D] (what the output projection reads; models.py:520-556)."""

    @staticmethod
    def forward(ctx, keys, values, ge_all, init_state, wq, bq, w_c, w_hh0, b_hh0, w_ih1, b_ih1, w_hh1, b_hh1, drop_p, drop_seed,
                seed_dev=None):
        keys, values, ge_all = _f32(keys), _f32(values), _f32(ge_all)
        B, T, K = keys.shape
        V = values.shape[2]
        U = ge_all.shape[0]
        D = w_hh0.shape[1]
        G = 3 * D
        assert ge_all.shape == (U, B, G) and init_state.shape == (2, D) and K % 4 == 0 and V % 4 == 0 and D % 4 == 0
        dev = keys.device
        st = _lib.stream()
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        wcat1 = torch.cat([w_hh1.detach(), wq.detach()], 0).contiguous()              # [3D + K, D]: s1 -> (gh1 | query)
        bcat1 = torch.cat([b_hh1.detach(), bq.detach()], 0).contiguous()
        w_c_, w_hh0_, w_ih1_ = w_c.detach().contiguous(), w_hh0.detach().contiguous(), w_ih1.detach().contiguous()
        N1 = G + K
        need = any(ctx.needs_input_grad)
        Wcat1, Whh0, Wc, Wih1 = (_W(w, need, B) for w in (wcat1, w_hh0_, w_c_, w_ih1_))
        s0 = f(U + 1, B, D); s1 = f(U + 1, B, D)                                        # [0] = initial state, [u+1] = after symbol u
        s0[0].copy_(init_state[0].detach().expand(B, D)); s1[0].copy_(init_state[1].detach().expand(B, D))
        g1 = f(U, B, N1)                                                                # gh1 | query of every step
        watt, ctxs, d0 = f(U, B, T), f(U, B, V), f(U, B, D)
        stash0, stash1 = f(U, B, 4 * D), f(U, B, 4 * D)
        gh0, gi0c, gi1 = f(B, G), f(B, G), f(B, G)
        inv_scale = 1.0 / math.sqrt(float(K))
        b_hh0_, b_ih1_ = b_hh0.detach().contiguous(), b_ih1.detach().contiguous()
        for u in range(U):
            Wcat1.nt(s1[u], B, g1[u], bcat1)                                            # s1 -> gh1 | q
            Whh0.nt(s0[u], B, gh0, b_hh0_)                                              # s0 -> gh0
            _lib.call("slu_attn_step_fwd", g1[u].data_ptr() + 4 * G, N1, keys.data_ptr(), values.data_ptr(), B, T, K, V, inv_scale,
                      watt[u].data_ptr(), ctxs[u].data_ptr(), st)
            Wc.nt(ctxs[u], B, gi0c)                                                     # context half of the first cell's input
            _lib.call("slu_grucell_fwd", ge_all[u].data_ptr(), G, gi0c.data_ptr(), G, gh0.data_ptr(), G, s0[u].data_ptr(), None, B, D,
                      float(drop_p), int(drop_seed), _lib.ptr(seed_dev), u, s0[u + 1].data_ptr(), stash0[u].data_ptr(), d0[u].data_ptr(), st)
            Wih1.nt(d0[u], B, gi1, b_ih1_)
            _lib.call("slu_grucell_fwd", gi1.data_ptr(), G, None, 0, g1[u].data_ptr(), N1, s1[u].data_ptr(), None, B, D,
                      0.0, 0, None, u, s1[u + 1].data_ptr(), stash1[u].data_ptr(), None, st)
        if need:
            ctx.save_for_backward(keys, values, s0, s1, g1, watt, ctxs, d0, stash0, stash1)
            ctx.weights = (Wcat1, Whh0, Wc, Wih1)
            ctx.drop = (float(drop_p), int(drop_seed), seed_dev)
            ctx.dims = (B, T, K, V, U, D)
        return s1[1:]

    @staticmethod
    def backward(ctx, ds_all):
        keys, values, s0, s1, g1, watt, ctxs, d0, stash0, stash1 = ctx.saved_tensors
        Wcat1, Whh0, Wc, Wih1 = ctx.weights
        drop_p, drop_seed, seed_dev = ctx.drop
        B, T, K, V, U, D = ctx.dims
        G, N1 = 3 * D, 3 * D + K
        dev = keys.device
        st = _lib.stream()
        ds_all = _f32(ds_all)
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        dkeys, dvalues = z(B, T, K), z(B, T, V)
        dg1, dgh0, dgi0, dgi1 = f(U, B, N1), f(U, B, G), f(U, B, G), f(U, B, G)          # pre-activation gradients of every step
        dd0, dctx = f(B, D), f(B, V)
        dir0, dir1 = f(2, B, D), f(2, B, D)                                             # direct dh paths (ping-pong over steps)
        rec0, rec1 = f(2, B, D), f(2, B, D)                                             # recurrent dh paths
        inv_scale = 1.0 / math.sqrt(float(K))
        have_next = False
        for u in range(U - 1, -1, -1):
            cur, nxt = u & 1, (u + 1) & 1
            # second cell: dh = dL/ds1[u] (+ what step u+1 sent back)
            _lib.call("slu_grucell_bwd", ds_all[u].data_ptr(), dir1[nxt].data_ptr() if have_next else None,
                      rec1[nxt].data_ptr() if have_next else None, stash1[u].data_ptr(), s1[u].data_ptr(), None, B, D, 0.0, 0, None, u,
                      dgi1[u].data_ptr(), G, dg1[u].data_ptr(), N1, dir1[cur].data_ptr(), st)
            Wih1.nn(dgi1[u], B, dd0)                                                    # -> d(dropped s0')
            # first cell: dh = dd0 * mask (+ what step u+1 sent back)
            _lib.call("slu_grucell_bwd", dd0.data_ptr(), dir0[nxt].data_ptr() if have_next else None,
                      rec0[nxt].data_ptr() if have_next else None, stash0[u].data_ptr(), s0[u].data_ptr(), None, B, D, drop_p, drop_seed,
                      _lib.ptr(seed_dev), u, dgi0[u].data_ptr(), G, dgh0[u].data_ptr(), G, dir0[cur].data_ptr(), st)
            Wc.nn(dgi0[u], B, dctx)
            _lib.call("slu_attn_step_bwd", dctx.data_ptr(), watt[u].data_ptr(), g1[u].data_ptr() + 4 * G, N1, keys.data_ptr(),
                      values.data_ptr(), B, T, K, V, inv_scale, dg1[u].data_ptr() + 4 * G, N1, dkeys.data_ptr(), dvalues.data_ptr(), st)
            Wcat1.nn(dg1[u], B, rec1[cur])                                              # (dgh1 | dq) -> ds1[u-1]
            Whh0.nn(dgh0[u], B, rec0[cur])                                              # dgh0 -> ds0[u-1]
            have_next = True
        # initial state: sum over the batch of what step 0 sent back (direct + recurrent)
        dinit = z(2, D)
        for l, (di, re) in enumerate(((dir0, rec0), (dir1, rec1))):
            _lib.call("slu_colsum_acc", di[0].data_ptr(), D, B, D, dinit[l].data_ptr(), st)
            _lib.call("slu_colsum_acc", re[0].data_ptr(), D, B, D, dinit[l].data_ptr(), st)
        # weight gradients: one reduction over all (symbol, utterance) rows per weight
        R = U * B
        wb = z(N1 * D + N1 + G * D + G + G * V + G * D + G)
        o = 0
        dwcat1 = wb[o:o + N1 * D].view(N1, D); o += N1 * D
        dbcat1 = wb[o:o + N1]; o += N1
        dw_hh0 = wb[o:o + G * D].view(G, D); o += G * D
        db_hh0 = wb[o:o + G]; o += G
        dw_c = wb[o:o + G * V].view(G, V); o += G * V
        dw_ih1 = wb[o:o + G * D].view(G, D); o += G * D
        db_ih1 = wb[o:o + G]
        ops.wgrad_tc(dg1, 0, N1, N1, s1, 0, D, D, 1, R, dwcat1, 0, D)
        ops.wgrad_tc(dgh0, 0, G, G, s0, 0, D, D, 1, R, dw_hh0, 0, D)
        ops.wgrad_tc(dgi0, 0, G, G, ctxs, 0, V, V, 1, R, dw_c, 0, V)
        ops.wgrad_tc(dgi1, 0, G, G, d0, 0, D, D, 1, R, dw_ih1, 0, D)
        _lib.call("slu_colsum_acc", dg1.data_ptr(), N1, R, N1, dbcat1.data_ptr(), st)
        _lib.call("slu_colsum_acc", dgh0.data_ptr(), G, R, G, db_hh0.data_ptr(), st)
        _lib.call("slu_colsum_acc", dgi1.data_ptr(), G, R, G, db_ih1.data_ptr(), st)
        return (dkeys, dvalues, dgi0, dinit, dwcat1[G:], dbcat1[G:], dw_c, dw_hh0, db_hh0, dw_ih1, db_ih1, dwcat1[:G], dbcat1[:G],
                None, None, None)


def teacher_forced_log_likelihood(dec, encoder_outputs, y, training):
    """Seq2SeqDecoder.forward on CUDA: log p(y | x) per example, [B].  `dec` holds the parameters (seq2seq.Seq2SeqDecoder).
    Per-example VALUES are exact; the gradient is the one of their mean (what Model.forward / the Trainer differentiate:
    models.py:827 `-log_probs.mean()`), because the fused output head yields the summed NLL, not per-row graph nodes."""
    if dec.rnn.num_layers != 2:
        raise NotImplementedError("slu_b200 CUDA path: the seq2seq decoder kernels take num_intent_decoder_layers = 2")
    B, U, S = y.shape
    att, cell0, cell1 = dec.attention, dec.rnn.layers[0], dec.rnn.layers[2]
    D = cell0.hidden_size
    y = y.float()
    y_prev = torch.zeros(B, U, S, device=y.device)
    y_prev[:, 0, dec.SOS] = 1.0
    y_prev[:, 1:] = y[:, :-1]
    pad = (-S) % 4
    y_prev = torch.nn.functional.pad(y_prev.transpose(0, 1), (0, pad)).contiguous()            # [U, B, S'] symbol-major
    keys = ops.LinearNT.apply(encoder_outputs, att.key_linear.weight, att.key_linear.bias)
    values = ops.LinearNT.apply(encoder_outputs, att.value_linear.weight, att.value_linear.bias)
    emb = ops.LinearNT.apply(y_prev, torch.nn.functional.pad(dec.embed.weight, (0, pad)), dec.embed.bias)      # [U, B, D]
    ge_all = ops.LinearNT.apply(emb, cell0.weight_ih[:, :D], cell0.bias_ih)                    # embedding half of cell 0's input
    p = dec.rnn.layers[1].p if training else 0.0
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0.0 else 0
    from . import engine
    seed_word = engine.graph_seed_word() if p > 0.0 else None          # set while the step is being captured as a CUDA graph
    states = DecoderStates.apply(keys, values, ge_all, dec.initial_state, att.query_linear.weight, att.query_linear.bias,
                                 cell0.weight_ih[:, D:], cell0.weight_hh, cell0.bias_hh, cell1.weight_ih, cell1.bias_ih,
                                 cell1.weight_hh, cell1.bias_hh, p, seed, seed_word)
    targets = y.argmax(-1).transpose(0, 1).contiguous().view(-1)                               # [U*B], symbol-major like `states`
    mean_nll, _, row_nll = ops.LinearCE.apply(states.reshape(U * B, D), dec.linear.weight, dec.linear.bias, targets)
    total = -float(U) * mean_nll                                                               # = mean_b log p(y_b | x_b)
    return -row_nll.view(U, B).sum(0) + (total - total.detach())


# ---- one decoding step for beam search (Seq2SeqDecoder.infer, reference models.py:558-651): forward only ----------------------
class StepCache:
    """What does not change over the symbols of one `infer` call: keys / values of the encoder states and the weights in the
    operand forms of the step projections (the reference re-projects the encoder states in every step of every hypothesis)."""

    def __init__(self, dec, encoder_outputs):
        att, cell0, cell1 = dec.attention, dec.rnn.layers[0], dec.rnn.layers[2]
        B = encoder_outputs.shape[0]
        self.keys = ops.LinearNT.apply(encoder_outputs, att.key_linear.weight, att.key_linear.bias).contiguous()
        self.values = ops.LinearNT.apply(encoder_outputs, att.value_linear.weight, att.value_linear.bias).contiguous()
        S = dec.embed.weight.shape[1]
        self.pad = (-S) % 4
        c = lambda t: t.detach().contiguous()
        self.w_embed = _W(c(torch.nn.functional.pad(dec.embed.weight.detach(), (0, self.pad))), False, B)
        self.w_q, self.w_ih0, self.w_hh0 = _W(c(att.query_linear.weight), False, B), _W(c(cell0.weight_ih), False, B), _W(c(cell0.weight_hh), False, B)
        self.w_ih1, self.w_hh1, self.w_out = _W(c(cell1.weight_ih), False, B), _W(c(cell1.weight_hh), False, B), _W(c(dec.linear.weight), False, B)
        self.b = [c(t) for t in (dec.embed.bias, att.query_linear.bias, cell0.bias_ih, cell0.bias_hh, cell1.bias_ih, cell1.bias_hh, dec.linear.bias)]


def beam_step(dec, cache, y_prev, state):
    """Seq2SeqDecoder._step on the library's kernels: attention over the cached keys / values, the two GRUCells, the output
    projection; returns (new state [B, 2, D], log-probabilities [B, |S|]).  No dropout (beam search runs in eval mode)."""
    if dec.rnn.num_layers != 2:
        raise NotImplementedError("slu_b200 CUDA path: the seq2seq decoder kernels take num_intent_decoder_layers = 2")
    B = y_prev.shape[0]
    D = dec.rnn.layers[0].hidden_size
    K, V, T = cache.keys.shape[2], cache.values.shape[2], cache.keys.shape[1]
    dev = y_prev.device
    st = _lib.stream()
    f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    b_emb, b_q, b_ih0, b_hh0, b_ih1, b_hh1, b_out = cache.b
    s0, s1 = state[:, 0].contiguous(), state[:, 1].contiguous()
    yp = torch.nn.functional.pad(y_prev.float(), (0, cache.pad)).contiguous()
    x0 = f(B, D + V)                                   # cell-0 input: embedding | context (models.py:540 torch.cat)
    emb, q, ctx, w = f(B, D), f(B, K), f(B, V), f(B, T)
    cache.w_embed.nt(yp, B, emb, b_emb)
    cache.w_q.nt(s1, B, q, b_q)
    _lib.call("slu_attn_step_fwd", q.data_ptr(), K, cache.keys.data_ptr(), cache.values.data_ptr(), B, T, K, V, 1.0 / math.sqrt(float(K)),
              w.data_ptr(), ctx.data_ptr(), st)
    x0[:, :D].copy_(emb); x0[:, D:].copy_(ctx)
    gi, gh, n0, n1 = f(B, 3 * D), f(B, 3 * D), f(B, D), f(B, D)
    cache.w_ih0.nt(x0, B, gi, b_ih0)
    cache.w_hh0.nt(s0, B, gh, b_hh0)
    _lib.call("slu_grucell_fwd", gi.data_ptr(), 3 * D, None, 0, gh.data_ptr(), 3 * D, s0.data_ptr(), None, B, D, 0.0, 0, None, 0,
              n0.data_ptr(), None, None, st)
    cache.w_ih1.nt(n0, B, gi, b_ih1)
    cache.w_hh1.nt(s1, B, gh, b_hh1)
    _lib.call("slu_grucell_fwd", gi.data_ptr(), 3 * D, None, 0, gh.data_ptr(), 3 * D, s1.data_ptr(), None, B, D, 0.0, 0, None, 0,
              n1.data_ptr(), None, None, st)
    S = dec.linear.weight.shape[0]
    logits = f(B, S)
    cache.w_out.nt(n1, B, logits, b_out)
    return torch.stack([n0, n1], dim=1), torch.log_softmax(logits, dim=1)
