"""Sequence-to-sequence intent module (character-level attention decoder) of the drop-in `models.py`.

Reference: models.py:381-411 (Seq2SeqEncoder), 413-436 (Attention), 438-484 (DecoderRNN), 486-651 (sort_beam,
Seq2SeqDecoder.forward / infer).  SURVEY.md section 8 keeps the decoder OUT of the hot path ("stays stock PyTorch on top of
the encoder kernels"): on a CUDA device the encoder's bidirectional GRU runs on the sm_100a persistent-GRU kernel and the teacher-forced
decoder (`Seq2SeqDecoder.forward`, the training path) on the library's decoder kernels (end-to-end-slu_b200/decoder.py:
batched tcgen05 GEMMs outside the recurrence, fused attention / GRUCell step kernels inside it, hand-written backward
through time); beam search (`infer`) runs its per-symbol step on the same kernels (decoder.beam_step: keys / values projected
once per call, attention, both GRUCells and the output projection) and keeps only the hypothesis bookkeeping (top-k, sort,
gather) in torch, written batched (gather/scatter instead of the reference's per-element python loops) with identical
results.  The CPU path is plain torch ops.  Parameter names match the reference so its seq2seq
checkpoints (`encoder.layers.0.*`, `decoder.{initial_state, embed, attention.*, rnn.layers.{0,2,..}, linear}`) load strictly.
"""
import torch
import torch.nn.functional as F


class RNNSelect(torch.nn.Module):
    def forward(self, input):
        return input[0]


class Seq2SeqEncoder(torch.nn.Module):
    """Stack of bidirectional GRUs + Dropout(0.5) over the word-module features (models.py:381-411)."""

    def __init__(self, input_dim, num_layers, encoder_dim):
        super().__init__()
        layers, dim = [], input_dim
        for idx in range(num_layers):
            gru = torch.nn.GRU(input_size=dim, hidden_size=encoder_dim, batch_first=True, bidirectional=True)
            gru.name = "intent_encoder_rnn%d" % idx
            sel = RNNSelect(); sel.name = "intent_encoder_rnn_select%d" % idx
            drop = torch.nn.Dropout(p=0.5); drop.name = "intent_encoder_dropout%d" % idx
            layers += [gru, sel, drop]
            dim = 2 * encoder_dim
        self.layers = torch.nn.ModuleList(layers)

    def forward(self, x):
        if x.is_cuda:
            import importlib
            eng = importlib.import_module("end-to-end-slu_b200").engine
            rnns = [(l, self.layers[i + 2].p, 1) for i, l in enumerate(self.layers) if isinstance(l, torch.nn.GRU)]
            if all(g.hidden_size == 128 for g, _, _ in rnns):
                return eng._run_rnns(x, rnns, self.training)
            raise NotImplementedError("slu_b200 CUDA path: intent_encoder_dim must be 128")
        for layer in self.layers:
            x = layer(x)
        return x


class Attention(torch.nn.Module):
    """Scaled dot-product attention of one decoder state over the encoder states (models.py:413-436)."""

    def __init__(self, encoder_dim, decoder_dim, key_dim, value_dim):
        super().__init__()
        self.scale_factor = torch.sqrt(torch.tensor(key_dim).float())
        self.key_linear = torch.nn.Linear(encoder_dim, key_dim)
        self.query_linear = torch.nn.Linear(decoder_dim, key_dim)
        self.value_linear = torch.nn.Linear(encoder_dim, value_dim)
        self.softmax = torch.nn.Softmax(dim=1)

    def forward(self, encoder_states, decoder_state):
        keys = self.key_linear(encoder_states)                       # (B, T, K)
        values = self.value_linear(encoder_states)                   # (B, T, V)
        query = self.query_linear(decoder_state).unsqueeze(2)        # (B, K, 1)
        scores = torch.matmul(keys, query) / self.scale_factor.to(keys.device)
        weights = self.softmax(scores).transpose(1, 2)               # (B, 1, T)
        return torch.matmul(weights, values).squeeze(1)


class DecoderRNN(torch.nn.Module):
    """GRUCell stack with Dropout after each cell (models.py:438-484)."""

    def __init__(self, num_decoder_layers, num_decoder_hidden, input_size, dropout):
        super().__init__()
        self.num_layers = num_decoder_layers
        layers = []
        for index in range(num_decoder_layers):
            cell = torch.nn.GRUCell(input_size=input_size if index == 0 else num_decoder_hidden, hidden_size=num_decoder_hidden)
            cell.name = "gru%d" % index
            drop = torch.nn.Dropout(p=dropout); drop.name = "dropout%d" % index
            layers += [cell, drop]
        self.layers = torch.nn.ModuleList(layers)

    def forward(self, input, previous_state):
        """input (B, input_size); previous_state (B, L, D) -> new state (B, L, D)."""
        out, states = input, []
        for index in range(self.num_layers):
            h = self.layers[2 * index](out, previous_state[:, index])
            states.append(h)
            out = self.layers[2 * index + 1](h)
        return torch.stack(states, dim=1)


class Seq2SeqDecoder(torch.nn.Module):
    """Attention decoder: teacher-forced log-likelihood (`forward`) and beam search (`infer`) (models.py:500-651)."""

    def __init__(self, num_labels, num_layers, encoder_dim, decoder_dim, key_dim, value_dim, SOS=0):
        super().__init__()
        embedding_dim = decoder_dim
        self.embed = torch.nn.Linear(num_labels, embedding_dim)
        self.attention = Attention(encoder_dim * 2, decoder_dim, key_dim, value_dim)
        self.rnn = DecoderRNN(num_layers, decoder_dim, embedding_dim + value_dim, dropout=0.5)
        self.initial_state = torch.nn.Parameter(torch.randn(num_layers, decoder_dim))
        self.linear = torch.nn.Linear(decoder_dim, num_labels)
        self.log_softmax = torch.nn.LogSoftmax(dim=1)
        self.SOS = SOS

    def _step(self, encoder_outputs, y_prev, state, cache=None):
        if cache is not None:             # CUDA: attention / GRUCells / projections on the library's kernels (decoder.beam_step)
            import importlib
            return importlib.import_module("end-to-end-slu_b200").decoder.beam_step(self, cache, y_prev, state)
        context = self.attention(encoder_outputs, state[:, -1])
        state = self.rnn(torch.cat([self.embed(y_prev), context], dim=1), state)
        return state, self.log_softmax(self.linear(state[:, -1]))

    def forward(self, encoder_outputs, y, y_lengths=None):
        """encoder_outputs (B, T, 2*encoder_dim); y (B, U, num_labels) one-hot.  Returns log p(y|x) per example."""
        if encoder_outputs.is_cuda:       # the library's decoder kernels (end-to-end-slu_b200/decoder.py); no torch fallback on CUDA
            import importlib
            dec = importlib.import_module("end-to-end-slu_b200").decoder
            return dec.teacher_forced_log_likelihood(self, encoder_outputs, y, self.training)
        B, U, S = y.shape
        state = self.initial_state.unsqueeze(0).expand(B, -1, -1)
        y_prev = torch.zeros(B, S, device=y.device)
        y_prev[:, self.SOS] = 1.
        log_p = 0
        for u in range(U):
            state, out = self._step(encoder_outputs, y_prev, state)
            log_p = log_p + (out * y[:, u, :]).sum(dim=1)
            y_prev = y[:, u, :]
        return log_p

    @torch.no_grad()
    def infer(self, encoder_outputs, Sy, B=4, debug=False, y_lengths=None):
        """Beam search (width B).  Returns (beam_scores (B, batch), beam (B, batch, U, |Sy|) one-hot); hypothesis 0 is the best.
        Same procedure as models.py:558-651 (first step feeds an all-zero previous symbol, no end-of-sequence stopping,
        U = 200 steps unless y_lengths is given), batched over hypotheses instead of per-element loops."""
        dev = encoder_outputs.device
        batch, S = encoder_outputs.shape[0], len(Sy)
        U = 200 if y_lengths is None else max(y_lengths)
        L, D = self.initial_state.shape
        cache = None
        if dev.type == "cuda":
            import importlib
            cache = importlib.import_module("end-to-end-slu_b200").decoder.StepCache(self, encoder_outputs)
        beam = torch.zeros(B, batch, U, S, device=dev)
        scores = torch.zeros(B, batch, device=dev)
        states = torch.zeros(B, batch, L, D, device=dev)
        ar = torch.arange(batch, device=dev)
        for u in range(U):
            n_hyp = 1 if u == 0 else B
            cand_scores, cand_ext, cand_ptr, new_states = [], [], [], []
            for b in range(n_hyp):
                if u == 0:
                    state = self.initial_state.unsqueeze(0).expand(batch, -1, -1)
                    y_prev = torch.zeros(batch, S, device=dev)
                else:
                    state, y_prev = states[b], beam[b, :, u - 1, :]
                state, out = self._step(encoder_outputs, y_prev, state, cache)
                new_states.append(state)
                top_s, top_i = out.topk(B)                                   # (batch, B)
                cand_scores.append(top_s.t() + scores[b])                    # (B, batch)
                cand_ext.append(top_i.t())
                cand_ptr.append(torch.full((B, batch), b, dtype=torch.long, device=dev))
            cand_scores = torch.cat(cand_scores); cand_ext = torch.cat(cand_ext); cand_ptr = torch.cat(cand_ptr)
            new_states = torch.stack(new_states)                             # (n_hyp, batch, L, D)
            order = cand_scores.sort(dim=0, descending=True)[1][:B]          # (B, batch)
            scores = cand_scores.gather(0, order)
            ptr = cand_ptr.gather(0, order)
            ext = cand_ext.gather(0, order)
            beam = beam[ptr, ar.unsqueeze(0)]                                # re-parent the hypotheses
            beam[:, :, u, :] = F.one_hot(ext, S).float()
            states = new_states[ptr, ar.unsqueeze(0)]
        return scores, beam
