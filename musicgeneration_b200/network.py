"""Host mirror of MT/network.py: ``MusicTransformer`` with the reference constructor, the
``forward(x) -> logits`` contract, the reference state_dict layout and ``generate``.

forward (train / eval) = MT/network.py:35-40 on the CUDA kernels.  ``generate`` replaces the
reference's O(len^2)-per-step full recompute (:52-62) by a KV-cached single-token step (K7) and
its OneHotCategorical draw (:73-74) by the fused temperature / top-k sampler (K8); the literal
reference loop (no mask, sliding window) is kept as ``generate_literal``.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional

import torch

from . import _lib as L
from . import config, engine, ops, utils
from .engine import Mask
from .layers import Encoder, _LinearFunction, _PrecisionMixin, _PRECISIONS, default_precision


class MusicTransformer(torch.nn.Module, _PrecisionMixin):
    def __init__(self, embedding_dim=256, vocab_size=388 + 2, num_layer=6, max_seq=2048, dropout=0.2,
                 debug=False, loader_path=None, dist=False, writer=None, precision=None):
        super().__init__()
        self.infer = False
        if loader_path is not None:
            self.load_config_file(loader_path)     # undefined upstream as well -> AttributeError
        else:
            self._debug = debug
            self.max_seq = max_seq
            self.num_layer = num_layer
            self.embedding_dim = embedding_dim
            self.vocab_size = vocab_size
            self.dist = dist
        self.writer = writer
        self.Decoder = Encoder(num_layers=self.num_layer, d_model=self.embedding_dim,
                               input_vocab_size=self.vocab_size, rate=dropout, max_len=max_seq)
        self.fc = torch.nn.Linear(self.embedding_dim, self.vocab_size)
        self.precision = default_precision()
        if precision is not None:
            self.set_precision(precision)
        # sampling knobs of generate() (the reference parses -T and never uses it)
        self.temperature = 1.0
        self.top_k = 0
        self.greedy = False

    # ------------------------------------------------------------------------------------
    def forward(self, x, length=None, writer=None):
        if self.training or not self.infer:
            _, _, look_ahead_mask = utils.get_masked_with_pad_tensor(self.max_seq, x, x, config.pad_token)
            # parameters living in a FlatAdam buffer: ONE cast of the flat buffer gives every layer its
            # 16-bit weight operand (instead of a cast launch per matrix)
            from . import optim as _optim
            opt = getattr(self.fc.weight, "_mt_opt", None)
            act = self.Decoder.cfg().act
            if opt is not None and act != torch.float32 and x.is_cuda:
                opt.refresh_lp(act)
                _optim._ACTIVE_SHADOW[0] = opt
            try:
                decoder, w = self.Decoder(x, mask=look_ahead_mask)
                fc = _LinearFunction.apply(self.Decoder.cfg(), decoder, self.fc.weight, self.fc.bias)
            finally:
                _optim._ACTIVE_SHADOW[0] = None
            return fc.contiguous() if self.training else (fc.contiguous(), [weight.contiguous() for weight in w])
        else:
            return self.generate(x, length, None).contiguous().tolist()

    def test(self):
        self.eval()
        self.infer = True

    # ------------------------------------------------------------------------------------
    # KV-cached sampling
    # ------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, prior: torch.Tensor, length=2048, tf_board_writer=None,
                 temperature: Optional[float] = None, top_k: Optional[int] = None,
                 greedy: Optional[bool] = None, uniforms: Optional[torch.Tensor] = None,
                 return_logits: bool = False, graph: Optional[bool] = None):
        """prior [B, P] int -> [B, P+length] int64.  Causal KV-cached decode (the mask the
        reference builds at MT/network.py:55-56 and drops IS applied); needs P+length-1 <=
        max_seq (no sliding window -- see generate_literal for the reference's literal loop).
        ``uniforms`` [length, B] fixes the random draws (tests); default torch.rand.
        ``return_logits``: also the last-position logits of every generated event [length, B, V].
        ``graph``: None (default) = the persistent decode kernel (one launch for the whole generation) where it takes
        the model and pays (16-bit mode, B <= 32: _DecodeSession.persistent_default), else one captured CUDA graph
        replayed per event; True = the graph path; False = every step launched from the host."""
        if not prior.is_cuda:
            raise RuntimeError("musicgeneration_b200 runs on CUDA tensors only (no CPU fallback)")
        temperature = self.temperature if temperature is None else temperature
        top_k = self.top_k if top_k is None else top_k
        greedy = self.greedy if greedy is None else greedy
        B, P = prior.shape
        if P + length - 1 > self.max_seq:
            raise RuntimeError(f"prior ({P}) + length ({length}) - 1 exceeds max_seq ({self.max_seq}); "
                               "KV-cached decode does not slide the window")
        dec = _DecodeSession(self, B)
        ids = torch.zeros((B, P + length), dtype=torch.int32, device=prior.device)
        ids[:, :P] = prior.to(torch.int32)
        persistent = graph is None and dec.persistent_default()
        if graph is None:
            graph = persistent or not return_logits
        if graph and (persistent or P + length - 1 >= 4):
            # one CUDA graph of a whole decode step (device-resident step index), replayed per event
            if greedy:
                u = None
            elif uniforms is not None:
                u = uniforms[:length].to(torch.float32).contiguous()
            else:
                u = torch.rand((length, B), dtype=torch.float32, device=prior.device)
            rec = torch.empty((P + length - 1, B, self.vocab_size), dtype=torch.float32, device=prior.device) \
                if return_logits else None
            run = dec.run_persistent if persistent else dec.run_graph
            run(ids, P, P + length - 1, u, float(temperature), int(top_k), bool(greedy), logits_out=rec)
            return (ids.to(torch.int64), rec[P - 1:]) if return_logits else ids.to(torch.int64)
        step_logits = []
        logits = None
        for t in range(P + length - 1):
            logits = dec.step(ids[:, t], t)
            if t >= P - 1:
                u = None
                if not greedy:
                    u = uniforms[t - (P - 1)].contiguous() if uniforms is not None else \
                        torch.rand(B, dtype=torch.float32, device=prior.device)
                nxt = torch.empty((B,), dtype=torch.int32, device=prior.device)
                ops.sample(logits, u, nxt, float(temperature), int(top_k), bool(greedy))
                ids[:, t + 1] = nxt
                if return_logits:
                    step_logits.append(logits.clone())
        out = ids.to(torch.int64)
        return (out, torch.stack(step_logits)) if return_logits else out

    @torch.no_grad()
    def decode_logits(self, ids: torch.Tensor, graph: Optional[bool] = None) -> torch.Tensor:
        """Teacher-forced KV-cached pass over given ids [B, n] on the decode path (one graph replay per
        position): logits [n, B, V], entry t = the distribution of token t+1 given ids[:, :t+1] under the
        causal mask.  The same launches ``generate`` replays; used to score sequences and by the parity tests."""
        if not ids.is_cuda:
            raise RuntimeError("musicgeneration_b200 runs on CUDA tensors only (no CPU fallback)")
        B, n = ids.shape
        if n > self.max_seq:
            raise RuntimeError(f"sequence length {n} exceeds max_seq {self.max_seq}")
        dec = _DecodeSession(self, B)
        buf = torch.zeros((B, n + 1), dtype=torch.int32, device=ids.device)
        buf[:, :n] = ids.to(torch.int32)
        rec = torch.empty((n, B, self.vocab_size), dtype=torch.float32, device=ids.device)
        # prior_len n+1: nothing is sampled
        if graph is None and dec.persistent_default():
            dec.run_persistent(buf, n + 1, n, None, 1.0, 0, True, logits_out=rec)
        else:
            dec.run_graph(buf, n + 1, n, None, 1.0, 0, True, logits_out=rec)
        return rec

    @torch.no_grad()
    def generate_literal(self, prior: torch.Tensor, length=2048, greedy=True,
                         uniforms: Optional[torch.Tensor] = None):
        """The reference loop as written (MT/network.py:52-77): NO mask, full-stack recompute of
        the whole window every step, window slides at ``config.threshold_len``."""
        decode_array = prior
        result_array = prior
        was_training = self.training
        self.eval()
        try:
            for i in range(length):
                if decode_array.size(1) >= config.threshold_len:
                    decode_array = decode_array[:, 1:]
                hid, _ = _no_weights(self.Decoder, decode_array.contiguous())
                z = _LinearFunction.apply(self.Decoder.cfg(), hid[:, -1:, :].contiguous(), self.fc.weight,
                                          self.fc.bias)[:, 0].contiguous()
                nxt = torch.empty((z.shape[0],), dtype=torch.int32, device=z.device)
                u = None
                if not greedy:
                    u = uniforms[i].contiguous() if uniforms is not None else \
                        torch.rand(z.shape[0], dtype=torch.float32, device=z.device)
                ops.sample(z, u, nxt, 1.0, 0, bool(greedy))
                nxt = nxt.to(decode_array.dtype).unsqueeze(-1)
                decode_array = torch.cat((decode_array, nxt), dim=-1)
                result_array = torch.cat((result_array, nxt), dim=-1)
        finally:
            self.train(was_training)
        return result_array


def _no_weights(encoder: Encoder, ids: torch.Tensor):
    """Encoder forward with mask=None and without materialising the L x L weights."""
    from .layers import _EncoderFunction
    outs = _EncoderFunction.apply(encoder, None, False, ids, *encoder.params())
    return outs[0], None


class _DecodeSession:
    """Per-generation state: K/V caches [layers][B,h,max_seq,dh] and the per-layer operands."""

    def __init__(self, model: MusicTransformer, B: int):
        enc = model.Decoder
        self.cfg = enc.cfg()
        self.cfg.p_drop = 0.0
        cfg = self.cfg
        dev = enc.embedding.weight.device
        self.B = B
        self.enc = enc
        self.Ws = [l.weights(cfg.act) for l in enc.enc_layers]
        self.pe = enc.pos_encoding.table(dev)
        self.emb = enc.embedding.weight.data
        from .layers import _act_copy
        self.Wv = _act_copy(model.fc.weight.data, cfg.act)
        self.bv = model.fc.bias.data
        self.V = model.fc.weight.shape[0]
        shape = (B, cfg.h, cfg.max_seq, cfg.dh)
        # per-layer 16-bit type: the first layer's attention block runs on f16 operands in the bf16 mode
        # (engine.py docstring) -- here x, Wqkv, q/k/v, the KV cache, E, the attention output and Wfc
        self.lact = [cfg.act] * len(self.Ws)
        if self.Ws and self.Ws[0].Wqkv_hp is not None and cfg.dh == 64 and cfg.gemm_path != L.PATH_SIMT:
            self.lact[0] = torch.float16
            self.Ws[0].Wfc_hp = _act_copy(enc.enc_layers[0].rga.fc.weight.data, torch.float16)
        self.kc = [torch.zeros(shape, dtype=a, device=dev) for a in self.lact]
        self.vc = [torch.zeros(shape, dtype=a, device=dev) for a in self.lact]
        # pad bit per cached position: a generated/prior pad token is masked as a key, exactly as
        # the look-ahead mask of MT/utils.py:73 does in the reference's recompute
        self.pad_bits = torch.zeros((B, cfg.max_seq), dtype=torch.uint8, device=dev)

    # ---- persistent mode: the whole generation is one launch (csrc/decode_step.cu) ----------------
    def persistent_ok(self) -> bool:
        cfg = self.cfg
        return (cfg.act == torch.bfloat16 and cfg.dh == 64 and cfg.gemm_path != L.PATH_SIMT
                and cfg.attn_path != L.PATH_SIMT
                and ops.decode_run_supported(self.B, cfg.d, cfg.h, self.V, len(self.Ws)))

    def persistent_default(self) -> bool:
        """Whether generate() takes the persistent kernel by itself.  Measured on B200 at config B (us per event,
        persistent / graph path): 8 sequences 243 / 345, 32 sequences 408 / 472 -- the one launch wins most where the
        step is launch- and fill-latency (few sequences); it is the default up to 32 sequences (what was measured).  MT_DECODE_PERSISTENT=1 / 0 forces it
        on (wherever it is supported) / off."""
        env = os.environ.get("MT_DECODE_PERSISTENT")
        if env is not None:
            return env != "0" and self.persistent_ok()
        return self.B <= 32 and self.persistent_ok()

    def run_persistent(self, ids, prior_len, n_steps, u, temperature, top_k, greedy, logits_out=None):
        """Positions 0 .. n_steps-1 of ``ids`` [B, >= n_steps+1] in ONE launch (same contract as run_graph)."""
        cfg = self.cfg
        rows, f16, keep = [], [], []          # keep: the operand tensors stay referenced until the launch has been enqueued
        for li, W in enumerate(self.Ws):
            hp = self.lact[li] != cfg.act
            ts = [W.Wqkv_hp if hp else W.Wqkv, W.bqkv, W.Wfc_hp if hp else W.Wfc, W.bfc, W.Wpre, W.bpre, W.Wsuf, W.bsuf,
                  W.g1, W.b1, W.g2, W.b2, W.E_hp if hp else W.E, self.kc[li], self.vc[li]]
            for x in ts:
                if not x.is_contiguous():
                    raise RuntimeError("decode_run: non-contiguous layer operand")
            rows.append([x.data_ptr() for x in ts])
            f16.append(1 if hp else 0)
            keep.extend(ts)
        ptrs = torch.tensor(rows, dtype=torch.int64)
        flags = torch.tensor(f16, dtype=torch.int32)
        ws = torch.empty(L.load().mt_decode_run_workspace_bytes(self.B, cfg.d, self.V), dtype=torch.uint8,
                         device=self.emb.device)
        ops.decode_run(ids, 0, n_steps, prior_len, self.emb, self.pe, ptrs, flags, self.Wv, self.bv, cfg.d, cfg.h,
                       cfg.max_seq, config.pad_token, self.pad_bits, u, temperature, top_k, greedy, logits_out, ws)
        self._keep = (keep, ws)               # (stream-ordered use: released with the session)

    # ---- graph mode: every buffer preallocated, the position read from device memory ----------
    def _alloc_step_buffers(self):
        cfg, B = self.cfg, self.B
        d = cfg.d
        dev = self.emb.device
        lp = cfg.act != torch.float32
        f32 = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
        act = lambda *sh: torch.empty(sh, dtype=cfg.act, device=dev)
        nl = len(self.Ws)
        self.g_x = [f32(B, d) for _ in range(nl + 1)]
        self.g_xlp = [act(B, d) if lp else None for _ in range(nl + 1)]
        if lp and nl:
            self.g_xlp[0] = torch.empty((B, d), dtype=self.lact[0], device=dev)
        self.g_qkv = {a: torch.empty((B, 3 * d), dtype=a, device=dev) for a in set(self.lact)}
        self.g_o = {a: torch.empty((B, d), dtype=a, device=dev) for a in set(self.lact)}
        self.g_a = f32(B, d)
        self.g_out1, self.g_out1lp = f32(B, d), (act(B, d) if lp else None)
        self.g_hmid, self.g_f = act(B, d // 2), f32(B, d)
        self.g_mean, self.g_rstd = f32(B), f32(B)
        self.g_logits = f32(B, self.V)
        self.t_dev = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.g_ws = ops.decode_attend_workspace(B, cfg.h, cfg.dh, cfg.max_seq, dev)

    def _graph_step(self, ids, prior_len, u, temperature, top_k, greedy):
        """Enqueue one decode step at position *t_dev (embed -> layers -> vocabulary GEMM -> sample ->
        advance); the same launches serve every position, so they are captured once."""
        cfg, B = self.cfg, self.B
        d, h, dh = cfg.d, cfg.h, cfg.dh
        lp = cfg.act != torch.float32
        ops.decode_embed(ids, self.t_dev, self.emb, self.pe, self.g_x[0], self.g_xlp[0], math.sqrt(d),
                         config.pad_token, self.pad_bits)
        for li, W in enumerate(self.Ws):
            x, xl = self.g_x[li], (self.g_xlp[li] if lp else self.g_x[li])
            hp = self.lact[li] != cfg.act
            g_qkv, g_o = self.g_qkv[self.lact[li]], self.g_o[self.lact[li]]
            engine.linear_fwd(xl, W.Wqkv_hp if hp else W.Wqkv, W.bqkv, g_qkv, cfg)
            ops.decode_attend(g_qkv, 3 * d, self.kc[li], self.vc[li], W.E_hp if hp else W.E, self.pad_bits, g_o,
                              self.t_dev, B, h, dh, cfg.max_seq, self.g_ws, append=True)
            engine.linear_fwd(g_o, W.Wfc_hp if hp else W.Wfc, W.bfc, self.g_a, cfg)
            ops.add_ln_fwd(self.g_a, x, W.g1, W.b1, self.g_out1, self.g_out1lp, self.g_mean, self.g_rstd,
                           1e-6, 0.0, 0, 0)
            o1 = self.g_out1lp if lp else self.g_out1
            engine.linear_fwd(o1, W.Wpre, W.bpre, self.g_hmid, cfg, relu=True)
            engine.linear_fwd(self.g_hmid, W.Wsuf, W.bsuf, self.g_f, cfg)
            ops.add_ln_fwd(self.g_f, self.g_out1, W.g2, W.b2, self.g_x[li + 1], self.g_xlp[li + 1],
                           self.g_mean, self.g_rstd, 1e-6, 0.0, 0, 0)
        nl = len(self.Ws)
        engine.linear_fwd(self.g_xlp[nl] if lp else self.g_x[nl], self.Wv, self.bv, self.g_logits, cfg)
        ops.decode_sample(self.g_logits, u, ids, self.t_dev, prior_len, temperature, top_k, greedy)
        ops.decode_advance(self.t_dev)

    def run_graph(self, ids, prior_len, n_steps, u, temperature, top_k, greedy, logits_out=None):
        """Positions 0 .. n_steps-1 of ``ids`` [B, >= n_steps+1] (prior tokens kept, later ones sampled).
        ``logits_out`` [n_steps, B, V]: receives the step's logits after every replay."""
        self._alloc_step_buffers()
        args = (ids, prior_len, u, temperature, top_k, greedy)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._graph_step(*args)          # warm-up outside capture: function attributes, workspaces
            self.t_dev.zero_()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        chain = os.environ.get("MT_DECODE_CHAIN", "1") != "0"
        with torch.cuda.graph(graph, stream=side):
            ops.decode_chain(chain)          # kernels of the captured step launch programmatically dependent
            try:
                self._graph_step(*args)
            finally:
                ops.decode_chain(False)
        self.t_dev.zero_()
        for t in range(n_steps):
            graph.replay()
            if logits_out is not None:
                logits_out[t].copy_(self.g_logits)

    def step(self, tok: torch.Tensor, t: int) -> torch.Tensor:
        """tok int32 [B] at position t -> logits fp32 [B, V] for position t+1."""
        cfg, B = self.cfg, self.B
        d, h, dh = cfg.d, cfg.h, cfg.dh
        lp = cfg.act != torch.float32
        dev = self.emb.device
        ids = tok.reshape(B, 1).contiguous()
        self.pad_bits[:, t] = (tok == config.pad_token)
        x = torch.empty((B, d), dtype=torch.float32, device=dev)
        x_lp = torch.empty((B, d), dtype=self.lact[0] if self.lact else cfg.act, device=dev) if lp else None
        ops.embed_pos_fwd(ids, self.emb, self.pe, x, x_lp, t, math.sqrt(d), 0.0, 0, 0)
        xl = x_lp if lp else x
        for li, W in enumerate(self.Ws):
            la = self.lact[li]
            hp = la != cfg.act
            qkv = torch.empty((B, 3 * d), dtype=la, device=dev)
            engine.linear_fwd(xl, W.Wqkv_hp if hp else W.Wqkv, W.bqkv, qkv, cfg)
            ops.kv_append(qkv, self.kc[li], self.vc[li], B, h, dh, cfg.max_seq, t)
            o = torch.empty((B, d), dtype=la, device=dev)
            ops.rga_decode(qkv, 3 * d, self.kc[li], self.vc[li], W.E_hp if hp else W.E, self.pad_bits, o, B, h, dh,
                           cfg.max_seq, t)
            a = torch.empty((B, d), dtype=torch.float32, device=dev)
            engine.linear_fwd(o, W.Wfc_hp if hp else W.Wfc, W.bfc, a, cfg)
            out1 = torch.empty((B, d), dtype=torch.float32, device=dev)
            out1_lp = torch.empty((B, d), dtype=cfg.act, device=dev) if lp else None
            mean = torch.empty((B,), dtype=torch.float32, device=dev)
            rstd = torch.empty((B,), dtype=torch.float32, device=dev)
            ops.add_ln_fwd(a, x, W.g1, W.b1, out1, out1_lp, mean, rstd, 1e-6, 0.0, 0, 0)
            o1 = out1_lp if lp else out1
            hmid = torch.empty((B, d // 2), dtype=cfg.act, device=dev)
            engine.linear_fwd(o1, W.Wpre, W.bpre, hmid, cfg, relu=True)
            f = torch.empty((B, d), dtype=torch.float32, device=dev)
            engine.linear_fwd(hmid, W.Wsuf, W.bsuf, f, cfg)
            x = torch.empty((B, d), dtype=torch.float32, device=dev)
            x_lp = torch.empty((B, d), dtype=cfg.act, device=dev) if lp else None
            ops.add_ln_fwd(f, out1, W.g2, W.b2, x, x_lp, mean, rstd, 1e-6, 0.0, 0, 0)
            xl = x_lp if lp else x
        logits = torch.empty((B, self.V), dtype=torch.float32, device=dev)
        engine.linear_fwd(xl, self.Wv, self.bv, logits, cfg)
        return logits
