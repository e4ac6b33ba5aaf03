"""Host mirror of MT/network.py: ``MusicTransformer`` with the reference constructor, the
``forward(x) -> logits`` contract, the reference state_dict layout and ``generate``.

forward (train / eval) = MT/network.py:35-40 on the CUDA kernels.  ``generate`` replaces the
reference's O(len^2)-per-step full recompute (:52-62) by a KV-cached single-token step (K7) and
its OneHotCategorical draw (:73-74) by the fused temperature / top-k sampler (K8); the literal
reference loop (no mask, sliding window) is kept as ``generate_literal``.
"""
from __future__ import annotations

import math
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
            decoder, w = self.Decoder(x, mask=look_ahead_mask)
            fc = _LinearFunction.apply(self.Decoder.cfg(), decoder, self.fc.weight, self.fc.bias)
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
                 return_logits: bool = False):
        """prior [B, P] int -> [B, P+length] int64.  Causal KV-cached decode (the mask the
        reference builds at MT/network.py:55-56 and drops IS applied); needs P+length-1 <=
        max_seq (no sliding window -- see generate_literal for the reference's literal loop).
        ``uniforms`` [length, B] fixes the random draws (tests); default torch.rand."""
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
        ids = torch.empty((B, P + length), dtype=torch.int32, device=prior.device)
        ids[:, :P] = prior.to(torch.int32)
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
        self.kc = [torch.zeros(shape, dtype=cfg.act, device=dev) for _ in self.Ws]
        self.vc = [torch.zeros(shape, dtype=cfg.act, device=dev) for _ in self.Ws]
        # pad bit per cached position: a generated/prior pad token is masked as a key, exactly as
        # the look-ahead mask of MT/utils.py:73 does in the reference's recompute
        self.pad_bits = torch.zeros((B, cfg.max_seq), dtype=torch.uint8, device=dev)

    def step(self, tok: torch.Tensor, t: int) -> torch.Tensor:
        """tok int32 [B] at position t -> logits fp32 [B, V] for position t+1."""
        cfg, B = self.cfg, self.B
        d, h, dh = cfg.d, cfg.h, cfg.dh
        lp = cfg.act != torch.float32
        dev = self.emb.device
        ids = tok.reshape(B, 1).contiguous()
        self.pad_bits[:, t] = (tok == config.pad_token)
        x = torch.empty((B, d), dtype=torch.float32, device=dev)
        x_lp = torch.empty((B, d), dtype=cfg.act, device=dev) if lp else None
        ops.embed_pos_fwd(ids, self.emb, self.pe, x, x_lp, t, math.sqrt(d), 0.0, 0, 0)
        xl = x_lp if lp else x
        for li, W in enumerate(self.Ws):
            qkv = torch.empty((B, 3 * d), dtype=cfg.act, device=dev)
            engine.linear_fwd(xl, W.Wqkv, W.bqkv, qkv, cfg)
            ops.kv_append(qkv, self.kc[li], self.vc[li], B, h, dh, cfg.max_seq, t)
            o = torch.empty((B, d), dtype=cfg.act, device=dev)
            ops.rga_decode(qkv, 3 * d, self.kc[li], self.vc[li], W.E, self.pad_bits, o, B, h, dh, cfg.max_seq, t)
            a = torch.empty((B, d), dtype=torch.float32, device=dev)
            engine.linear_fwd(o, W.Wfc, W.bfc, a, cfg)
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
