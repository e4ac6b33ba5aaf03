"""Host mirror of MT/layers.py: same class names, constructor signatures, attribute names and
state_dict layout; every forward/backward runs on the hand-written sm_100a kernels through
libmt_b200 (see engine.py).  There is no CPU path.

Classes: DynamicPositionEmbedding (MT/layers.py:22-39), RelativeGlobalAttention (:42-133),
EncoderLayer (:136-161), Encoder (:207-233).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import engine, ops
from .engine import LayerWeights, Mask, StackCfg

_PRECISIONS = {"fp32": torch.float32, "bf16": torch.bfloat16}
_seed_counter = [0]


def default_precision() -> str:
    return os.environ.get("MT_B200_PRECISION", "fp32")


def _next_seed() -> int:
    """Dropout seed for one forward call: torch's seed mixed with a call counter and the rank."""
    _seed_counter[0] += 1
    rank = int(os.environ.get("RANK", "0"))
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0x100000001B3 + rank * 7919) \
        & 0xFFFFFFFFFFFFFFFF


def sinusoid_table(max_seq: int, d: int) -> np.ndarray:
    """float64 [1, max_seq, d]; element (pos, i) = sin(pos * e^{-ln(1e4) i/d} * e^{ln(1e4)/d * (i%2)}
    + pi/2 * (i%2)) -- the table MT/layers.py:25-34 builds with scalar loops, vectorised."""
    pos = np.arange(max_seq, dtype=np.float64).reshape(-1, 1)
    idx = np.arange(d, dtype=np.float64).reshape(1, -1)
    odd = (np.arange(d) % 2).reshape(1, -1)
    ang = pos * np.exp(-math.log(10000) * idx / d) * np.exp(math.log(10000) / d * odd) + 0.5 * math.pi * odd
    return np.sin(ang)[None]


def as_mask(mask, B: int, Lq: int) -> Optional[Mask]:
    """Accepts None, an engine.Mask, or the reference's bool tensor [B,1,L,L] (True = masked).
    A tensor must have the look-ahead structure  pad_keys[b,j] | (j > i)  -- that is the only mask
    the reference ever builds (MT/utils.py:58-83); it is decomposed here (O(B L^2) check)."""
    if mask is None or isinstance(mask, Mask):
        return mask
    if not isinstance(mask, torch.Tensor):
        raise TypeError("mask must be None, engine.Mask or a bool tensor")
    m = mask.to(torch.bool)
    if m.dim() != 4 or m.size(-1) != Lq or m.size(-2) != Lq:
        raise RuntimeError(f"The size of mask {tuple(m.shape)} must match the sequence length {Lq}")
    m = m.expand(B, 1, Lq, Lq)
    pad = m[:, 0, -1, :]
    tri = torch.triu(torch.ones(Lq, Lq, dtype=torch.bool, device=m.device), 1)
    if torch.equal(m[:, 0], pad[:, None, :] | tri):
        return Mask(True, pad.to(torch.uint8).contiguous() if bool(pad.any()) else None)
    if torch.equal(m[:, 0], pad[:, None, :].expand(B, Lq, Lq)):
        return Mask(False, pad.to(torch.uint8).contiguous() if bool(pad.any()) else None)
    raise NotImplementedError("only look-ahead (causal | key-padding) masks are supported")


class DynamicPositionEmbedding(torch.nn.Module):
    def __init__(self, embedding_dim, max_seq=2048):
        super().__init__()
        self.positional_embedding = sinusoid_table(max_seq, embedding_dim)   # numpy, not a buffer
        self._dev_cache = {}

    def table(self, device) -> torch.Tensor:
        """fp32 [max_seq, d] resident on ``device`` (the reference re-uploads it every call)."""
        key = str(device)
        t = self._dev_cache.get(key)
        if t is None:
            t = torch.from_numpy(self.positional_embedding[0].astype(np.float32)).to(device).contiguous()
            self._dev_cache[key] = t
        return t

    def forward(self, x):
        return x + self.table(x.device)[None, :x.size(1), :].to(x.dtype)


class _PrecisionMixin:
    def set_precision(self, precision: str):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        for m in self.modules():
            if isinstance(m, _PrecisionMixin):
                m.precision = precision
        return self


def _pack_rows(params: Sequence[torch.nn.Parameter]) -> torch.Tensor:
    """Make ``params`` (same trailing shape) adjacent rows of ONE buffer and return that buffer;
    each ``param.data`` becomes a view, so state_dict keys / optimizers are unaffected."""
    first = params[0]
    tail = tuple(first.shape[1:])
    rows = sum(p.shape[0] for p in params)
    es = first.data.element_size()
    store = first.data.untyped_storage()
    total = sum(p.data.numel() for p in params)
    # adjacent means: views of ONE storage at consecutive offsets (address adjacency of separate
    # allocations does not count -- set_() past the end of a storage would reallocate it)
    ok = all(p.data.is_contiguous() and p.data.untyped_storage().data_ptr() == store.data_ptr()
             for p in params)
    ok = ok and (first.data.storage_offset() + total) * es <= store.nbytes()
    if ok:
        off = first.data.storage_offset()
        for p in params:
            if p.data.storage_offset() != off:
                ok = False
                break
            off += p.data.numel()
    if ok:
        return torch.empty(0, dtype=first.dtype, device=first.device).set_(
            store, first.data.storage_offset(), (rows,) + tail)
    flat = torch.empty((rows,) + tail, dtype=first.dtype, device=first.device)
    r = 0
    for p in params:
        flat[r:r + p.shape[0]].copy_(p.data)
        p.data = flat[r:r + p.shape[0]]
        r += p.shape[0]
    return flat


class RelativeGlobalAttention(torch.nn.Module, _PrecisionMixin):
    """Relative global attention of Music Transformer (Huang et al. 2018), MT/layers.py:42-133."""

    def __init__(self, h=4, d=256, add_emb=False, max_seq=2048, **kwargs):
        super().__init__()
        self.len_k = None
        self.max_seq = max_seq
        self.h = h
        self.d = d
        self.dh = d // h
        self.Wq = torch.nn.Linear(self.d, self.d)
        self.Wk = torch.nn.Linear(self.d, self.d)
        self.Wv = torch.nn.Linear(self.d, self.d)
        self.fc = torch.nn.Linear(d, d)
        self.additional = add_emb
        self.E = torch.nn.Parameter(torch.randn([self.max_seq, int(self.dh)]))
        if self.additional:
            self.Radd = None
        self.precision = default_precision()
        self.need_weights = None      # None: weights only outside training (see forward)
        self.gemm_path = L.PATH_AUTO
        self.attn_path = L.PATH_AUTO

    # -- operands for the kernels ---------------------------------------------------------
    def params(self) -> List[torch.nn.Parameter]:
        return [self.Wq.weight, self.Wq.bias, self.Wk.weight, self.Wk.bias, self.Wv.weight,
                self.Wv.bias, self.fc.weight, self.fc.bias, self.E]

    def cfg(self, p_drop=0.0) -> StackCfg:
        return StackCfg(d=self.d, h=self.h, max_seq=self.max_seq, p_drop=p_drop,
                        act=_PRECISIONS[self.precision], gemm_path=self.gemm_path,
                        attn_path=self.attn_path)

    def packed(self):
        wqkv = _pack_rows([self.Wq.weight, self.Wk.weight, self.Wv.weight])
        bqkv = _pack_rows([self.Wq.bias, self.Wk.bias, self.Wv.bias])
        return wqkv, bqkv

    def forward(self, inputs, mask=None, **kwargs):
        """inputs = [Q, K, V] ([B,L,d] each; the stack passes [x,x,x]); returns (out [B,L,d],
        attention weights [B,h,L,L]).  The L x L weights are materialised only when
        ``need_weights`` is True, or -- by default -- when the module is not training (the one
        place the reference consumes them is eval mode, MT/network.py:40)."""
        xq, xk, xv = inputs[0], inputs[1], inputs[2]
        if xk.shape != xq.shape or xv.shape != xq.shape:
            raise NotImplementedError("len_k != len_q is not on the MusicTransformer path")
        B, Lq, _ = xq.shape
        if Lq > self.max_seq:
            raise RuntimeError(f"sequence length {Lq} exceeds max_seq {self.max_seq}")
        self.len_k = self.len_q = Lq
        m = as_mask(mask, B, Lq)
        want_w = (not self.training) if self.need_weights is None else bool(self.need_weights)
        out, w = _RGAFunction.apply(self, m, want_w, xq, xk, xv, *self.params())
        return out, w


def _act_copy(x_f32_2d: torch.Tensor, act: torch.dtype) -> torch.Tensor:
    if act == torch.float32:
        return x_f32_2d
    y = torch.empty(x_f32_2d.shape, dtype=act, device=x_f32_2d.device)
    ops.cast(x_f32_2d, y)
    return y


def _weight_copy(w_f32: torch.Tensor, act: torch.dtype) -> torch.Tensor:
    """16-bit operand copy of a weight: a view of the optimizer's refreshed shadow when one is active
    (optim.weight_shadow), else a cast launch."""
    if act == torch.float32:
        return w_f32
    from .optim import weight_shadow
    sh = weight_shadow(w_f32, act)
    return sh if sh is not None else _act_copy(w_f32, act)


def _as_f32_2d(x: torch.Tensor, d: int) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("musicgeneration_b200 runs on CUDA tensors only (no CPU fallback)")
    x2 = x.reshape(-1, d)
    if x2.dtype != torch.float32:
        x2 = x2.float()
    return x2.contiguous()


def hp_first_layer() -> bool:
    """f16 operands for the first layer's attention in the bf16 mode (engine.py docstring); MT_B200_L0_F16=0
    turns it off (A/B measurements of the all-bf16 arithmetic)."""
    return os.environ.get("MT_B200_L0_F16", "1") != "0"


def _rga_weights_for(rga: RelativeGlobalAttention, act: torch.dtype, ffn=None, lns=None, hp=False) -> LayerWeights:
    """LayerWeights for one attention block (and, when given, the FFN / LayerNorm params).  ``hp``: also
    the f16 copies of Wq/Wk/Wv and E (first layer of a stack in the bf16 mode)."""
    wqkv, bqkv = rga.packed()
    hp = hp and act == torch.bfloat16 and hp_first_layer()

    def a(t):
        return _weight_copy(t.data, act)

    z = torch.empty(0, device=wqkv.device)
    return LayerWeights(
        Wqkv=_weight_copy(wqkv, act), bqkv=bqkv,
        Wfc=a(rga.fc.weight), bfc=rga.fc.bias.data,
        Wpre=a(ffn[0].weight) if ffn else z, bpre=ffn[0].bias.data if ffn else z,
        Wsuf=a(ffn[1].weight) if ffn else z, bsuf=ffn[1].bias.data if ffn else z,
        E=a(rga.E),
        g1=lns[0].weight.data if lns else z, b1=lns[0].bias.data if lns else z,
        g2=lns[1].weight.data if lns else z, b2=lns[1].bias.data if lns else z,
        Wqkv_hp=_act_copy(wqkv, torch.float16) if hp else None,
        E_hp=_act_copy(rga.E.data, torch.float16) if hp else None)


class _RGAFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rga, mask, want_w, xq, xk, xv, *params):
        cfg = rga.cfg()
        B, Lq, d = xq.shape
        same = all(t.data_ptr() == xq.data_ptr() and t.shape == xq.shape and t.stride() == xq.stride()
                   for t in (xk, xv))
        q2 = _act_copy(_as_f32_2d(xq, d), cfg.act)
        k2 = q2 if same else _act_copy(_as_f32_2d(xk, d), cfg.act)
        v2 = q2 if same else _act_copy(_as_f32_2d(xv, d), cfg.act)
        W = _rga_weights_for(rga, cfg.act)
        a, saved, P = engine.rga_block_fwd(q2, k2, v2, W, cfg, B, Lq, mask, want_w, keep_p=any(ctx.needs_input_grad))
        ctx.rga, ctx.cfg, ctx.W, ctx.saved = rga, cfg, W, saved
        ctx.shape = (B, Lq, d)
        if P is not None:
            ctx.mark_non_differentiable(P)
        if a.dtype != torch.float32:         # the module surface returns fp32 (MT/layers.py:108); the stack consumes `a` as is
            a32 = torch.empty((B * Lq, d), dtype=torch.float32, device=a.device)
            ops.cast(a, a32)
            a = a32
        return a.view(B, Lq, d), P

    @staticmethod
    def backward(ctx, d_out, _dP):
        cfg, W, s = ctx.cfg, ctx.W, ctx.saved
        B, Lq, d = ctx.shape
        d_a = _act_copy(_as_f32_2d(d_out, d), cfg.act)
        g = {}
        dxq, dxk, dxv = engine.rga_block_bwd(d_a, s, W, cfg, g, dx_addend=None)
        if s["same"]:
            # one tensor fed all three inputs: autograd sums the three returned grads, so hand
            # the combined gradient to the first slot only
            gx = (dxq.view(B, Lq, d), None, None)
        else:
            gx = (dxq.view(B, Lq, d), dxk.view(B, Lq, d), dxv.view(B, Lq, d))
        gw, gb = g["Wqkv"], g["bqkv"]
        return (None, None, None) + gx + (gw[0:d], gb[0:d], gw[d:2 * d], gb[d:2 * d],
                                          gw[2 * d:3 * d], gb[2 * d:3 * d], g["Wfc"], g["bfc"], g["E"])


class EncoderLayer(torch.nn.Module, _PrecisionMixin):
    """Post-LN block: LN1(dropout(rga(x)) + x) -> LN2(. + dropout(FFN_suf(relu(FFN_pre(.)))));
    FFN hidden = d/2, LN eps 1e-6 (MT/layers.py:136-161)."""

    def __init__(self, d_model, rate=0.1, h=16, additional=False, max_seq=2048):
        super().__init__()
        self.d_model = d_model
        self.rga = RelativeGlobalAttention(h=h, d=d_model, max_seq=max_seq, add_emb=additional)
        self.FFN_pre = torch.nn.Linear(self.d_model, self.d_model // 2)
        self.FFN_suf = torch.nn.Linear(self.d_model // 2, self.d_model)
        self.layernorm1 = torch.nn.LayerNorm(self.d_model, eps=1e-6)
        self.layernorm2 = torch.nn.LayerNorm(self.d_model, eps=1e-6)
        self.dropout1 = torch.nn.Dropout(rate)
        self.dropout2 = torch.nn.Dropout(rate)
        self.precision = default_precision()
        # True for the first layer of an Encoder: its input is the un-normalised embedding, so in the bf16
        # mode its attention runs on f16 operands (engine.py); a standalone layer keeps plain bf16
        self.hp_attention = False

    def params(self) -> List[torch.nn.Parameter]:
        return self.rga.params() + [self.FFN_pre.weight, self.FFN_pre.bias, self.FFN_suf.weight,
                                    self.FFN_suf.bias, self.layernorm1.weight, self.layernorm1.bias,
                                    self.layernorm2.weight, self.layernorm2.bias]

    def cfg(self) -> StackCfg:
        c = self.rga.cfg(p_drop=float(self.dropout1.p))
        c.act = _PRECISIONS[self.precision]
        return c

    def weights(self, act) -> LayerWeights:
        return _rga_weights_for(self.rga, act, ffn=(self.FFN_pre, self.FFN_suf),
                                lns=(self.layernorm1, self.layernorm2), hp=self.hp_attention)

    def forward(self, x, mask=None, **kwargs):
        B, Lq, _ = x.shape
        m = as_mask(mask, B, Lq)
        want_w = not self.training
        out, w = _LayerFunction.apply(self, m, want_w, x, *self.params())
        return out, w


def _claim_direct(params) -> Optional[list]:
    """Gradient views of ``params`` that a kernel may write in place: every parameter must belong
    to a FlatAdam whose zero_grad() ran since the parameter's gradient was last written (so the view
    is zero and plain assignment equals accumulation).  Claims them (a second backward before the
    next zero_grad() goes through autograd's accumulation again)."""
    for p in params:
        opt = getattr(p, "_mt_opt", None)
        if opt is None or p.grad is None or id(p) not in opt.fresh or p.grad.dtype != torch.float32 \
                or not p.grad.is_contiguous():
            return None
    for p in params:
        p._mt_opt.fresh.discard(id(p))
    return [p.grad for p in params]


def _adjacent_rows(ts) -> Optional[torch.Tensor]:
    """One [sum rows, ...] view over tensors that sit back to back in one storage, else None."""
    first = ts[0]
    store = first.untyped_storage()
    off = first.storage_offset()
    for t in ts:
        if t.untyped_storage().data_ptr() != store.data_ptr() or t.storage_offset() != off \
                or t.shape[1:] != first.shape[1:] or not t.is_contiguous():
            return None
        off += t.numel()
    rows = sum(t.shape[0] for t in ts)
    return torch.empty(0, dtype=first.dtype, device=first.device).set_(
        store, first.storage_offset(), (rows,) + tuple(first.shape[1:]))


def layer_direct_dst(layer) -> Optional[dict]:
    """Destination dict for engine.layer_bwd (keys of its grads dict) over the layer's parameter
    gradients, or None when they cannot be written in place."""
    ps = layer.params()
    probe = [getattr(p, "_mt_opt", None) is not None and p.grad is not None for p in ps]
    if not all(probe):
        return None
    wq, bq, wk, bk, wv, bv = ps[0], ps[1], ps[2], ps[3], ps[4], ps[5]
    gw = _adjacent_rows([wq.grad, wk.grad, wv.grad])
    gb = _adjacent_rows([bq.grad, bk.grad, bv.grad])
    if gw is None or gb is None:
        return None
    grads = _claim_direct(ps)
    if grads is None:
        return None
    names = ["Wfc", "bfc", "E", "Wpre", "bpre", "Wsuf", "bsuf", "g1", "b1", "g2", "b2"]
    dst = {"Wqkv": gw, "bqkv": gb}
    dst.update({n: grads[6 + i] for i, n in enumerate(names)})
    return dst


def layer_grads_in_param_order(g, d):
    """engine grad dict -> tuple ordered like EncoderLayer.params()."""
    gw, gb = g["Wqkv"], g["bqkv"]
    return (gw[0:d], gb[0:d], gw[d:2 * d], gb[d:2 * d], gw[2 * d:3 * d], gb[2 * d:3 * d], g["Wfc"],
            g["bfc"], g["E"], g["Wpre"], g["bpre"], g["Wsuf"], g["bsuf"], g["g1"], g["b1"], g["g2"],
            g["b2"])


N_LAYER_PARAMS = 17


class _LayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, mask, want_w, x, *params):
        cfg = layer.cfg()
        B, Lq, d = x.shape
        x2 = _as_f32_2d(x, d)
        W = layer.weights(cfg.act)
        out, _, saved, P = engine.layer_fwd(x2, _act_copy(x2, cfg.act), W, cfg, B, Lq, mask, _next_seed(),
                                            1, layer.training, want_w, keep_p=any(ctx.needs_input_grad))
        ctx.cfg, ctx.W, ctx.saved, ctx.shape = cfg, W, saved, (B, Lq, d)
        if P is not None:
            ctx.mark_non_differentiable(P)
        return out.view(B, Lq, d), P

    @staticmethod
    def backward(ctx, d_out, _dP):
        B, Lq, d = ctx.shape
        dx, g = engine.layer_bwd(_as_f32_2d(d_out, d).clone(), ctx.saved, ctx.W, ctx.cfg)
        return (None, None, None, dx.view(B, Lq, d)) + layer_grads_in_param_order(g, d)


class Encoder(torch.nn.Module, _PrecisionMixin):
    """Embedding * sqrt(d) + sinusoid + dropout, then ``num_layers`` EncoderLayers with
    h = d_model // 64 heads (MT/layers.py:207-233)."""

    def __init__(self, num_layers, d_model, input_vocab_size, rate=0.1, max_len=None):
        super().__init__()
        self.d_model = d_model
        self.num_layers = num_layers
        self.embedding = torch.nn.Embedding(num_embeddings=input_vocab_size, embedding_dim=d_model)
        self.pos_encoding = DynamicPositionEmbedding(self.d_model, max_seq=max_len)
        self.enc_layers = torch.nn.ModuleList(
            [EncoderLayer(d_model, rate, h=self.d_model // 64, additional=False, max_seq=max_len)
             for _ in range(num_layers)])
        self.dropout = torch.nn.Dropout(rate)
        self.max_len = max_len
        self.precision = default_precision()
        if num_layers > 0:
            self.enc_layers[0].hp_attention = True

    def params(self) -> List[torch.nn.Parameter]:
        ps = [self.embedding.weight]
        for l in self.enc_layers:
            ps += l.params()
        return ps

    def cfg(self) -> StackCfg:
        c = self.enc_layers[0].cfg()
        c.p_drop = float(self.dropout.p)
        c.act = _PRECISIONS[self.precision]
        return c

    def forward(self, x, mask=None):
        """x int [B, L] -> (hidden [B,L,d] fp32, [attention weights per layer] (None entries
        while training))."""
        if not x.is_cuda:
            raise RuntimeError("musicgeneration_b200 runs on CUDA tensors only (no CPU fallback)")
        B, Lq = x.shape
        if Lq > self.max_len:
            raise RuntimeError(f"sequence length {Lq} exceeds max_seq {self.max_len}")
        m = as_mask(mask, B, Lq)
        want_w = not self.training
        outs = _EncoderFunction.apply(self, m, want_w, x, *self.params())
        return outs[0], list(outs[1:]) if want_w else [None] * self.num_layers


class _EncoderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, mask, want_w, ids, *params):
        cfg = enc.cfg()
        B, Lq = ids.shape
        ids32 = ids.to(torch.int32).contiguous()
        Ws = [l.weights(cfg.act) for l in enc.enc_layers]
        pe = enc.pos_encoding.table(ids.device)
        hid, hid_lp, saved, weights = engine.encoder_fwd(ids32, enc.embedding.weight.data, pe, Ws, cfg,
                                                         mask, _next_seed(), enc.training, want_w,
                                                         keep_p=any(ctx.needs_input_grad))
        ctx.cfg, ctx.Ws, ctx.saved, ctx.enc = cfg, Ws, saved, enc
        ctx.V = enc.embedding.weight.shape[0]
        ctx.shape = (B, Lq, cfg.d)
        ctx.n_w = len(weights) if want_w else 0
        out = hid.view(B, Lq, cfg.d)
        out._mt_lp = hid_lp          # bf16 copy written by the last LayerNorm kernel (if any)
        if want_w:
            ctx.mark_non_differentiable(*weights)
            return (out,) + tuple(weights)
        return (out,)

    @staticmethod
    def backward(ctx, d_hid, *_dw):
        B, Lq, d = ctx.shape
        enc = ctx.enc
        # gradients go straight into the optimizer's flat buffer where that is allowed (FlatAdam,
        # freshly zeroed); autograd then gets None for those parameters
        dsts = [layer_direct_dst(l) for l in enc.enc_layers]
        emb_dst = _claim_direct([enc.embedding.weight])
        opt = getattr(enc.embedding.weight, "_mt_opt", None)

        def layer_done(li):          # data-parallel: this layer's gradient bucket may go on the wire now
            if opt is not None and dsts[li] is not None:
                opt.grads_ready(enc.enc_layers[li].params())

        demb, lg = engine.encoder_bwd(_as_f32_2d(d_hid, d).clone(), ctx.saved, ctx.Ws, ctx.cfg, ctx.V,
                                      dsts=dsts, demb_dst=emb_dst[0] if emb_dst else None, on_layer_done=layer_done)
        if opt is not None and emb_dst:
            opt.grads_ready([enc.embedding.weight])
        grads = (None,) if emb_dst else (demb,)
        for g, dst in zip(lg, dsts):
            grads += (None,) * N_LAYER_PARAMS if dst is not None else layer_grads_in_param_order(g, d)
        return (None, None, None, None) + grads


class _LinearFunction(torch.autograd.Function):
    """nn.Linear on the GEMM kernel (the vocabulary projection, MT/network.py:39)."""

    @staticmethod
    def forward(ctx, cfg, x, weight, bias):
        shape = x.shape
        K = shape[-1]
        N = weight.shape[0]
        x2 = _as_f32_2d(x, K)
        lp = getattr(x, "_mt_lp", None)
        xa = lp if (lp is not None and cfg.act != torch.float32) else _act_copy(x2, cfg.act)
        Wa = _weight_copy(weight.data, cfg.act)
        out = torch.empty((x2.shape[0], N), dtype=torch.float32, device=x.device)
        engine.linear_fwd(xa, Wa, bias.data, out, cfg)
        ctx.cfg, ctx.xa, ctx.Wa, ctx.shape = cfg, xa, Wa, shape
        ctx.weight, ctx.bias = weight, bias
        return out.view(*shape[:-1], N)

    @staticmethod
    def backward(ctx, d_out):
        cfg, xa, Wa = ctx.cfg, ctx.xa, ctx.Wa
        N, K = Wa.shape
        d2 = _as_f32_2d(d_out, N)
        T = d2.shape[0]
        if cfg.act != torch.float32 and N % 8 != 0:
            # the TMA-fed GEMM needs a 16-byte row pitch: copy dY into a zero-padded [T, ld] buffer
            ld = (N + 7) // 8 * 8
            dy_buf = torch.empty((T, ld), dtype=cfg.act, device=d2.device)
            ops.cast2d(d2, N, dy_buf, ld, T, N)
            dy = dy_buf[:, :N]
        else:
            dy = _act_copy(d2, cfg.act)
        direct = _claim_direct([ctx.weight, ctx.bias])
        dW = direct[0] if direct else torch.empty((N, K), dtype=torch.float32, device=dy.device)
        db = direct[1] if direct else torch.empty((N,), dtype=torch.float32, device=dy.device)
        engine.linear_wgrad(dy, xa, dW, db, cfg)
        if direct:
            ctx.weight._mt_opt.grads_ready([ctx.weight, ctx.bias])
        dx = torch.empty((T, K), dtype=torch.float32, device=dy.device)
        engine.linear_dgrad(dy, Wa, dx, cfg)
        return (None, dx.view(ctx.shape), None, None) if direct else (None, dx.view(ctx.shape), dW, db)
