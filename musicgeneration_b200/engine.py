"""Host-side orchestration of the MusicTransformer hot path over the C-ABI ops.

Mirrors the op order of the reference (MT/layers.py:152-161 EncoderLayer, :223-233 Encoder,
:64-109 RelativeGlobalAttention) but every arithmetic step is one of our CUDA kernels; this
file only allocates buffers, keeps what the backward needs and sequences launches.

Precision modes
  * "fp32": all activations fp32, FFMA kernels -- the parity mode (logits/loss 1e-5, greedy
    ids exact).
  * "bf16": fp32 master weights and fp32 residual stream / LayerNorm / softmax / accumulators;
    GEMM and attention operands bf16 (activations are written once in bf16 next to the fp32
    residual by the producing kernel) -- the throughput mode.  The FIRST layer's attention sees the
    un-normalised embedding (MT/layers.py:226-229: |logit| ~ 1e3, softmax nearly one-hot), where a
    bf16 operand error moves the attention targets (logits error 1.6e-2 at config B, above the 1e-2
    bar); its x, Wq/Wk/Wv, q/k/v and E operands are therefore f16 (same tensor-core rate, 11-bit
    mantissa; "hp" below): 3.9e-3 (scripts/sim_precision.py, measured values in DESIGN.md).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from . import ops


@dataclass
class StackCfg:
    d: int
    h: int
    max_seq: int
    p_drop: float = 0.0
    act: torch.dtype = torch.float32          # activation / GEMM operand dtype
    gemm_path: int = L.PATH_AUTO
    attn_path: int = L.PATH_AUTO

    @property
    def dh(self) -> int:
        return self.d // self.h


@dataclass
class LayerWeights:
    """Per-layer operands in the layout the kernels want (act dtype for GEMM operands)."""
    Wqkv: torch.Tensor      # [3d, d] act   (Wq;Wk;Wv stacked: nn.Linear weights are [out,in])
    bqkv: torch.Tensor      # [3d] f32
    Wfc: torch.Tensor       # [d, d] act
    bfc: torch.Tensor
    Wpre: torch.Tensor      # [d/2, d] act
    bpre: torch.Tensor
    Wsuf: torch.Tensor      # [d, d/2] act
    bsuf: torch.Tensor
    E: torch.Tensor         # [max_seq, dh] act
    g1: torch.Tensor
    b1: torch.Tensor
    g2: torch.Tensor
    b2: torch.Tensor
    # f16 operand copies of the first layer's attention in the bf16 mode (None elsewhere)
    Wqkv_hp: Optional[torch.Tensor] = None
    E_hp: Optional[torch.Tensor] = None
    Wfc_hp: Optional[torch.Tensor] = None     # decode only (there the attention output stays f16)


@dataclass
class Mask:
    """Structured form of the reference's look-ahead mask (MT/utils.py:58-83): causal part +
    per-key pad bits.  ``None`` pad_keys = no pad tokens."""
    causal: bool
    pad_keys: Optional[torch.Tensor] = None   # uint8 [B, L]


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


# ---------------------------------------------------------------------------------------
# linear helpers (all row-major; nn.Linear weight W is [out, in])
# ---------------------------------------------------------------------------------------
def linear_fwd(x, W, b, out, cfg: StackCfg, relu=False, ldc=None):
    """out[T, N] = x[T, K] . W[N, K]^T + b"""
    T, K = x.shape
    N = W.shape[0]
    ops.gemm(x, W, out, T, N, K, x.stride(0), W.stride(0), ldc or out.stride(0), False, True,
             bias=b, relu=relu, path=cfg.gemm_path)


def linear_dgrad(dy, W, out, cfg: StackCfg, addend=None, relu_mask_aux=None, ldy=None):
    """out[T, K] = dy[T, N] . W[N, K] (+ addend) (masked by aux > 0)"""
    T = dy.shape[0]
    N, K = W.shape
    ops.gemm(dy, W, out, T, K, N, ldy or dy.stride(0), W.stride(0), out.stride(0), False, False,
             addend=addend, aux=relu_mask_aux, relu_mask=relu_mask_aux is not None,
             path=cfg.gemm_path)


def linear_wgrad(dy, x, dW, db, cfg: StackCfg, ldy=None, n=None):
    """dW[N, K] = dy[T, N]^T . x[T, K];  db[N] = colsum(dy)"""
    T, K = x.shape
    N = n or dy.shape[1]
    ld = ldy or dy.stride(0)
    if db is not None and ops.wgrad_bias_supported(dy, x, N, K, cfg.gemm_path):
        ops.wgrad_bias(dy, x, dW, db, N, K, T, ld, x.stride(0), dW.stride(0))      # column sums ride along in the GEMM
        return
    ops.gemm(dy, x, dW, N, K, T, ld, x.stride(0), dW.stride(0), True, False, path=cfg.gemm_path)
    if db is not None:
        ops.colsum(dy, db, T, N, ld)


# ---------------------------------------------------------------------------------------
# relative global attention block: QKV projection -> fused attention -> fc
# ---------------------------------------------------------------------------------------
def hp_eligible(W: LayerWeights, cfg: StackCfg, mask: Optional[Mask]) -> bool:
    """Whether this layer's attention runs with f16 q/k/v/E operands (the mixed mode of the tcgen05
    kernels: causal mask, head dim 64, bf16 activations elsewhere)."""
    return (W.Wqkv_hp is not None and W.E_hp is not None and cfg.act == torch.bfloat16 and cfg.dh == 64
            and mask is not None and bool(mask.causal)
            and cfg.attn_path != L.PATH_SIMT and cfg.gemm_path != L.PATH_SIMT)


def rga_block_fwd(xq, xk, xv, W: LayerWeights, cfg: StackCfg, B: int, Lq: int, mask: Optional[Mask],
                  need_weights: bool, x_hp: Optional[torch.Tensor] = None, x_f32: Optional[torch.Tensor] = None,
                  keep_p: bool = False):
    """xq/xk/xv: [T, d] act dtype (the same tensor for self-attention).  Returns
    (a [T,d] act dtype = fc output incl. bias, saved dict, P or None).
    ``x_hp`` (f16 copy of the self-attention input) selects the f16-operand mode of the first layer;
    ``xq`` may then be None (the bf16 copy the weight gradient needs is cast from ``x_f32`` in the backward).
    ``keep_p``: a backward will follow -- on the tcgen05 path (causal mask, head dim 64, 16-bit operands) the attention
    kernel then keeps its P tiles in a per-layer stash for it (the reference keeps the softmax output through
    autograd, MT/layers.py:97-99)."""
    d, h, dh = cfg.d, cfg.h, cfg.dh
    T = B * Lq
    hp = x_hp is not None
    like = x_hp if hp else xq
    qkv = _empty((T, 3 * d), torch.float16 if hp else cfg.act, like)
    same = hp or ((xq is xk) and (xk is xv))
    E = W.E_hp if hp else W.E
    if hp:
        linear_fwd(x_hp, W.Wqkv_hp, W.bqkv, qkv, cfg)
    elif same:
        linear_fwd(xq, W.Wqkv, W.bqkv, qkv, cfg)
    else:
        for i, x in enumerate((xq, xk, xv)):
            linear_fwd(x, W.Wqkv[i * d:(i + 1) * d], W.bqkv[i * d:(i + 1) * d], qkv[:, i * d:(i + 1) * d],
                       cfg, ldc=3 * d)
    q, k, v = qkv[:, 0:d], qkv[:, d:2 * d], qkv[:, 2 * d:3 * d]
    strides = (Lq * 3 * d, 3 * d, dh)          # (batch, position, head) in elements
    O = _empty((T, d), cfg.act, like)          # (bf16 also in the f16-operand mode: the kernel packs O separately)
    ostrides = (Lq * d, d, dh)
    lse = _empty((B, h, Lq), torch.float32, like)
    causal = bool(mask.causal) if mask is not None else False
    pad = mask.pad_keys if mask is not None else None
    stash = None
    if keep_p and causal and cfg.attn_path != L.PATH_SIMT and qkv.dtype != torch.float32:
        stash = ops.rga_stash_new(q, E, O, B, h, Lq, dh)          # None when the tcgen05 pair does not take the shape
    ops.rga_fwd(q, k, v, strides, E, pad, O, ostrides, lse, B, h, Lq, dh, cfg.max_seq, causal,
                path=cfg.attn_path, stash=stash)
    P = None
    if need_weights:
        P = _empty((B, h, Lq, Lq), torch.float32, like)
        ops.rga_weights(q, k, strides, E, pad, lse, P, B, h, Lq, dh, cfg.max_seq, causal)
    # sublayer output in the activation dtype (bf16 mode: what a bf16 nn.Linear returns; it is read twice more,
    # by the residual+LayerNorm forward and backward, so the narrower type saves 3 x T x d x 2 bytes per sublayer)
    a = _empty((T, d), cfg.act, like)
    linear_fwd(O, W.Wfc, W.bfc, a, cfg)
    saved = dict(xq=xq, xk=xk, xv=xv, same=same, qkv=qkv, O=O, lse=lse, causal=causal, pad=pad,
                 strides=strides, ostrides=ostrides, B=B, L=Lq, hp=hp, x_f32=x_f32 if hp else None, stash=stash)
    return a, saved, P


def _gbuf(dst, name, shape, like, zero=False):
    """Gradient destination: the caller's buffer (a zeroed view of the flat gradient buffer) or a temporary."""
    if dst is not None and name in dst:
        return dst[name]
    if zero:
        return torch.zeros(shape, dtype=torch.float32, device=like.device)
    return _empty(shape, torch.float32, like)


def rga_block_bwd(d_a, s, W: LayerWeights, cfg: StackCfg, g: Dict[str, torch.Tensor],
                  dx_addend: Optional[torch.Tensor], dst: Optional[Dict[str, torch.Tensor]] = None):
    """d_a [T,d] act dtype = grad wrt fc output.  Fills g[...] (fp32 grads) and returns
    (dxq, dxk, dxv) fp32 [T,d]; for self-attention the three are one tensor that already
    includes ``dx_addend`` (the residual-path gradient)."""
    d, h, dh = cfg.d, cfg.h, cfg.dh
    B, Lq = s["B"], s["L"]
    T = B * Lq
    dev = d_a
    g["Wfc"] = _gbuf(dst, "Wfc", (d, d), dev)
    have_bfc = "bfc" in g                              # already produced by the LayerNorm backward
    if not have_bfc:
        g["bfc"] = _gbuf(dst, "bfc", (d,), dev)
    linear_wgrad(d_a, s["O"], g["Wfc"], None if have_bfc else g["bfc"], cfg)
    dO = _empty((T, d), cfg.act, dev)
    linear_dgrad(d_a, W.Wfc, dO, cfg)
    dqkv = _empty((T, 3 * d), cfg.act, dev)
    delta = _empty((B, h, Lq), torch.float32, dev)
    g["E"] = _gbuf(dst, "E", (cfg.max_seq, dh), dev, zero=True)
    qkv = s["qkv"]
    ops.rga_bwd(qkv[:, 0:d], qkv[:, d:2 * d], qkv[:, 2 * d:3 * d], s["strides"], W.E_hp if s["hp"] else W.E, s["pad"],
                s["O"], dO, s["ostrides"], s["lse"], delta, dqkv[:, 0:d], dqkv[:, d:2 * d],
                dqkv[:, 2 * d:3 * d], g["E"], B, h, Lq, dh, cfg.max_seq, s["causal"],
                path=cfg.attn_path, stash=s.get("stash"))
    s["stash"] = None           # (its last use: the caching allocator may hand the block to the next layer's scratch)
    g["Wqkv"] = _gbuf(dst, "Wqkv", (3 * d, d), dev)
    g["bqkv"] = _gbuf(dst, "bqkv", (3 * d,), dev)
    if s["same"]:
        xw = s["xq"]
        if xw is None:              # f16-operand layer: the weight gradient takes a bf16 copy of the input (dqkv is bf16)
            xw = _empty((T, d), cfg.act, dev)
            ops.cast(s["x_f32"], xw)
        linear_wgrad(dqkv, xw, g["Wqkv"], g["bqkv"], cfg)
        # dx = dx_addend + dqkv . Wqkv, accumulated IN PLACE into the residual-path gradient (its last
        # use): the GEMM epilogue then needs no addend read (red.global.add, see gemm_tc.cu)
        dx = dx_addend if dx_addend is not None else _empty((T, d), torch.float32, dev)
        linear_dgrad(dqkv, W.Wqkv, dx, cfg, addend=dx_addend)
        return dx, dx, dx
    outs = []
    for i, x in enumerate((s["xq"], s["xk"], s["xv"])):
        sl = dqkv[:, i * d:(i + 1) * d]
        linear_wgrad(sl, x, g["Wqkv"][i * d:(i + 1) * d], g["bqkv"][i * d:(i + 1) * d], cfg,
                     ldy=3 * d, n=d)
        dx = _empty((T, d), torch.float32, dev)
        linear_dgrad(sl, W.Wqkv[i * d:(i + 1) * d], dx, cfg, ldy=3 * d)
        outs.append(dx)
    return tuple(outs)


# ---------------------------------------------------------------------------------------
# encoder layer  (MT/layers.py:152-161)
# ---------------------------------------------------------------------------------------
def layer_fwd(x_f32, x_lp, W: LayerWeights, cfg: StackCfg, B: int, Lq: int, mask: Optional[Mask],
              seed: int, site0: int, training: bool, need_weights: bool, keep_p: bool = False):
    d = cfg.d
    T = B * Lq
    p = cfg.p_drop if training else 0.0
    lp = cfg.act != torch.float32
    if hp_eligible(W, cfg, mask):
        # first layer, bf16 mode: f16 operands for the projection and the attention logits
        if x_lp is not None and x_lp.dtype == torch.float16:
            x_hp, x_b = x_lp, None
        else:
            x_hp, x_b = _empty((T, d), torch.float16, x_f32), x_lp
            ops.cast(x_f32, x_hp)
        a, s_att, P = rga_block_fwd(x_b, x_b, x_b, W, cfg, B, Lq, mask, need_weights, x_hp=x_hp, x_f32=x_f32,
                                    keep_p=keep_p)
    else:
        a, s_att, P = rga_block_fwd(x_lp, x_lp, x_lp, W, cfg, B, Lq, mask, need_weights, keep_p=keep_p)
    out1 = _empty((T, d), torch.float32, x_f32)
    out1_lp = _empty((T, d), cfg.act, x_f32) if lp else None
    mean1 = _empty((T,), torch.float32, x_f32)
    rstd1 = _empty((T,), torch.float32, x_f32)
    ops.add_ln_fwd(a, x_f32, W.g1, W.b1, out1, out1_lp, mean1, rstd1, 1e-6, p, seed, site0)
    o1 = out1_lp if lp else out1
    hmid = _empty((T, d // 2), cfg.act, x_f32)
    linear_fwd(o1, W.Wpre, W.bpre, hmid, cfg, relu=True)
    f = _empty((T, d), cfg.act, x_f32)
    linear_fwd(hmid, W.Wsuf, W.bsuf, f, cfg)
    out2 = _empty((T, d), torch.float32, x_f32)
    out2_lp = _empty((T, d), cfg.act, x_f32) if lp else None
    mean2 = _empty((T,), torch.float32, x_f32)
    rstd2 = _empty((T,), torch.float32, x_f32)
    ops.add_ln_fwd(f, out1, W.g2, W.b2, out2, out2_lp, mean2, rstd2, 1e-6, p, seed, site0 + 1)
    saved = dict(att=s_att, x=x_f32, a=a, out1=out1, o1=o1, mean1=mean1, rstd1=rstd1, hmid=hmid,
                 f=f, mean2=mean2, rstd2=rstd2, p=p, seed=seed, site0=site0)
    return out2, (out2_lp if lp else out2), saved, P


def layer_bwd(dout, s, W: LayerWeights, cfg: StackCfg,
              dst: Optional[Dict[str, torch.Tensor]] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """dout [T,d] f32 (may be overwritten).  Returns (dx f32, grads dict keyed like LayerWeights).
    ``dst``: zeroed gradient buffers (keys of the grads dict) to write into instead of temporaries."""
    d = cfg.d
    T = dout.shape[0]
    g: Dict[str, torch.Tensor] = {}
    dev = dout
    p, seed, site0 = s["p"], s["seed"], s["site0"]
    # LN2
    g["g2"] = _gbuf(dst, "g2", (d,), dev)
    g["b2"] = _gbuf(dst, "b2", (d,), dev)
    d_f = _empty((T, d), cfg.act, dev)
    g["bsuf"] = _gbuf(dst, "bsuf", (d,), dev)         # colsum(d_f) comes out of the LayerNorm backward
    ops.add_ln_bwd(dout, s["f"], s["out1"], W.g2, s["mean2"], s["rstd2"], dout, d_f, g["g2"], g["b2"],
                   p, seed, site0 + 1, dbias=g["bsuf"])
    dz2 = dout
    # FFN_suf
    g["Wsuf"] = _gbuf(dst, "Wsuf", (d, d // 2), dev)
    linear_wgrad(d_f, s["hmid"], g["Wsuf"], None, cfg)
    dh_ = _empty((T, d // 2), cfg.act, dev)
    linear_dgrad(d_f, W.Wsuf, dh_, cfg, relu_mask_aux=s["hmid"])
    # FFN_pre
    g["Wpre"] = _gbuf(dst, "Wpre", (d // 2, d), dev)
    g["bpre"] = _gbuf(dst, "bpre", (d // 2,), dev)
    linear_wgrad(dh_, s["o1"], g["Wpre"], g["bpre"], cfg)
    linear_dgrad(dh_, W.Wpre, dz2, cfg, addend=dz2)          # d_out1 = dz2 + dh . Wpre (in place)
    d_out1 = dz2
    # LN1
    g["g1"] = _gbuf(dst, "g1", (d,), dev)
    g["b1"] = _gbuf(dst, "b1", (d,), dev)
    d_a = _empty((T, d), cfg.act, dev)
    g["bfc"] = _gbuf(dst, "bfc", (d,), dev)           # colsum(d_a), likewise
    ops.add_ln_bwd(d_out1, s["a"], s["x"], W.g1, s["mean1"], s["rstd1"], d_out1, d_a, g["g1"], g["b1"],
                   p, seed, site0, dbias=g["bfc"])
    dz1 = d_out1
    dx, _, _ = rga_block_bwd(d_a, s["att"], W, cfg, g, dx_addend=dz1, dst=dst)
    return dx, g


# ---------------------------------------------------------------------------------------
# whole stack  (MT/layers.py:223-233)
# ---------------------------------------------------------------------------------------
def encoder_fwd(ids: torch.Tensor, emb: torch.Tensor, pe: torch.Tensor, Ws: List[LayerWeights],
                cfg: StackCfg, mask: Optional[Mask], seed: int, training: bool, need_weights: bool,
                pos0: int = 0, keep_p: bool = False):
    B, Lq = ids.shape
    d = cfg.d
    T = B * Lq
    lp = cfg.act != torch.float32
    p = cfg.p_drop if training else 0.0
    x = _empty((T, d), torch.float32, emb)
    # the first layer's 16-bit input copy is f16 when that layer runs its attention on f16 operands
    lp0 = torch.float16 if (Ws and hp_eligible(Ws[0], cfg, mask)) else cfg.act
    x_lp = _empty((T, d), lp0, emb) if lp else None
    ops.embed_pos_fwd(ids, emb, pe, x, x_lp, pos0, math.sqrt(d), p, seed, 0)
    xl = x_lp if lp else x
    saved_layers = []
    weights = []
    for li, W in enumerate(Ws):
        x, xl, s, P = layer_fwd(x, xl, W, cfg, B, Lq, mask, seed, 1 + 2 * li, training, need_weights, keep_p=keep_p)
        saved_layers.append(s)
        weights.append(P)
    saved = dict(ids=ids, layers=saved_layers, p=p, seed=seed, B=B, L=Lq)
    return x, xl, saved, weights


def encoder_bwd(dhid: torch.Tensor, saved, Ws: List[LayerWeights], cfg: StackCfg, V: int,
                dsts: Optional[List[Optional[Dict[str, torch.Tensor]]]] = None,
                demb_dst: Optional[torch.Tensor] = None, on_layer_done=None):
    """dhid [T,d] f32 -> (demb [V,d] f32, [layer grad dicts]).  ``dsts`` / ``demb_dst``: zeroed
    gradient buffers to write into (see layer_bwd).  ``on_layer_done(li)`` is called once layer li's
    backward kernels have been enqueued (its gradients are final in stream order: the data-parallel exchange
    of that layer's bucket may start while the earlier layers' backward runs)."""
    dx = dhid
    layer_grads: List[Dict[str, torch.Tensor]] = [None] * len(Ws)
    for li in range(len(Ws) - 1, -1, -1):
        dx, layer_grads[li] = layer_bwd(dx, saved["layers"][li], Ws[li], cfg,
                                        dst=dsts[li] if dsts is not None else None)
        if on_layer_done is not None:
            on_layer_done(li)
    demb = demb_dst if demb_dst is not None else \
        torch.zeros((V, cfg.d), dtype=torch.float32, device=dhid.device)
    ops.embed_pos_bwd(saved["ids"], dx, demb, math.sqrt(cfg.d), saved["p"], saved["seed"], 0)
    return demb, layer_grads
