"""Thin torch-tensor wrappers over the C ABI (one function per exported op).

PyTorch is only plumbing here: it owns device memory and the current stream; every function
below enqueues exactly the kernels of one ``mt_*`` entry point on that stream."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L

_DT = {torch.float32: L.MT_F32, torch.bfloat16: L.MT_BF16, torch.float16: L.MT_F16}


def dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    return _DT[d]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


LAUNCH_CALLS = [0]      # number of launching C-ABI calls made so far (each enqueues >= 1 kernel)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = torch.cuda.current_device


def _stream():
    """cudaStream_t of torch's current stream on the current device.  (The public
    torch.cuda.current_stream() builds a Stream object through four Python layers: 14 us per call, a
    quarter of the host time of a train step at ~140 calls.)  The library launches on the CURRENT device
    (include/mt_b200.h); _need_cuda() checks that the tensors live there."""
    LAUNCH_CALLS[0] += 1
    if _raw_stream is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("musicgeneration_b200 runs on CUDA tensors only (no CPU fallback); "
                               "move the module and its inputs to a B200 device")
        if cur is None:
            cur = _cur_device()
        if t.device.index != cur:
            # kernels are enqueued on the current device's stream: a tensor of another device would be
            # dereferenced in the wrong context (one process per GPU is the supported layout)
            raise RuntimeError(f"musicgeneration_b200: tensor on cuda:{t.device.index} but the current device is "
                               f"cuda:{cur}; wrap the call in torch.cuda.device(...) or call torch.cuda.set_device")


_ws_cache = {}


def workspace(nbytes: int, device) -> Optional[torch.Tensor]:
    """Per-device scratch reused by ops that take a caller-owned workspace (stream-ordered)."""
    if nbytes <= 0:
        return None
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def embed_pos_fwd(ids, emb, pe, out_f32, out_lp, pos0, scale, p_drop, seed, site):
    _need_cuda(ids, emb, pe, out_f32)
    B, Lq = ids.shape
    V, d = emb.shape
    L.check(L.load().mt_embed_pos_fwd(_ptr(ids), _ptr(emb), _ptr(pe), _ptr(out_f32), _ptr(out_lp),
                                      dt(out_lp) if out_lp is not None else L.MT_F32, B, Lq, d, V,
                                      pos0, scale, p_drop, seed, site, _stream()), "embed_pos_fwd")


def embed_pos_bwd(ids, dout, demb, scale, p_drop, seed, site):
    _need_cuda(ids, dout, demb)
    B, Lq = ids.shape
    V, d = demb.shape
    L.check(L.load().mt_embed_pos_bwd(_ptr(ids), _ptr(dout), _ptr(demb), B, Lq, d, V, scale, p_drop,
                                      seed, site, _stream()), "embed_pos_bwd")


def add_ln_fwd(a, resid, gamma, beta, out_f32, out_lp, mean, rstd, eps, p_drop, seed, site):
    _need_cuda(a, resid, out_f32)
    T, d = resid.shape
    L.check(L.load().mt_add_ln_fwd(_ptr(a), dt(a), _ptr(resid), _ptr(gamma), _ptr(beta),
                                   _ptr(out_f32), _ptr(out_lp),
                                   dt(out_lp) if out_lp is not None else L.MT_F32, _ptr(mean),
                                   _ptr(rstd), T, d, eps, p_drop, seed, site, _stream()), "add_ln_fwd")


def add_ln_bwd(dout, a, resid, gamma, mean, rstd, dz, da, dgamma, dbeta, p_drop, seed, site, dbias=None):
    """``dbias`` (optional, [d] fp32) receives colsum(da): the bias gradient of the linear producing ``a``."""
    _need_cuda(dout, a, resid)
    T, d = resid.shape
    lib = L.load()
    nparts = lib.mt_add_ln_bwd_parts(T)
    part = torch.empty((3, nparts, d), dtype=torch.float32, device=dout.device)
    L.check(lib.mt_add_ln_bwd(_ptr(dout), _ptr(a), dt(a), _ptr(resid), _ptr(gamma), _ptr(mean),
                              _ptr(rstd), _ptr(dz), _ptr(da), dt(da), _ptr(part), T, d, p_drop, seed,
                              site, _stream()), "add_ln_bwd")
    L.check(lib.mt_ln_param_grad(_ptr(part), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), nparts, d, _stream()),
            "ln_param_grad")


def gemm(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, bias=None, addend=None, aux=None,
         relu=False, relu_mask=False, path=L.PATH_AUTO):
    """C[M,N] = epi(op(A) . op(B)); see include/mt_b200.h."""
    _need_cuda(A, B, C)
    lib = L.load()
    epi = 0
    if bias is not None:
        epi |= L.EPI_BIAS
    if addend is not None:
        epi |= L.EPI_ADD
    if relu:
        epi |= L.EPI_RELU
    if relu_mask:
        epi |= L.EPI_RELU_MASK
    nws = lib.mt_gemm_workspace_bytes(M, N, K, dt(A), path)
    ws = workspace(nws, A.device)
    L.check(lib.mt_gemm(_ptr(A), _ptr(B), _ptr(C), _ptr(bias), _ptr(addend), _ptr(aux), M, N, K, lda,
                        ldb, ldc, int(transA), int(transB), dt(A), dt(C), epi, path, _ptr(ws),
                        ws.numel() if ws is not None else 0, _stream()), "gemm")


def wgrad_bias_supported(dy, x, M, N, path=L.PATH_AUTO) -> bool:
    """Whether mt_wgrad_bias takes dW[M,N] = dy^T x with db = colsum(dy) (tcgen05 weight-gradient form)."""
    return (path != L.PATH_SIMT and dy.dtype in (torch.bfloat16, torch.float16) and x.dtype == dy.dtype
            and N % 128 == 0 and M >= 64 and dy.stride(0) % 8 == 0 and x.stride(0) % 8 == 0)


def wgrad_bias(dy, x, dW, db, M, N, K, lddy, ldx, lddw):
    """dW[M,N] = dy[K,M]^T . x[K,N], db[M] = colsum(dy): one pass over dy (see include/mt_b200.h)."""
    _need_cuda(dy, x, dW, db)
    lib = L.load()
    nws = lib.mt_gemm_workspace_bytes(M, N, K, dt(dy), L.PATH_AUTO)
    ws = workspace(nws, dy.device)
    L.check(lib.mt_wgrad_bias(_ptr(dy), _ptr(x), _ptr(dW), _ptr(db), M, N, K, lddy, ldx, lddw, dt(dy), _ptr(ws),
                              ws.numel() if ws is not None else 0, _stream()), "wgrad_bias")


def colsum(X, out, M, N, ldx):
    _need_cuda(X, out)
    lib = L.load()
    nws = lib.mt_colsum_workspace_bytes(M, N)
    ws = workspace(nws, X.device)
    L.check(lib.mt_colsum(_ptr(X), dt(X), _ptr(out), M, N, ldx, _ptr(ws),
                          ws.numel() if ws is not None else 0, _stream()), "colsum")


def cast(src, dst, n=None):
    _need_cuda(src, dst)
    n = src.numel() if n is None else n
    L.check(L.load().mt_cast(_ptr(src), dt(src), _ptr(dst), dt(dst), n, _stream()), "cast")


def cast2d(src, lds, dst, ldd, rows, cols):
    _need_cuda(src, dst)
    L.check(L.load().mt_cast2d(_ptr(src), dt(src), lds, _ptr(dst), dt(dst), ldd, rows, cols, _stream()),
            "cast2d")


def transpose_cast(src, dst, rows, cols):
    _need_cuda(src, dst)
    L.check(L.load().mt_transpose_cast(_ptr(src), dt(src), _ptr(dst), dt(dst), rows, cols, _stream()),
            "transpose_cast")


def _rga_dtype(q, E, other) -> int:
    """dtype code of an attention call: the tensors' common type, or MT_F16_BF16 for the mixed mode of the
    first encoder layer (f16 q/k/v/E, bf16 everything else)."""
    if E.dtype != q.dtype:
        raise RuntimeError(f"rga: q/k/v are {q.dtype} but E is {E.dtype}")
    if q.dtype == torch.float16 and other.dtype == torch.bfloat16:
        return L.MT_F16_BF16
    if other.dtype != q.dtype:
        raise RuntimeError(f"rga: unsupported dtype combination q {q.dtype} / out {other.dtype}")
    return dt(q)


def rga_stash_new(q, E, O, B, h, Lq, dh) -> Optional[torch.Tensor]:
    """A P stash for one training forward/backward pair of the tcgen05 path (mt_rga_stash_bytes), or None
    when that path does not take the problem.  One per layer: it lives until the layer's backward."""
    need = L.load().mt_rga_stash_bytes(B, h, Lq, dh, _rga_dtype(q, E, O))
    if need == 0:
        return None
    return torch.empty(need, dtype=torch.uint8, device=q.device)


def rga_fwd(q, k, v, strides, E, pad_keys, O, ostrides, lse, B, h, Lq, dh, max_seq, causal,
            path=L.PATH_AUTO, stash: Optional[torch.Tensor] = None):
    """stash: keep the P tiles for the backward (training step on the tcgen05 path, causal only)."""
    _need_cuda(q, k, v, E, O, lse)
    sb, sl, sh = strides
    ob, ol, oh = ostrides
    if stash is not None:
        L.check(L.load().mt_rga_fwd_stash(_ptr(q), _ptr(k), _ptr(v), sb, sl, sh, _ptr(E), _ptr(pad_keys),
                                          _ptr(O), ob, ol, oh, _ptr(lse), B, h, Lq, dh, max_seq, int(causal),
                                          _rga_dtype(q, E, O), _ptr(stash), stash.numel(), _stream()), "rga_fwd_stash")
        return
    L.check(L.load().mt_rga_fwd(_ptr(q), _ptr(k), _ptr(v), sb, sl, sh, _ptr(E), _ptr(pad_keys),
                                _ptr(O), ob, ol, oh, _ptr(lse), B, h, Lq, dh, max_seq, int(causal),
                                _rga_dtype(q, E, O), path, _stream()), "rga_fwd")


def rga_weights(q, k, strides, E, pad_keys, lse, P, B, h, Lq, dh, max_seq, causal):
    _need_cuda(q, k, E, lse, P)
    sb, sl, sh = strides
    L.check(L.load().mt_rga_weights(_ptr(q), _ptr(k), sb, sl, sh, _ptr(E), _ptr(pad_keys), _ptr(lse),
                                    _ptr(P), B, h, Lq, dh, max_seq, int(causal), dt(q), _stream()),
            "rga_weights")


_RGA_WS = {}     # device index -> uint8 scratch shared by every layer's attention backward (one stream)


def rga_bwd_workspace(q, B, h, Lq, dh, code=None):
    """Scratch for the dS-spill variant of the tcgen05 backward (grow-only, one per device); None when
    that variant does not take the problem."""
    need = L.load().mt_rga_bwd_workspace_bytes(B, h, Lq, dh, dt(q) if code is None else code)
    if need == 0:
        return None
    key = q.device.index
    ws = _RGA_WS.get(key)
    if ws is None or ws.numel() < need:
        _RGA_WS[key] = ws = torch.empty(need, dtype=torch.uint8, device=q.device)
    return ws


def rga_bwd(q, k, v, strides, E, pad_keys, O, dO, ostrides, lse, delta, dq, dk, dv, dE, B, h, Lq, dh,
            max_seq, causal, path=L.PATH_AUTO, spill=True, stash: Optional[torch.Tensor] = None):
    """spill=True: give the tcgen05 path its dS workspace (S, P, dS computed once); False: every
    backward role recomputes them (no scratch memory).  stash: the P tiles the forward of this call kept
    (rga_fwd(..., stash=...)): nothing of S / P is rebuilt."""
    _need_cuda(q, k, v, E, O, dO, lse, delta, dq, dk, dv, dE)
    sb, sl, sh = strides
    ob, ol, oh = ostrides
    code = _rga_dtype(q, E, dq)
    if stash is not None:
        ws = rga_bwd_workspace(q, B, h, Lq, dh, code)
        L.check(L.load().mt_rga_bwd_stash(_ptr(q), _ptr(k), _ptr(v), sb, sl, sh, _ptr(E), _ptr(pad_keys),
                                          _ptr(O), _ptr(dO), ob, ol, oh, _ptr(lse), _ptr(delta), _ptr(dq),
                                          _ptr(dk), _ptr(dv), _ptr(dE), B, h, Lq, dh, max_seq, int(causal),
                                          code, _ptr(stash), stash.numel(), _ptr(ws),
                                          ws.numel() if ws is not None else 0, _stream()), "rga_bwd_stash")
        return
    ws = rga_bwd_workspace(q, B, h, Lq, dh, code) if (spill and path != L.PATH_SIMT) else None
    L.check(L.load().mt_rga_bwd_ws(_ptr(q), _ptr(k), _ptr(v), sb, sl, sh, _ptr(E), _ptr(pad_keys),
                                   _ptr(O), _ptr(dO), ob, ol, oh, _ptr(lse), _ptr(delta), _ptr(dq),
                                   _ptr(dk), _ptr(dv), _ptr(dE), B, h, Lq, dh, max_seq, int(causal),
                                   code, path, _ptr(ws), ws.numel() if ws is not None else 0, _stream()),
            "rga_bwd")


def smooth_ce_fwd(logits, target, row_ws, argmax, sums, eps, ignore):
    _need_cuda(logits, target)
    T, V = logits.shape
    L.check(L.load().mt_smooth_ce_fwd(_ptr(logits), _ptr(target), _ptr(row_ws), _ptr(argmax),
                                      _ptr(sums), T, V, eps, ignore, _stream()), "smooth_ce_fwd")


def smooth_ce_bwd(logits, target, row_ws, sums, grad_out, dlogits, eps, ignore):
    T, V = logits.shape
    L.check(L.load().mt_smooth_ce_bwd(_ptr(logits), _ptr(target), _ptr(row_ws), _ptr(sums),
                                      _ptr(grad_out), _ptr(dlogits), T, V, eps, ignore, _stream()),
            "smooth_ce_bwd")


def adam_step(p, g, m, v, p_lp, lr, beta1, beta2, eps, step, grad_scale):
    _need_cuda(p, g, m, v)
    L.check(L.load().mt_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_lp), p.numel(), lr,
                                  beta1, beta2, eps, step, grad_scale, _stream()), "adam_step")


def rga_decode(q, q_stride_b, kcache, vcache, E, pad_keys, out, B, h, dh, max_seq, t):
    L.check(L.load().mt_rga_decode(_ptr(q), q_stride_b, _ptr(kcache), _ptr(vcache), _ptr(E),
                                   _ptr(pad_keys), _ptr(out), B, h, dh, max_seq, t, dt(q), _stream()), "rga_decode")


def kv_append(qkv, kcache, vcache, B, h, dh, max_seq, t):
    L.check(L.load().mt_kv_append(_ptr(qkv), _ptr(kcache), _ptr(vcache), B, h, dh, max_seq, t,
                                  dt(qkv), _stream()), "kv_append")


def sample(logits, u, ids_out, temperature, top_k, greedy):
    B, V = logits.shape
    L.check(L.load().mt_sample(_ptr(logits), _ptr(u), _ptr(ids_out), B, V, temperature, top_k,
                               int(greedy), _stream()), "sample")


# ---- decode with a device-resident step index (CUDA-graph replayable) ----------------------
def decode_embed(ids, t_dev, emb, pe, out_f32, out_lp, scale, pad_token=0, pad_bits=None):
    B, ld = ids.shape
    V, d = emb.shape
    L.check(L.load().mt_decode_embed(_ptr(ids), ld, _ptr(t_dev), _ptr(emb), _ptr(pe), _ptr(out_f32),
                                     _ptr(out_lp), dt(out_lp) if out_lp is not None else L.MT_F32, B, d, V,
                                     scale, pad_token, _ptr(pad_bits),
                                     pad_bits.shape[1] if pad_bits is not None else 0, _stream()), "decode_embed")


def decode_kv_append(qkv, kcache, vcache, ids, pad_token, pad_bits, t_dev, B, h, dh, max_seq):
    L.check(L.load().mt_decode_kv_append(_ptr(qkv), _ptr(kcache), _ptr(vcache), _ptr(ids), ids.shape[1],
                                         pad_token, _ptr(pad_bits), _ptr(t_dev), B, h, dh, max_seq, dt(qkv),
                                         _stream()), "decode_kv_append")


def decode_attend_workspace(B, h, dh, max_seq, device):
    """Zeroed scratch for decode_attend (per-chunk partials + arrival counters); allocate once per session."""
    n = L.load().mt_decode_attend_workspace_bytes(B, h, dh, max_seq) + 16
    return torch.zeros(n, dtype=torch.uint8, device=device)


def decode_attend(q, q_stride_b, kcache, vcache, E, pad_bits, out, t_dev, B, h, dh, max_seq, ws, append=False):
    """append=True: ``q`` is the fused projection row [3, h, dh]; K / V of position t are stored by this launch."""
    L.check(L.load().mt_decode_attend(_ptr(q), q_stride_b, _ptr(kcache), _ptr(vcache), _ptr(E),
                                      _ptr(pad_bits), _ptr(out), _ptr(t_dev), B, h, dh, max_seq, dt(q),
                                      int(append), _ptr(ws), ws.numel(), _stream()), "decode_attend")


def decode_sample(logits, u, ids, t_dev, prior_len, temperature, top_k, greedy):
    B, V = logits.shape
    L.check(L.load().mt_decode_sample(_ptr(logits), _ptr(u), _ptr(ids), ids.shape[1], _ptr(t_dev),
                                      prior_len, B, V, temperature, top_k, int(greedy), _stream()),
            "decode_sample")


def decode_run_supported(B, d, h, V, layers) -> bool:
    return bool(L.load().mt_decode_run_supported(B, d, h, V, layers))


def decode_run(ids, t0, n_steps, prior_len, emb, pe, layer_ptrs, layer_f16, Wv, bv, d, h, max_seq, pad_token, pad_bits,
               uniforms, temperature, top_k, greedy, logits_out, ws):
    """The whole generation in one persistent launch (include/mt_b200.h: mt_decode_run).  ``layer_ptrs``: host int64
    tensor [layers, 15] of device addresses, ``layer_f16``: host int32 tensor [layers]."""
    B, ld = ids.shape
    V = Wv.shape[0]
    L.check(L.load().mt_decode_run(_ptr(ids), ld, B, t0, n_steps, prior_len, _ptr(emb), _ptr(pe),
                                   layer_ptrs.data_ptr(), layer_f16.data_ptr(), layer_ptrs.shape[0], _ptr(Wv), _ptr(bv),
                                   d, h, V, max_seq, pad_token, _ptr(pad_bits), _ptr(uniforms), temperature, top_k,
                                   int(greedy), _ptr(logits_out), _ptr(ws), ws.numel(), _stream()), "decode_run")


def decode_chain(enable: bool):
    """Programmatic dependent launch for the kernels of a decode step (see include/mt_b200.h)."""
    L.load().mt_decode_chain(int(enable))


def decode_advance(t_dev):
    L.check(L.load().mt_decode_advance(_ptr(t_dev), _stream()), "decode_advance")


def window_gather(arena, starts, x, y=None, y_shift=0):
    """x[b, :] = arena[starts[b] : +L], y[b, :] = arena[starts[b] + y_shift : +y_len] (int32 outputs)."""
    _need_cuda(arena, starts, x, y)
    if arena.dtype not in (torch.uint8, torch.uint16) or starts.dtype != torch.int64 or x.dtype != torch.int32:
        raise RuntimeError("window_gather: arena uint8/uint16, starts int64, outputs int32")
    B, Lx = x.shape
    L.check(L.load().mt_window_gather(_ptr(arena), arena.element_size(), _ptr(starts), _ptr(x), _ptr(y), B, Lx,
                                      0 if y is None else y.shape[1], y_shift, _stream()), "window_gather")


def window_sample(file_off, eligible, need, seed, step, starts, files=None):
    """On-device draw of B distinct eligible files and one window start per file (arena offsets)."""
    _need_cuda(file_off, eligible, starts, files)
    L.check(L.load().mt_window_sample(_ptr(file_off), _ptr(eligible), eligible.numel(), need, seed, step,
                                      _ptr(starts), _ptr(files), starts.numel(), _stream()), "window_sample")
