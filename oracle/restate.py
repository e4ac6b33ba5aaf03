"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the MusicTransformer hot path.

This file restates, in plain functional PyTorch-on-CPU / NumPy, the algorithm of the
reference path named by BASELINE.json (``mg/model/MusicTransformer``; ``MT/`` below):
mask build, embedding + sinusoid, relative global attention with the skew, post-LN encoder
layer, vocabulary projection, label-smoothed cross entropy, the step metrics and the
autoregressive sampling loop.  It is the checker for the CUDA path and the ``cpu_baseline``
("port") of bench.py.  The product package ``musicgeneration_b200`` never imports it.

Pinning: the reference ships no tests / golden vectors, so the oracle is
pinned against the UNMODIFIED reference modules executed in the build container
(``oracle/ref_import.py`` + ``oracle/make_golden.py``; results under ``tests/golden``) --
``tests/test_oracle.py`` re-checks those fixtures everywhere and re-runs the live comparison
when ``/root/reference`` is present.

Two statements of the attention are kept on purpose:
  * ``rga_forward``       -- follows the reference op by op (einsum with E, QE masking, the
                             pad+reshape skew, matmul, scale, additive -1e9 mask, softmax);
  * ``rga_closed_form``   -- the index form  S[i,j] = (q_i.k_j + [j<=i] q_i.E[max_seq-1-(i-j)])
                             / sqrt(dh)  that the CUDA kernels implement.
The tests require both to agree with each other and with the reference.

Parameters are passed as a plain ``dict`` keyed exactly like the reference ``state_dict``
(``Decoder.embedding.weight``, ``Decoder.enc_layers.{i}.rga.E`` ... ``fc.bias``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------
# MT/layers.py:22-39  DynamicPositionEmbedding
# ----------------------------------------------------------------------------------------
def sinusoid_table(max_seq: int, d: int) -> np.ndarray:
    """float64 ``[max_seq, d]`` table: sin(pos*exp(-ln(1e4)*i/d)*exp(ln(1e4)/d*(i%2)) + pi/2*(i%2))

    MT/layers.py:25-34 builds this with Python scalar loops; the vectorised form below performs
    the same float64 operations in the same order per element."""
    pos = np.arange(max_seq, dtype=np.float64)[:, None]
    i = np.arange(d, dtype=np.float64)[None, :]
    par = np.arange(d)[None, :] % 2
    ang = pos * np.exp(-math.log(10000) * i / d) * np.exp(math.log(10000) / d * par) \
        + 0.5 * math.pi * par
    return np.sin(ang)


# ----------------------------------------------------------------------------------------
# MT/utils.py:58-83,183-188  get_masked_with_pad_tensor / sequence_mask
# ----------------------------------------------------------------------------------------
def look_ahead_mask(x: torch.Tensor, pad: int, size: Optional[int] = None) -> torch.Tensor:
    """bool ``[B,1,size,size]``: True = masked.  mask[b,0,i,j] = (x[b,j]==pad) | (j>i).

    MT/utils.py:73 builds the pad part as ``trg == pad`` broadcast from [B,1,1,L];
    :75 the causal part as ``~(arange(size) < arange(1,size+1)[:,None])``; :77 ORs them.  The OR
    broadcasts [B,1,1,L] against [size,size], which is why forward() needs L == max_seq."""
    L = x.size(1)
    size = L if size is None else size
    padm = (x == pad)[:, None, None, :]
    ar = torch.arange(size)
    causal = ~(ar[None, :] < (ar + 1)[:, None])
    return padm | causal


# ----------------------------------------------------------------------------------------
# MT/layers.py:64-133  RelativeGlobalAttention.forward
# ----------------------------------------------------------------------------------------
def _split_heads(t: torch.Tensor, h: int) -> torch.Tensor:
    B, L, d = t.shape
    return t.reshape(B, L, h, d // h).permute(0, 2, 1, 3)


def rga_scores(q: torch.Tensor, k: torch.Tensor, E: torch.Tensor, max_seq: int) -> torch.Tensor:
    """Scaled, unmasked logits following MT/layers.py:89-97 literally (len_q == len_k)."""
    B, h, L, dh = q.shape
    e = E[max(0, max_seq - L):, :]                                  # :111-114
    qe = torch.einsum("bhld,md->bhlm", q, e)                        # :90
    # :127-133  keep column m of row l only if m >= L-1-l
    lengths = torch.arange(L - 1, -1, -1)
    keep = ~(torch.arange(L)[None, :] < lengths[:, None])
    qe = keep.to(qe.dtype) * qe
    # :116-119  pad one column on the left, view as [L+1, L], drop the first row
    padded = F.pad(qe, [1, 0])
    srel = padded.reshape(B, h, L + 1, L)[:, :, 1:, :]
    qk = torch.matmul(q, k.transpose(-1, -2))                       # :94-95
    return (qk + srel) / math.sqrt(dh)                              # :96-97


def rga_forward(x: torch.Tensor, p: Params, prefix: str, h: int, max_seq: int,
                mask: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [B,L,d] -> (out [B,L,d], attention weights [B,h,L,L]); inputs are [x,x,x] as in
    MT/layers.py:153."""
    q = _split_heads(F.linear(x, p[prefix + "Wq.weight"], p[prefix + "Wq.bias"]), h)
    k = _split_heads(F.linear(x, p[prefix + "Wk.weight"], p[prefix + "Wk.bias"]), h)
    v = _split_heads(F.linear(x, p[prefix + "Wv.weight"], p[prefix + "Wv.bias"]), h)
    logits = rga_scores(q, k, p[prefix + "E"], max_seq)
    if mask is not None:
        logits = logits + (mask.to(torch.int64) * -1e9).to(logits.dtype)   # :99-100
    w = F.softmax(logits, -1)                                        # :102
    a = torch.matmul(w, v)                                           # :103
    B, _, L, _ = a.shape
    out = a.permute(0, 2, 1, 3).reshape(B, L, -1)                    # :105-106
    out = F.linear(out, p[prefix + "fc.weight"], p[prefix + "fc.bias"])   # :108
    return out, w


def rga_closed_form(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, E: torch.Tensor,
                    max_seq: int, causal: bool,
                    pad_keys: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Index form used by the kernels.  q,k,v [B,h,L,dh]; returns (O [B,h,L,dh], LSE [B,h,L]).

    S[i,j] = (q_i.k_j + [j<=i] q_i.E[max_seq-1-(i-j)]) / sqrt(dh); masked entries (j>i when
    causal, pad keys) are excluded (the reference adds -1e9, which is the same thing in
    floating point for every row that keeps at least one key)."""
    B, h, L, dh = q.shape
    i = torch.arange(L)[:, None]
    j = torch.arange(L)[None, :]
    idx = (max_seq - 1 - (i - j)).clamp(0, max_seq - 1)              # [L,L]
    qe = torch.einsum("bhld,md->bhlm", q, E)                         # all rows of E
    srel = torch.gather(qe, 3, idx.expand(B, h, L, L)) * (j <= i).to(q.dtype)
    s = (torch.matmul(q, k.transpose(-1, -2)) + srel) / math.sqrt(dh)
    dead = torch.zeros(B, 1, L, L, dtype=torch.bool)
    if causal:
        dead = dead | (j > i)
    if pad_keys is not None:
        dead = dead | pad_keys[:, None, None, :]
    s = s.masked_fill(dead, float("-inf"))
    lse = torch.logsumexp(s, -1)
    return torch.matmul(torch.exp(s - lse[..., None]), v), lse


# ----------------------------------------------------------------------------------------
# MT/layers.py:152-161  EncoderLayer.forward ; :223-233 Encoder.forward (dropout = identity)
# ----------------------------------------------------------------------------------------
def encoder_layer_forward(x: torch.Tensor, p: Params, prefix: str, h: int, max_seq: int,
                          mask: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    d = x.size(-1)
    a, w = rga_forward(x, p, prefix + "rga.", h, max_seq, mask)
    o1 = F.layer_norm(a + x, (d,), p[prefix + "layernorm1.weight"],
                      p[prefix + "layernorm1.bias"], 1e-6)
    f = F.relu(F.linear(o1, p[prefix + "FFN_pre.weight"], p[prefix + "FFN_pre.bias"]))
    f = F.linear(f, p[prefix + "FFN_suf.weight"], p[prefix + "FFN_suf.bias"])
    o2 = F.layer_norm(o1 + f, (d,), p[prefix + "layernorm2.weight"],
                      p[prefix + "layernorm2.bias"], 1e-6)
    return o2, w


def num_layers_of(p: Params) -> int:
    n = 0
    while f"Decoder.enc_layers.{n}.rga.E" in p:
        n += 1
    return n


def encoder_forward(ids: torch.Tensor, p: Params, max_seq: int,
                    mask: Optional[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    emb = p["Decoder.embedding.weight"]
    d = emb.size(1)
    h = d // 64                                                       # MT/layers.py:219
    x = emb[ids.long()] * math.sqrt(d)                                # :226-227
    pe = torch.from_numpy(sinusoid_table(max_seq, d)[None, :ids.size(1), :])
    x = x + pe.to(x.dtype)                                            # :38
    ws = []
    for l in range(num_layers_of(p)):
        x, w = encoder_layer_forward(x, p, f"Decoder.enc_layers.{l}.", h, max_seq, mask)
        ws.append(w)
    return x, ws


# ----------------------------------------------------------------------------------------
# MT/network.py:35-42  MusicTransformer.forward (train / eval)
# ----------------------------------------------------------------------------------------
def model_forward(ids: torch.Tensor, p: Params, max_seq: int, pad: int,
                  return_weights: bool = False):
    mask = look_ahead_mask(ids, pad, max_seq)                         # network.py:37
    hid, ws = encoder_forward(ids, p, max_seq, mask)
    logits = F.linear(hid, p["fc.weight"], p["fc.bias"]).contiguous()
    return (logits, ws) if return_weights else logits


# ----------------------------------------------------------------------------------------
# MT/criterion.py:43-67  SmoothCrossEntropyLoss ; MT/metrics.py:40-60
# ----------------------------------------------------------------------------------------
def smooth_ce(logits: torch.Tensor, target: torch.Tensor, eps: float, vocab: int,
              ignore: int) -> torch.Tensor:
    dead = (target == ignore).unsqueeze(-1)
    q = F.one_hot(target.long(), vocab).to(torch.float32)
    qp = ((1.0 - eps) * q + eps / vocab).masked_fill(dead, 0)
    ce = -(qp * (logits - logits.logsumexp(-1, keepdim=True))).sum(-1)
    return ce.sum() / (target != ignore).sum()


def categorical_accuracy(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """MT/metrics.py:50-52 + Accuracy base: mean(argmax(softmax(z)) == y) over ALL positions."""
    return (logits.softmax(-1).argmax(-1) == target).to(torch.float32).mean()


def logits_bucket(logits: torch.Tensor) -> torch.Tensor:
    return logits.argmax(-1).flatten().to(torch.int32)               # MT/metrics.py:60


# ----------------------------------------------------------------------------------------
# MT/network.py:44-80  generate  (two oracles)
# ----------------------------------------------------------------------------------------
def generate_literal_greedy(prior: torch.Tensor, length: int, p: Params, max_seq: int,
                            threshold_len: int) -> torch.Tensor:
    """network.py:52-77 as written -- NO mask (``self.Decoder(decode_array, None)``), full
    recompute, sliding window at ``threshold_len`` -- with the reference's own (dead) greedy
    branch arithmetic (:68-71) in place of OneHotCategorical sampling."""
    dec = prior.clone()
    res = prior.clone()
    for _ in range(length):
        if dec.size(1) >= threshold_len:
            dec = dec[:, 1:]
        hid, _ = encoder_forward(dec, p, max_seq, None)
        pr = F.linear(hid, p["fc.weight"], p["fc.bias"]).softmax(-1)
        nxt = pr[:, -1].argmax(-1).to(dec.dtype).unsqueeze(-1)
        dec = torch.cat((dec, nxt), -1)
        res = torch.cat((res, nxt), -1)
    return res


def generate_causal_greedy(prior: torch.Tensor, length: int, p: Params, max_seq: int,
                           pad: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Oracle for the KV-cached decode: same loop, but the look-ahead mask the reference builds
    at network.py:55-56 (and then drops) IS passed to the stack; no sliding window
    (prior+length <= max_seq).  Returns (ids [B,P+length], last-position logits of every step
    [length,B,V])."""
    dec = prior.clone()
    steps = []
    for _ in range(length):
        mask = look_ahead_mask(dec, pad, dec.size(1))
        hid, _ = encoder_forward(dec, p, max_seq, mask)
        z = F.linear(hid[:, -1], p["fc.weight"], p["fc.bias"])
        steps.append(z)
        nxt = z.argmax(-1).to(dec.dtype).unsqueeze(-1)
        dec = torch.cat((dec, nxt), -1)
    return dec, torch.stack(steps)


def sample_topk_from_uniform(logits: torch.Tensor, u: torch.Tensor, temperature: float,
                             top_k: int) -> torch.Tensor:
    """Our sampling semantics (absent in MT; temperature as in
    mg/model/Event_MelodyRNN/network.py:90-96): z/T -> keep the top-k (ties: lower id first)
    -> softmax -> inverse CDF over ids in ascending id order with the given uniforms u [B]."""
    z = logits.to(torch.float32) / temperature
    V = z.size(-1)
    if 0 < top_k < V:
        # stable selection: sort by (-value, id)
        order = torch.argsort(-z, dim=-1, stable=True)
        keep = torch.zeros_like(z, dtype=torch.bool)
        keep.scatter_(1, order[:, :top_k], True)
        z = z.masked_fill(~keep, float("-inf"))
    pr = torch.softmax(z, -1)
    cdf = torch.cumsum(pr, -1)
    tgt = u[:, None] * cdf[:, -1:]
    idx = (cdf <= tgt).sum(-1).clamp(max=V - 1)
    # never return an excluded id
    alive = pr > 0
    last_alive = (alive * torch.arange(V)[None, :]).max(-1).values
    return torch.minimum(idx, last_alive)


# ----------------------------------------------------------------------------------------
# MT/criterion.py:70-96  CustomSchedule.rate
# ----------------------------------------------------------------------------------------
def noam_rate(step: int, d_model: int, warmup: int = 4000) -> float:
    return d_model ** (-0.5) * min(step ** (-0.5), step * warmup ** (-1.5))


# ----------------------------------------------------------------------------------------
# deterministic synthetic inputs
# ----------------------------------------------------------------------------------------
def synthetic_ids(B: int, L: int, pad: int, seed: int = 1234) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, pad, (B, L), generator=g, dtype=torch.int32)
    y = torch.randint(0, pad, (B, L), generator=g, dtype=torch.int32)
    return x, y


def init_params(d: int, vocab: int, layers: int, max_seq: int, seed: int = 0) -> Params:
    """Random init with the reference's parameter shapes and distributions (torch defaults for
    Embedding/Linear/LayerNorm, E ~ randn -- MT/layers.py:60) drawn in state_dict order from a
    private generator.  NOT bit-identical to constructing the reference module under
    torch.manual_seed (module construction order differs); parity tests that need identical
    weights load the golden state_dict instead."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def lin(name, out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        p[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
        p[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * b

    p["Decoder.embedding.weight"] = torch.randn(vocab, d, generator=g)
    for l in range(layers):
        pre = f"Decoder.enc_layers.{l}."
        lin(pre + "rga.Wq", d, d)
        lin(pre + "rga.Wk", d, d)
        lin(pre + "rga.Wv", d, d)
        lin(pre + "rga.fc", d, d)
        p[pre + "rga.E"] = torch.randn(max_seq, 64, generator=g)
        lin(pre + "FFN_pre", d // 2, d)
        lin(pre + "FFN_suf", d, d // 2)
        for n in ("layernorm1", "layernorm2"):
            p[pre + n + ".weight"] = torch.ones(d)
            p[pre + n + ".bias"] = torch.zeros(d)
    lin("fc", vocab, d)
    return p


# ----------------------------------------------------------------------------------------
# MT/data.py:10-107  Data (batch builders) -- host restatement that re-reads the files
# ----------------------------------------------------------------------------------------
def write_token_corpus(root: str, n_files: int = 24, seed: int = 5, lo: int = 70, hi: int = 420,
                       vocab: int = 388, dtype=np.uint16) -> List[str]:
    """Synthetic corpus in the reference's on-disk format: one ``torch.save(np.ndarray)`` per piece,
    named ``*.data`` (REF/mg/model/utils/preprocess_MIDI_like.py:21-41).  Deterministic in ``seed``;
    a few pieces sit exactly at lengths the tests use as window sizes (boundary cases of
    ``_get_seq``).  Returns the file names in creation order."""
    import os as _os
    rng = np.random.RandomState(seed)
    names = []
    for i in range(n_files):
        n = int(rng.randint(lo, hi))
        if i % 7 == 3:
            n = 65            # == 64 + 1: random.randrange(0, 0) -> ValueError for slide batches of 64
        if i % 11 == 5:
            n = 40            # passes a max_length = 30 filter, too short for most windows -> IndexError
        arr = rng.randint(0, vocab, size=n).astype(dtype)
        sub = _os.path.join(root, f"d{i % 3}")
        _os.makedirs(sub, exist_ok=True)
        name = _os.path.join(sub, f"piece{i:03d}-{rng.randint(0, 1 << 30):08x}.data")
        torch.save(arr, name)
        names.append(name)
    return names


class DataOracle:
    """MT/data.py:10-107 restated on the host with numpy: same ``random`` call sequence (one
    ``random.sample`` per batch, one ``random.randrange`` per drawn file), same split, same errors.
    ``files`` may be given to fix the listing order (the reference takes ``os.walk`` order)."""

    def __init__(self, dir_path, max_length, files=None):
        import os as _os
        if files is None:
            files = []
            for path, _, names in _os.walk(dir_path):                       # MT/utils.py:19-22
                files += [_os.path.join(path, n) for n in names if n.lower().endswith('.data')]
        self.files = list(files)
        self._data = {f: np.asarray(torch.load(f, weights_only=False)) for f in self.files}
        n = len(self.files)
        keep = lambda fs: [f for f in fs if max_length <= len(self._data[f])]   # MT/data.py:33-40
        self.file_dict = {'train': keep(self.files[:int(n * 0.8)]),
                          'valid': keep(self.files[int(n * 0.8):int(n * 0.9)]),
                          'test': keep(self.files[int(n * 0.9):])}
        self._seq_file_name_idx = 0
        self._seq_idx = 0

    def _get_seq(self, fname, max_length=None):                             # MT/data.py:96-107
        import random as _random
        data = self._data[fname]
        if max_length is not None:
            if max_length <= len(data):
                start = _random.randrange(0, len(data) - max_length)
                data = data[start:start + max_length]
            else:
                raise IndexError
        return data

    def batch(self, batch_size, length, mode='train'):                      # MT/data.py:41-48
        import random as _random
        batch_files = _random.sample(self.file_dict[mode], k=batch_size)
        return np.array([self._get_seq(f, length) for f in batch_files], dtype=np.int16)

    def seq2seq_batch(self, batch_size, length, mode='train'):              # MT/data.py:50-54
        data = self.batch(batch_size, length * 2, mode)
        return data[:, :length], data[:, length:]

    def smallest_encoder_batch(self, batch_size, length, mode='train'):     # MT/data.py:56-60
        data = self.batch(batch_size, length * 2, mode)
        return data[:, :length // 100], data[:, length // 100:length // 100 + length]

    def slide_seq2seq_batch(self, batch_size, length, mode='train'):        # MT/data.py:62-66
        data = self.batch(batch_size, length + 1, mode)
        return data[:, :-1], data[:, 1:]

    def random_sequential_batch(self, batch_size, length):                  # MT/data.py:68-76
        import random as _random
        batch_files = _random.sample(self.files, k=batch_size)
        out = []
        for i in range(batch_size):
            data = self._get_seq(batch_files[i])
            for j in range(len(data) - length):
                out.append(data[j:j + length])
                if len(out) == batch_size:
                    return out
        return None

    def sequential_batch(self, batch_size, length):                         # MT/data.py:78-94
        out = []
        data = self._get_seq(self.files[self._seq_file_name_idx])
        while len(out) < batch_size:
            while self._seq_idx < len(data) - length:
                out.append(data[self._seq_idx:self._seq_idx + length])
                self._seq_idx += 1
                if len(out) == batch_size:
                    return out
            self._seq_idx = 0
            self._seq_file_name_idx += 1
            if self._seq_file_name_idx == len(self.files):
                self._seq_file_name_idx = 0
