"""Mask helpers of the path (MT/utils.py:58-83, :183-188) in structured form.

The reference materialises a bool [B,1,L,L] look-ahead mask and later re-expands it to int64
and fp32 (MT/layers.py:100).  The kernels only need the causal predicate and one pad bit per
key, so ``get_masked_with_pad_tensor`` returns an ``engine.Mask``; call ``materialize(mask)``
for the reference's dense tensor."""
from __future__ import annotations

import torch

from .engine import Mask


def sequence_mask(length: torch.Tensor, max_length=None) -> torch.Tensor:
    """TensorFlow-style sequence mask: out[r, c] = c < length[r]."""
    if max_length is None:
        max_length = int(length.max())
    cols = torch.arange(max_length, dtype=length.dtype, device=length.device)
    return cols[None, :] < length[:, None]


def pad_key_bits(x: torch.Tensor, pad_token: int):
    """uint8 [B, L] (1 = key is a pad token) or None when the batch has no pad token."""
    bits = (x == pad_token)
    return bits.to(torch.uint8).contiguous()


def get_masked_with_pad_tensor(size, src, trg, pad_token):
    """Returns (src_mask, trg_mask, look_ahead_mask) like the reference; the first two are the
    Python bools the reference computes with ``torch.equal`` (callers ignore them), the third is
    the structured look-ahead mask  (trg[b,j] == pad) | (j > i)."""
    if trg is None:
        return None, None, None
    if trg.size(1) != size:
        # same failure mode as the reference's broadcast of [B,1,1,L] against [size,size]
        raise RuntimeError(f"The size of tensor a ({trg.size(1)}) must match the size of tensor b "
                           f"({size}) at non-singleton dimension 3")
    return False, False, Mask(True, pad_key_bits(trg, pad_token))


def materialize(mask: Mask, L: int) -> torch.Tensor:
    """Dense bool [B,1,L,L] (True = masked) of a structured mask."""
    dev = mask.pad_keys.device if mask.pad_keys is not None else None
    ar = torch.arange(L, device=dev)
    m = (ar[None, :] > ar[:, None]) if mask.causal else torch.zeros(L, L, dtype=torch.bool, device=dev)
    m = m[None, None]
    if mask.pad_keys is not None:
        m = m | mask.pad_keys.bool()[:, None, None, :]
    return m


class ScalarReadback:
    """Device -> host read of per-step scalars (loss, accuracy) without stalling the launch queue.

    MT/train.py:267-283 reads ``metrics['loss']`` with an implicit synchronise every iteration, which
    drains the GPU before the host starts enqueueing the next step.  Here ``push(t)`` enqueues an async
    copy of the 0-d tensor into a pinned slot and records an event; ``pop()`` returns the OLDEST pushed
    value as a Python float, waiting only for that copy's event.  Calling ``pop()`` once ``depth - 1``
    steps later (depth 2: right after the next step has been enqueued) keeps every step's value read on
    the host while the device never idles; ``drain()`` returns what is still in flight."""

    def __init__(self, depth: int = 2, dtype=torch.float32):
        if not torch.cuda.is_available():
            raise RuntimeError("ScalarReadback needs a CUDA device")
        self.depth = max(1, depth)
        self._slots = torch.empty(self.depth, dtype=dtype).pin_memory()
        self._events = [torch.cuda.Event() for _ in range(self.depth)]
        self._head = 0      # next slot to fill
        self._n = 0         # values in flight

    def push(self, t: torch.Tensor):
        if self._n == self.depth:
            raise RuntimeError("ScalarReadback: pop() before pushing more than `depth` values")
        i = self._head
        self._slots[i:i + 1].copy_(t.detach().reshape(1), non_blocking=True)
        self._events[i].record()
        self._head = (i + 1) % self.depth
        self._n += 1

    def __len__(self):
        return self._n

    def pop(self) -> float:
        if self._n == 0:
            raise RuntimeError("ScalarReadback: nothing in flight")
        i = (self._head - self._n) % self.depth
        self._events[i].synchronize()
        self._n -= 1
        return float(self._slots[i])

    def drain(self):
        return [self.pop() for _ in range(self._n)]
