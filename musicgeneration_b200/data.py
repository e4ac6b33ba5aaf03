"""Data feed: the reference's ``Data`` class (MT/data.py:10-107) over an HBM-resident token arena.

Upstream every batch re-reads ``batch_size`` pickled files from disk (``torch.load`` in ``_get_seq``,
MT/data.py:96-97), slices a window on the host, stacks a numpy int16 array and converts it to a
device int32 tensor (MT/train.py:258-260).  Here all ``.data`` files (``torch.save`` of a 1-D numpy
uint8 / uint16 array, REF/mg/model/utils/preprocess_MIDI_like.py:21-41) are read ONCE, concatenated
into one unsigned arena on the device, and a batch is one ``mt_window_gather`` launch
(csrc/feed.cu) producing int32 ``x`` / ``y`` in place.

Same surface: ``Data(dir_path, max_length)``, ``file_dict`` with the 80/10/10 split in ``os.walk``
order, ``batch``, ``seq2seq_batch``, ``smallest_encoder_batch``, ``slide_seq2seq_batch``,
``random_sequential_batch``, ``sequential_batch`` and the error behaviour of ``_get_seq`` (IndexError
for a file shorter than the window -- MT/train.py:261 catches it and skips the batch -- and the
ValueError of ``random.randrange(0, 0)`` for a file of exactly the window length).  The host draws
files and window starts with the SAME ``random`` call sequence (``random.sample`` then one
``random.randrange`` per file), so under an equal ``random.seed`` the batches are bit-identical to
the reference's.  The numpy-returning methods copy the gathered windows back (API compatibility);
``*_device`` variants return the int32 CUDA tensors the train loop wants, and ``device_sampler=True``
draws on the device as well (no host->device traffic at all; different random stream).
"""
from __future__ import annotations

import os
import random
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import ops


def find_files_by_extensions(root, exts: Sequence[str] = ()):
    """MT/utils.py:10-22 -- os.walk order, case-insensitive suffix match."""
    for path, _, files in os.walk(root):
        for name in files:
            if not exts or any(name.lower().endswith(e) for e in exts):
                yield os.path.join(path, name)


def _load_tokens(fname: str) -> np.ndarray:
    data = torch.load(fname, weights_only=False)       # a pickled numpy array (torch >= 2.6 needs the flag)
    arr = np.asarray(data)
    if arr.ndim != 1:
        raise ValueError(f"{fname}: expected a 1-D token array, got shape {arr.shape}")
    if arr.size and (arr.min() < 0 or arr.max() > 0xFFFF):
        raise ValueError(f"{fname}: token ids outside uint16")
    return arr


class Data:
    def __init__(self, dir_path, max_length, device=None, seed: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("musicgeneration_b200.data.Data keeps its token arena in HBM: CUDA device required "
                               "(no CPU fallback)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.files: List[str] = list(find_files_by_extensions(dir_path, ['.data']))
        arrays = [_load_tokens(f) for f in self.files]
        self._len = {f: int(a.size) for f, a in zip(self.files, arrays)}
        n = len(self.files)
        self.file_dict = {
            'train': self.file_filter(self.files[:int(n * 0.8)], max_length),
            'valid': self.file_filter(self.files[int(n * 0.8): int(n * 0.9)], max_length),
            'test': self.file_filter(self.files[int(n * 0.9):], max_length),
        }
        self._seq_file_name_idx = 0
        self._seq_idx = 0
        # --- arena ------------------------------------------------------------------------
        off = np.zeros(n + 1, dtype=np.int64)
        if n:
            off[1:] = np.cumsum([a.size for a in arrays])
        self._off_host = off
        self._index = {f: i for i, f in enumerate(self.files)}
        wide = any(a.size and a.max() > 0xFF for a in arrays)
        np_dt, th_dt = (np.uint16, torch.uint16) if wide else (np.uint8, torch.uint8)
        host = torch.empty(max(int(off[-1]), 1), dtype=th_dt).pin_memory()
        flat = host.numpy()
        for a, o in zip(arrays, off[:-1]):
            flat[o:o + a.size] = a.astype(np_dt, copy=False)
        self.arena = host.to(self.device, non_blocking=True)
        self.file_off = torch.from_numpy(off).to(self.device)
        self._starts_host = None
        self._starts_slot = 0
        # the device sampler is keyed by (seed, step): mix the data-parallel rank in, so that the replicas of a
        # one-process-per-GPU job draw DIFFERENT files and windows under the default seed (the host-drawn path
        # follows Python's `random`, which the caller seeds per rank exactly as with the reference)
        self._sampler_seed = (int(seed) + 0x9E3779B97F4A7C15 * int(os.environ.get("RANK", "0"))) & 0xFFFFFFFFFFFFFFFF
        self._sampler_step = 0
        self._eligible = {}
        torch.cuda.current_stream(self.device).synchronize()      # the pinned staging buffer may go now

    def __repr__(self):
        return (f"<class Data has train: {len(self.file_dict['train'])}, val: {len(self.file_dict['valid'])},"
                f"test: {len(self.file_dict['test'])} files>")

    def file_filter(self, files, max_length):
        return [f for f in files if max_length <= self._len[f]]

    # --- window drawing (host, reference random stream) ------------------------------------
    def _draw(self, batch_size: int, length, mode: str) -> Tuple[np.ndarray, List[str]]:
        """random.sample + per-file random.randrange in the order MT/data.py:42-47,98-101 makes them."""
        batch_files = random.sample(self.file_dict[mode], k=batch_size)
        starts = np.empty(batch_size, dtype=np.int64)
        for i, f in enumerate(batch_files):
            n = self._len[f]
            if length <= n:
                starts[i] = self._off_host[self._index[f]] + random.randrange(0, n - length)
            else:
                raise IndexError
        return starts, batch_files

    _RING = 4       # pinned staging slots for the window starts

    def _starts_to_device(self, starts: np.ndarray) -> torch.Tensor:
        """Host window starts -> device.  The copy is asynchronous and the train loop lets the host run ahead of
        the GPU, so ONE pinned buffer could be refilled with the next batch's starts before the queued copy of the
        previous batch had run (batch N would train on batch N+1's windows).  A ring of pinned slots, each guarded
        by an event recorded after its copy, closes that: a slot is rewritten only once its last copy has run."""
        B = starts.size
        ring = self._starts_host
        if ring is None or ring[0][0].numel() < B:
            if ring is not None:
                for _, ev in ring:
                    if ev is not None:
                        ev.synchronize()
            ring = self._starts_host = [[torch.empty(max(B, 64), dtype=torch.int64).pin_memory(), None]
                                        for _ in range(self._RING)]
            self._starts_slot = 0
        slot = ring[self._starts_slot]
        self._starts_slot = (self._starts_slot + 1) % self._RING
        if slot[1] is not None:
            slot[1].synchronize()                      # the copy that last read this slot has completed
        slot[0][:B].copy_(torch.from_numpy(starts))
        with torch.cuda.device(self.device):
            out = slot[0][:B].to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        slot[1] = ev
        return out

    def _gather(self, starts_dev: torch.Tensor, x_len: int, y_len: int = 0, y_shift: int = 0):
        B = starts_dev.numel()
        with torch.cuda.device(self.device):
            x = torch.empty((B, x_len), dtype=torch.int32, device=self.device)
            y = torch.empty((B, y_len), dtype=torch.int32, device=self.device) if y_len else None
            ops.window_gather(self.arena, starts_dev, x, y, y_shift)
        return x, y

    def _device_draw(self, batch_size: int, need: int, mode: str) -> torch.Tensor:
        key = (mode, need)
        el = self._eligible.get(key)
        if el is None:
            idx = [self._index[f] for f in self.file_dict[mode] if self._len[f] > need]
            el = torch.tensor(idx, dtype=torch.int64, device=self.device)
            self._eligible[key] = el
        if el.numel() < batch_size:
            raise ValueError("Sample larger than population or is negative")
        starts = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            ops.window_sample(self.file_off, el, need, self._sampler_seed, self._sampler_step, starts)
        self._sampler_step += 1
        return starts

    # --- device-resident batches (what the train loop consumes) ----------------------------
    def batch_device(self, batch_size, length, mode='train', device_sampler=False) -> torch.Tensor:
        starts = (self._device_draw(batch_size, length, mode) if device_sampler
                  else self._starts_to_device(self._draw(batch_size, length, mode)[0]))
        return self._gather(starts, length)[0]

    def slide_seq2seq_batch_device(self, batch_size, length, mode='train', device_sampler=False):
        """x = w[:, :-1], y = w[:, 1:] of a (length+1)-token window: both written by one launch."""
        starts = (self._device_draw(batch_size, length + 1, mode) if device_sampler
                  else self._starts_to_device(self._draw(batch_size, length + 1, mode)[0]))
        return self._gather(starts, length, length, 1)

    def seq2seq_batch_device(self, batch_size, length, mode='train', device_sampler=False):
        starts = (self._device_draw(batch_size, length * 2, mode) if device_sampler
                  else self._starts_to_device(self._draw(batch_size, length * 2, mode)[0]))
        return self._gather(starts, length, length, length)

    # --- reference surface (numpy int16 on the host, MT/data.py:41-67) ----------------------
    @staticmethod
    def _np(t: torch.Tensor) -> np.ndarray:
        return t.cpu().numpy().astype(np.int16)

    def batch(self, batch_size, length, mode='train'):
        return self._np(self.batch_device(batch_size, length, mode))

    def seq2seq_batch(self, batch_size, length, mode='train'):
        x, y = self.seq2seq_batch_device(batch_size, length, mode)
        return self._np(x), self._np(y)

    def smallest_encoder_batch(self, batch_size, length, mode='train'):
        starts = self._starts_to_device(self._draw(batch_size, length * 2, mode)[0])
        x, y = self._gather(starts, length // 100, length, length // 100) if length // 100 else \
            (torch.empty((batch_size, 0), dtype=torch.int32, device=self.device),
             self._gather(starts, length)[0])
        return self._np(x), self._np(y)

    def slide_seq2seq_batch(self, batch_size, length, mode='train'):
        x, y = self.slide_seq2seq_batch_device(batch_size, length, mode)
        return self._np(x), self._np(y)

    def _windows(self, pairs: List[Tuple[int, int]], length: int):
        """pairs = (file index, start within the file) -> list of 1-D host arrays, one launch."""
        if not pairs:
            return []
        starts = np.array([self._off_host[f] + s for f, s in pairs], dtype=np.int64)
        x, _ = self._gather(self._starts_to_device(starts), length)
        dt = np.uint16 if self.arena.dtype == torch.uint16 else np.uint8
        return list(x.cpu().numpy().astype(dt))

    def random_sequential_batch(self, batch_size, length):
        """MT/data.py:69-77: consecutive windows of the sampled files until the batch is full (None if the
        sampled files cannot fill it, as upstream falls off the end of its loop)."""
        batch_files = random.sample(self.files, k=batch_size)
        pairs = []
        for f in batch_files:
            for j in range(self._len[f] - length):
                pairs.append((self._index[f], j))
                if len(pairs) == batch_size:
                    return self._windows(pairs, length)
        return None

    def sequential_batch(self, batch_size, length):
        """MT/data.py:79-94: a cursor (file, offset) sliding one token at a time through the file list."""
        pairs = []
        n = self._len[self.files[self._seq_file_name_idx]]
        fi = self._seq_file_name_idx       # upstream keeps reading the file it loaded at entry
        while len(pairs) < batch_size:
            while self._seq_idx < n - length:
                pairs.append((fi, self._seq_idx))
                self._seq_idx += 1
                if len(pairs) == batch_size:
                    return self._windows(pairs, length)
            self._seq_idx = 0
            self._seq_file_name_idx = self._seq_file_name_idx + 1
            if self._seq_file_name_idx == len(self.files):
                self._seq_file_name_idx = 0
                print('iter intialized')

    def _get_seq(self, fname, max_length=None):
        """Host view of one file (or of a random window of it) -- kept for callers that poke at it."""
        i = self._index[fname]
        lo, hi = int(self._off_host[i]), int(self._off_host[i + 1])
        if max_length is not None:
            if max_length <= hi - lo:
                start = random.randrange(0, hi - lo - max_length)
                lo, hi = lo + start, lo + start + max_length
            else:
                raise IndexError
        return self.arena[lo:hi].cpu().numpy()
