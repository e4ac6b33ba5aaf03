"""Time one batch of the data feed (16 x 2048 window pairs, config B) three ways on the same corpus:
  * arena + mt_window_gather, window starts drawn on the host (reference random stream)
  * arena + mt_window_sample + mt_window_gather (nothing crosses PCIe)
  * the reference's way (MT/data.py:41-67 + MT/train.py:258-260): torch.load of 16 files, numpy stack,
    two pageable H2D copies with an int16 -> int32 conversion on the device  (host port, timed beside it)
Wall-clock per batch including a final device synchronize; medians of N iterations."""
import os
import random
import statistics
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from musicgeneration_b200 import data as mdata  # noqa: E402

N = int(os.environ.get("N", "200"))
B, L = 16, 2048


def ref_batch(files, lens):
    fs = random.sample(files, k=B)
    rows = []
    for f in fs:
        d = torch.load(f, weights_only=False)
        s = random.randrange(0, len(d) - (L + 1))
        rows.append(d[s:s + L + 1])
    w = np.array(rows, dtype=np.int16)
    x = torch.from_numpy(w[:, :-1]).contiguous().to("cuda", non_blocking=True, dtype=torch.int)
    y = torch.from_numpy(w[:, 1:]).contiguous().to("cuda", non_blocking=True, dtype=torch.int)
    return x, y


def timed(fn):
    ts = []
    for _ in range(N):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    return statistics.median(ts)


with tempfile.TemporaryDirectory() as tmp:
    rng = np.random.RandomState(0)
    for i in range(256):
        torch.save(rng.randint(0, 388, size=int(rng.randint(3000, 20000))).astype(np.uint16),
                   os.path.join(tmp, f"p{i:04d}.data"))
    D = mdata.Data(tmp, L + 1)
    lens = D._len
    for _ in range(5):
        D.slide_seq2seq_batch_device(B, L)
        D.slide_seq2seq_batch_device(B, L, device_sampler=True)
        ref_batch(D.file_dict['train'], lens)
    a = timed(lambda: D.slide_seq2seq_batch_device(B, L))
    b = timed(lambda: D.slide_seq2seq_batch_device(B, L, device_sampler=True))
    c = timed(lambda: ref_batch(D.file_dict['train'], lens))
    # kernel alone (CUDA events on the current stream)
    st = torch.randint(0, 1000, (B,), dtype=torch.int64, device="cuda")
    x = torch.empty(B, L, dtype=torch.int32, device="cuda")
    y = torch.empty(B, L, dtype=torch.int32, device="cuda")
    from musicgeneration_b200 import ops
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        ops.window_gather(D.arena, st, x, y, 1)
    e1.record()
    torch.cuda.synchronize()
    k = e0.elapsed_time(e1) * 10
    print(f"feed 16x2048: arena host-drawn {a:.1f} us  arena device-drawn {b:.1f} us  reference-style {c:.1f} us  "
          f"gather kernel {k:.2f} us/launch back to back ({(B * (L + 1) * 2 + 2 * B * L * 4) / k / 1e3:.1f} GB/s)")
