"""One warm-up + one measured call of the tcgen05 attention forward/backward at the BASELINE
config-B per-layer shape (B=16, h=8, L=2048, dh=64, bf16).  Used under ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musicgeneration_b200 import ops

B, h, L, dh, max_seq = int(os.environ.get("PB", 16)), 8, 2048, 64, 2048
dev = torch.device("cuda:0")
d = h * dh
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, L, 3, h, dh, generator=g).to(torch.bfloat16).to(dev)
E = torch.randn(max_seq, dh, generator=g).to(torch.bfloat16).to(dev)
dO = torch.randn(B, L, h, dh, generator=g).to(torch.bfloat16).to(dev)
strides, ostr = (L * 3 * d, 3 * d, dh), (L * d, d, dh)
Od = torch.empty(B, L, h, dh, dtype=torch.bfloat16, device=dev)
lse = torch.empty(B, h, L, device=dev)
dqkv = torch.zeros(B, L, 3, h, dh, dtype=torch.bfloat16, device=dev)
dE = torch.zeros(max_seq, dh, device=dev)
delta = torch.empty(B, h, L, device=dev)
SPILL = os.environ.get("SPILL", "1") == "1"
STASH = os.environ.get("STASH", "0") == "1"      # the training pair: the forward keeps its P tiles, the backward reads them
st = ops.rga_stash_new(qkv[:, :, 0], E, Od, B, h, L, dh) if STASH else None
ITERS = int(os.environ.get("PITER", 3))
hist = []
for it in range(ITERS):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    ops.rga_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, ostr, lse, B, h, L, dh, max_seq, True, path=2, stash=st)
    e[1].record()
    ops.rga_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, dO, ostr, lse, delta,
                dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], dE, B, h, L, dh, max_seq, True, path=2, spill=SPILL, stash=st)
    e[2].record()
    torch.cuda.synchronize()
    hist.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    if ITERS <= 3:
        print("fwd ms", hist[-1][0], "bwd ms", hist[-1][1])
if ITERS > 3:
    f = sorted(h[0] for h in hist[2:])
    b = sorted(h[1] for h in hist[2:])
    print(f"fwd ms median {f[len(f) // 2]:.4f} min {f[0]:.4f}   bwd ms median {b[len(b) // 2]:.4f} min {b[0]:.4f}")
