"""Times the residual+LayerNorm forward / backward kernels at config B's shape (T = 32768, d = 512) with
and without dropout, bf16 and fp32 sublayer input, rotating over enough buffers to defeat the L2
(each set is ~300 MB).  Prints us per launch and GB/s of algorithmic bytes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musicgeneration_b200 import ops

dev = torch.device("cuda:0")
T, d = 32768, 512
NSET = 3
g = torch.ones(d, device=dev)
b = torch.zeros(d, device=dev)
for adt in (torch.bfloat16, torch.float32):
    sets = []
    for _ in range(NSET):
        sets.append(dict(a=torch.randn(T, d, device=dev).to(adt), x=torch.randn(T, d, device=dev),
                         out=torch.empty(T, d, device=dev), lp=torch.empty(T, d, device=dev, dtype=torch.bfloat16),
                         mean=torch.empty(T, device=dev), rstd=torch.empty(T, device=dev),
                         dout=torch.randn(T, d, device=dev), da=torch.empty(T, d, device=dev, dtype=torch.bfloat16),
                         dg=torch.empty(d, device=dev), db=torch.empty(d, device=dev), dbias=torch.empty(d, device=dev)))
    for p in (0.0, 0.2):
        def fwd(s):
            ops.add_ln_fwd(s["a"], s["x"], g, b, s["out"], s["lp"], s["mean"], s["rstd"], 1e-6, p, 1, 2)

        def bwd(s):
            ops.add_ln_bwd(s["dout"], s["a"], s["x"], g, s["mean"], s["rstd"], s["dout"], s["da"], s["dg"], s["db"],
                           p, 1, 2, dbias=s["dbias"])
        es = adt.itemsize if hasattr(adt, "itemsize") else (2 if adt == torch.bfloat16 else 4)
        for name, fn, byts in (("fwd", fwd, T * d * (es + 4 + 4 + 2)), ("bwd", bwd, T * d * (es + 4 + 4 + 4 + 2))):
            for s in sets:
                fn(s)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(30):
                fn(sets[i % NSET])
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 30 * 1e3
            print(f"add_ln_{name} a={str(adt)[6:]:8s} p={p}: {us:6.1f} us  {byts / us / 1e3:7.1f} GB/s")
