import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musicgeneration_b200 import ops
dev = torch.device("cuda:0")
B, h, dh, ms, d = 32, 8, 64, 2048, 512
bf = torch.bfloat16
kcs = [torch.randn(B, h, ms, dh, device=dev).to(bf) for _ in range(6)]
vcs = [torch.randn(B, h, ms, dh, device=dev).to(bf) for _ in range(6)]
E = torch.randn(ms, dh, device=dev).to(bf)
qkv = torch.randn(B, 3 * d, device=dev).to(bf)
o = torch.empty(B, d, device=dev, dtype=bf)
pad = torch.zeros(B, ms, dtype=torch.uint8, device=dev)
ws = ops.decode_attend_workspace(B, h, dh, ms, dev)
for t in (128, 512, 1000, 1664, 2040):
    td = torch.tensor([t], dtype=torch.int32, device=dev)
    for mode in ("host_t", "dev_t"):
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(10):
                for l in range(6):
                    if mode == "host_t":
                        ops.rga_decode(qkv, 3 * d, kcs[l], vcs[l], E, pad, o, B, h, dh, ms, t)
                    else:
                        ops.decode_attend(qkv, 3 * d, kcs[l], vcs[l], E, pad, o, td, B, h, dh, ms, ws)
            e1.record(); torch.cuda.synchronize()
        print(t, mode, "us per call", e0.elapsed_time(e1) * 1000 / 60)
    o1 = o.clone(); ops.rga_decode(qkv, 3 * d, kcs[0], vcs[0], E, pad, o1, B, h, dh, ms, t); o2 = o.clone(); ops.decode_attend(qkv, 3 * d, kcs[0], vcs[0], E, pad, o2, td, B, h, dh, ms, ws); print("  max |split - single|", float((o1.float() - o2.float()).abs().max()), "max |o|", float(o1.float().abs().max()))
