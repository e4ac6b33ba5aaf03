import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musicgeneration_b200 import ops
dev = torch.device("cuda:0")
T, N, K = 32768, 1536, 512
bf = torch.bfloat16
x = torch.randn(T, K).to(bf).to(dev); W = torch.randn(N, K).to(bf).to(dev); b = torch.randn(N).to(dev)
out = torch.empty(T, N, dtype=bf, device=dev)
for _ in range(3):
    ops.gemm(x, W, out, T, N, K, K, K, N, False, True, bias=b)
torch.cuda.synchronize()
