"""Times the GEMM shapes of BASELINE config B (T = 16 x 2048 tokens, d = 512) through the C ABI:
forward (x.W^T), dgrad (dy.W) and wgrad (dy^T.x), bf16 operands."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musicgeneration_b200 import ops

dev = torch.device("cuda:0")
T, d, V = 32768, 512, 390
bf = torch.bfloat16
shapes = [("qkv", 3 * d, d), ("fc", d, d), ("pre", d // 2, d), ("suf", d, d // 2), ("vocab", V, d)]
g = torch.Generator().manual_seed(0)
res = []
for name, N, K in shapes:
    x = torch.randn(T, K, generator=g).to(bf).to(dev)
    W = torch.randn(N, K, generator=g).to(bf).to(dev)
    ldn = (N + 7) // 8 * 8
    dy = torch.randn(T, ldn, generator=g).to(bf).to(dev)
    b = torch.randn(N).to(dev)
    for kind in ("fwd_bf16", "fwd_f32", "dgrad", "dgrad_add", "dgrad_mask", "wgrad"):
        if kind.startswith("fwd"):
            out = torch.empty(T, N, dtype=bf if kind == "fwd_bf16" else torch.float32, device=dev)
            fn = lambda: ops.gemm(x, W, out, T, N, K, K, K, N, False, True, bias=b)
            byts = x.numel() * 2 + W.numel() * 2 + out.numel() * out.element_size()
        elif kind == "dgrad":
            out = torch.empty(T, K, dtype=torch.float32, device=dev)
            fn = lambda: ops.gemm(dy, W, out, T, K, N, ldn, K, K, False, False)
            byts = T * N * 2 + W.numel() * 2 + out.numel() * 4
        elif kind == "dgrad_add":       # residual-path gradient added in place (fp32), as in engine.layer_bwd
            out = torch.zeros(T, K, dtype=torch.float32, device=dev)
            fn = lambda: ops.gemm(dy, W, out, T, K, N, ldn, K, K, False, False, addend=out)
            byts = T * N * 2 + W.numel() * 2 + out.numel() * 8
        elif kind == "dgrad_mask":      # bf16 output masked by the saved ReLU activations
            out = torch.empty(T, K, dtype=bf, device=dev)
            aux = torch.randn(T, K, generator=g).to(bf).to(dev)
            fn = lambda: ops.gemm(dy, W, out, T, K, N, ldn, K, K, False, False, aux=aux, relu_mask=True)
            byts = T * N * 2 + W.numel() * 2 + out.numel() * 4
        else:
            out = torch.empty(N, K, dtype=torch.float32, device=dev)
            fn = lambda: ops.gemm(dy, x, out, N, K, T, ldn, K, K, True, False)
            byts = T * N * 2 + x.numel() * 2 + out.numel() * 4
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        fl = 2.0 * T * N * K
        print(f"{name:6s} {kind:9s} N={N:5d} K={K:4d}  {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s  {byts / us / 1e3:7.1f} GB/s")
