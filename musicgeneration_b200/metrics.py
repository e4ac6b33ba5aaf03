"""Step metrics of MT/metrics.py:40-75 with the same call surface.  When fed the logits the
criterion just consumed they could reuse ``SmoothCrossEntropyLoss.last_argmax``; standalone
they run the CE kernel's arg-max pass themselves (no softmax pass: argmax(softmax(z)) ==
argmax(z))."""
from __future__ import annotations

from typing import Dict

import torch

from . import ops


def _argmax(logits: torch.Tensor) -> torch.Tensor:
    if not logits.is_cuda:
        raise RuntimeError("musicgeneration_b200 runs on CUDA tensors only (no CPU fallback)")
    V = logits.shape[-1]
    z = logits.detach().reshape(-1, V).float().contiguous()
    T = z.shape[0]
    t = torch.zeros((T,), dtype=torch.int32, device=z.device)
    row_ws = torch.empty((3, T), dtype=torch.float32, device=z.device)
    am = torch.empty((T,), dtype=torch.int32, device=z.device)
    sums = torch.empty((4,), dtype=torch.float32, device=z.device)
    ops.smooth_ce_fwd(z, t, row_ws, am, sums, 0.0, -1)
    return am


class _Metric(torch.nn.Module):
    def forward(self, input: torch.Tensor, target: torch.Tensor):
        raise NotImplementedError()


class Accuracy(_Metric):
    def forward(self, input: torch.Tensor, target: torch.Tensor):
        """input [B, T] predicted ids, target [B, T]: fraction of equal positions (pads included)."""
        hit = (input.reshape(-1) == target.reshape(-1).to(input.dtype))
        return hit.to(torch.float32).mean()


class MockAccuracy(Accuracy):
    pass


class CategoricalAccuracy(Accuracy):
    def forward(self, input: torch.Tensor, target: torch.Tensor):
        """input [B, T, V] logits."""
        return super().forward(_argmax(input), target)


class LogitsBucketting(_Metric):
    def __init__(self, vocab_size):
        super().__init__()

    def forward(self, input: torch.Tensor, target: torch.Tensor):
        return _argmax(input)


class MetricsSet(object):
    def __init__(self, metric_dict: Dict):
        super().__init__()
        self.metrics = metric_dict

    def __call__(self, input: torch.Tensor, target: torch.Tensor):
        return self.forward(input=input, target=target)

    def forward(self, input: torch.Tensor, target: torch.Tensor):
        return {k: metric(input.to(target.device), target) for k, metric in self.metrics.items()}
