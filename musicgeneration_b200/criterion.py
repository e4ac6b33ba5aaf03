"""Host mirror of MT/criterion.py: SmoothCrossEntropyLoss (:28-67) on the fused CE kernel and
the Noam schedule wrapper CustomSchedule (:70-96)."""
from __future__ import annotations

import torch
from torch.nn.modules.loss import _Loss

from . import ops


class _SmoothCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, eps, vocab, ignore):
        if not logits.is_cuda:
            raise RuntimeError("musicgeneration_b200 runs on CUDA tensors only (no CPU fallback)")
        V = logits.shape[-1]
        if V != vocab:
            raise RuntimeError(f"logits have {V} classes, criterion was built for {vocab}")
        z = logits.reshape(-1, V)
        if z.dtype != torch.float32:
            z = z.float()
        z = z.contiguous()
        t = target.reshape(-1).to(torch.int32).contiguous()
        T = z.shape[0]
        row_ws = torch.empty((3, T), dtype=torch.float32, device=z.device)
        argmax = torch.empty((T,), dtype=torch.int32, device=z.device)
        sums = torch.empty((4,), dtype=torch.float32, device=z.device)
        ops.smooth_ce_fwd(z, t, row_ws, argmax, sums, eps, ignore)
        ctx.save_for_backward(z, t, row_ws, sums)
        ctx.eps, ctx.ignore, ctx.shape = eps, ignore, logits.shape
        ctx.mark_non_differentiable(argmax, sums)
        return sums[3], sums[0], argmax, sums

    @staticmethod
    def backward(ctx, g_mean, g_sum, _ga, _gs):
        z, t, row_ws, sums = ctx.saved_tensors
        dz = torch.empty_like(z)
        # d(mean)/dz = (softmax - q')/n_valid ; d(sum)/dz = n_valid times that
        g = None
        if g_mean is not None and g_sum is not None:
            g = (g_mean + g_sum * sums[1]).reshape(1).float().contiguous()
        elif g_mean is not None:
            g = g_mean.reshape(1).float().contiguous()
        elif g_sum is not None:
            g = (g_sum * sums[1]).reshape(1).float().contiguous()
        ops.smooth_ce_bwd(z, t, row_ws, sums, g, dz, ctx.eps, ctx.ignore)
        return dz.view(ctx.shape), None, None, None, None


class SmoothCrossEntropyLoss(_Loss):
    """Label-smoothed cross entropy (arXiv:1512.00567) with ``ignore_index`` rows dropped from both
    the sum and the mean's denominator.  After a call, ``last_argmax`` (int32 [T]) and
    ``last_sums`` ([loss_sum, n_valid, n_correct, mean]) hold the step metrics the same kernel
    pass produced (MT/metrics.py:50-60)."""
    __constants__ = ['label_smoothing', 'vocab_size', 'ignore_index', 'reduction']

    def __init__(self, label_smoothing, vocab_size, ignore_index=-100, reduction='mean', is_logits=True):
        assert 0.0 <= label_smoothing <= 1.0
        super().__init__(reduction=reduction)
        self.label_smoothing = label_smoothing
        self.vocab_size = vocab_size
        self.ignore_index = ignore_index
        self.input_is_logits = is_logits
        self.last_argmax = None
        self.last_sums = None

    def forward(self, input, target):
        mean, total, argmax, sums = _SmoothCEFunction.apply(
            input, target, float(self.label_smoothing), int(self.vocab_size), int(self.ignore_index))
        self.last_argmax, self.last_sums = argmax, sums
        if self.reduction == 'mean':
            return mean
        elif self.reduction == 'sum':
            return total
        raise NotImplementedError


class CustomSchedule:
    """lr = d_model^-0.5 * min(step^-0.5, step * warmup^-1.5), applied to every param group
    before ``optimizer.step()``."""

    def __init__(self, d_model, warmup_steps=4000, optimizer=None):
        self.d_model = d_model
        self.optimizer = optimizer
        self.warmup_steps = warmup_steps
        self._step = 0
        self._rate = 0

    def step(self):
        self._step += 1
        rate = self.rate()
        for group in self.optimizer.param_groups:
            group['lr'] = rate
        self._rate = rate
        self.optimizer.step()

    def rate(self, step=None):
        if step is None:
            step = self._step
        return self.d_model ** (-0.5) * min(step ** (-0.5), step * (self.warmup_steps ** -1.5))
