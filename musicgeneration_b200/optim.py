"""Optimizer side of the train step ("next" row 1 of SURVEY 8f): Adam with the reference's
hyper-parameters (MT/train.py:143: betas (0.9, 0.98), eps 1e-9, lr driven by the Noam
schedule of MT/criterion.py:70-96) as ONE fused kernel over a flat fp32 parameter buffer, plus
the data-parallel gradient exchange: one NCCL all-reduce over the flat gradient buffer.

``FlatAdam`` re-points every ``param.data`` / ``param.grad`` at views of two flat buffers (the
state_dict keys and shapes are untouched; Wq/Wk/Wv stay adjacent so the fused QKV GEMM needs no
repacking)."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import ops


def _ordered(params: List[torch.nn.Parameter]) -> List[torch.nn.Parameter]:
    return list(params)


# The optimizer whose 16-bit weight shadow has been refreshed for the forward pass in progress
# (MusicTransformer.forward sets / clears it).  While set, the layers take their 16-bit weight
# operands as views of the shadow instead of casting ~30 matrices one launch each.
_ACTIVE_SHADOW = [None]


def weight_shadow(t: torch.Tensor, act: torch.dtype) -> Optional[torch.Tensor]:
    """The 16-bit copy of fp32 weight tensor ``t`` inside the active shadow, or None."""
    opt = _ACTIVE_SHADOW[0]
    if opt is None or opt.flat_lp is None or opt.flat_lp.dtype != act or t.dtype != torch.float32 \
            or not t.is_contiguous():
        return None
    off = t.data_ptr() - opt.flat_p.data_ptr()
    if off < 0 or off + t.numel() * 4 > opt.n * 4 or off % 4:
        return None
    off //= 4
    return opt.flat_lp[off:off + t.numel()].view(t.shape)


class FlatAdam:
    def __init__(self, model: torch.nn.Module, lr: float = 0.0, betas=(0.9, 0.98), eps: float = 1e-9,
                 process_group=None, grad_accum: int = 1):
        from .layers import RelativeGlobalAttention
        seen, order = set(), []

        def add(p):
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                order.append(p)

        # keep Wq/Wk/Wv weights (and biases) adjacent: the packed [3d, d] operand of the QKV GEMM
        for mod in model.modules():
            if isinstance(mod, RelativeGlobalAttention):
                for p in (mod.Wq.weight, mod.Wk.weight, mod.Wv.weight, mod.Wq.bias, mod.Wk.bias, mod.Wv.bias):
                    add(p)
        for p in model.parameters():
            add(p)
        self.params = order
        dev = order[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs the model on a CUDA device (no CPU fallback)")
        sizes = [(p.numel() + 63) // 64 * 64 for p in order]    # 256-byte aligned fp32 slots = 128-byte aligned slots of the 16-bit shadow (TMA needs 16)
        self.n = sum(sizes)
        self.flat_p = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        for p, sz in zip(order, sizes):
            view = self.flat_p[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
            off += sz
        # Kernels may write a gradient straight into its flat_g view (instead of handing a temporary
        # to autograd, which then launches one `grad += tmp` per parameter) as long as that view is
        # known to be zero: ``fresh`` holds the ids of the parameters not written since zero_grad().
        self.fresh = set()
        self.flat_lp = None           # 16-bit shadow of flat_p, (re)written by refresh_lp()
        for p in order:
            p._mt_opt = self
        self.lr, self.betas, self.eps = lr, betas, eps
        self.param_groups = [{"lr": lr, "params": order}]     # what CustomSchedule.step() touches
        self.step_count = 0
        self.pg = process_group
        self.grad_accum = grad_accum

    def zero_grad(self, set_to_none: bool = False):
        self.flat_g.zero_()
        self.fresh = {id(p) for p in self.params}

    def refresh_lp(self, act: torch.dtype) -> None:
        """One cast launch over the whole flat parameter buffer (always from the current fp32 values, so
        parameters changed behind the optimizer's back -- load_state_dict, manual edits -- are picked up)."""
        if self.flat_lp is None or self.flat_lp.dtype != act:
            self.flat_lp = torch.empty(self.n, dtype=act, device=self.flat_p.device)
        ops.cast(self.flat_p, self.flat_lp)

    def all_reduce_grads(self):
        """Data-parallel exchange: sum of the flat gradient over ranks (NCCL, one call)."""
        from .parallel import all_reduce_flat_
        return all_reduce_flat_(self.flat_g, self.pg)

    def step(self):
        world = self.all_reduce_grads()
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, None, lr, self.betas[0], self.betas[1],
                      self.eps, self.step_count, 1.0 / (world * self.grad_accum))

    def state_dict(self):
        return {"step": self.step_count, "m": self.m, "v": self.v, "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.param_groups[0]["lr"] = sd["lr"]
