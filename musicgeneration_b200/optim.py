"""Optimizer side of the train step ("next" row 1 of SURVEY 8f): Adam with the reference's
hyper-parameters (MT/train.py:143: betas (0.9, 0.98), eps 1e-9, lr driven by the Noam
schedule of MT/criterion.py:70-96) as ONE fused kernel over a flat fp32 parameter buffer, plus
the data-parallel gradient exchange: one NCCL all-reduce over the flat gradient buffer.

``FlatAdam`` re-points every ``param.data`` / ``param.grad`` at views of two flat buffers (the
state_dict keys and shapes are untouched; Wq/Wk/Wv stay adjacent so the fused QKV GEMM needs no
repacking)."""
from __future__ import annotations

import os

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

        # keep Wq/Wk/Wv weights (and biases) adjacent: the packed [3d, d] operand of the QKV GEMM.  A triple
        # shares ONE slot (padding only after it), so adjacency holds for every d, not only d % 64 == 0
        glue = set()            # ids of parameters that must directly follow their predecessor (no padding between)
        # ... emitted at the position of the triple's first member in model.parameters() order, so every encoder
        # layer's parameters stay one contiguous range of the flat buffers (= one gradient-exchange bucket)
        trip_of = {}
        for mod in model.modules():
            if isinstance(mod, RelativeGlobalAttention):
                for trip in ((mod.Wq.weight, mod.Wk.weight, mod.Wv.weight), (mod.Wq.bias, mod.Wk.bias, mod.Wv.bias)):
                    if all(p.requires_grad for p in trip):
                        for p in trip:
                            trip_of[id(p)] = trip
        for p in model.parameters():
            trip = trip_of.get(id(p))
            if trip is not None and all(id(q) not in seen for q in trip):
                for i, q in enumerate(trip):
                    add(q)
                    if i:
                        glue.add(id(q))
            else:
                add(p)
        self.params = order
        self._model_order = [p for p in model.parameters() if p.requires_grad]
        dev = order[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs the model on a CUDA device (no CPU fallback)")
        # 256-byte aligned fp32 slots = 128-byte aligned slots of the 16-bit shadow (TMA needs 16)
        sizes = []
        for i, p in enumerate(order):
            last_of_group = i + 1 == len(order) or id(order[i + 1]) not in glue
            sizes.append((p.numel() + 63) // 64 * 64 if last_of_group else p.numel())
        # a glued group starts 64-aligned because every group END is padded; inside, offsets follow numel
        self.n = sum(sizes)
        self.flat_p = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        self._offsets = []
        for p, sz in zip(order, sizes):
            view = self.flat_p[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
            self._offsets.append(off)
            off += sz
        # Kernels may write a gradient straight into its flat_g view (instead of handing a temporary
        # to autograd, which then launches one `grad += tmp` per parameter) as long as that view is
        # known to be zero: ``fresh`` holds the ids of the parameters not written since zero_grad().
        self.fresh = set()
        self.flat_lp = None           # 16-bit shadow of flat_p, (re)written by refresh_lp()
        # ---- gradient-exchange buckets: the backward finishes the vocabulary projection first, then
        # the encoder layers last to first, then the embedding; a bucket = one layer's contiguous range (the
        # embedding and the vocabulary projection are buckets of their own).  Models without
        # ``enc_layers.<i>.`` parameter names get a single bucket.
        import re
        names = {id(p): n for n, p in model.named_parameters()}
        self._names = names
        key_of = []
        for p in order:
            mname = re.search(r"enc_layers\.(\d+)\.", names.get(id(p), ""))
            key_of.append(int(mname.group(1)) if mname else None)
        first_layer = next((k for k in key_of if k is not None), None)
        cur, keys = (first_layer if first_layer is not None else 0), []
        # MT_DDP_SPLIT_EMB=1: the parameters in front of the first layer (the embedding, whose gradient is the LAST
        # thing the backward produces) get a bucket of their own, so that layer 0's bucket -- 16 x larger -- goes on
        # the wire one kernel earlier
        if os.environ.get("MT_DDP_SPLIT_EMB", "1") != "0" and first_layer is not None:
            cur = -1
        seen_layer = False
        for k in key_of:
            if k is not None:
                cur, seen_layer = k, True
            elif seen_layer:
                cur = 1 << 30                 # parameters after the last layer (the vocabulary projection)
            keys.append(cur)
        self._bucket_of, bounds = {}, {}
        for p, off, sz, k in zip(order, self._offsets, sizes, keys):
            lo, hi = bounds.get(k, (off, off))
            bounds[k] = (min(lo, off), max(hi, off + sz))
        self._bucket_keys = sorted(bounds)
        for p, k in zip(order, keys):
            self._bucket_of[id(p)] = self._bucket_keys.index(k)
        from .parallel import BucketedExchange
        self.exchange = BucketedExchange(self.flat_g, [bounds[k] for k in self._bucket_keys], process_group)
        self._bucket_size = [sum(1 for p in order if self._bucket_of[id(p)] == b) for b in range(len(self._bucket_keys))]
        self._bucket_ready = [set() for _ in self._bucket_keys]
        self.ready_order = []         # buckets in the order the backward completed them (since the last zero_grad)
        self._hook_hits = {}
        # MT_DDP_MODE: "overlap" (default: buckets go on the wire during the backward), "single" (everything at
        # step(), the round-1 behaviour), "none" (no exchange at all -- timing experiments only)
        self._ddp_mode = os.environ.get("MT_DDP_MODE", "overlap")
        self.sync_grads = True        # False on the non-final micro-batches of an accumulation window (DDP no_sync)
        for p in order:
            p._mt_opt = self
            # a gradient written by autograd's own accumulation (a second loss, a standalone layer backward) makes
            # the view non-zero: kernels must then accumulate through autograd again instead of assigning
            p.register_post_accumulate_grad_hook(lambda q, _s=self: _s._autograd_wrote(q))
        self.lr, self.betas, self.eps = lr, betas, eps
        self.param_groups = [{"lr": lr, "params": order}]     # what CustomSchedule.step() touches
        self.step_count = 0
        self.pg = process_group
        self.grad_accum = grad_accum

    def _autograd_wrote(self, p):
        """Post-accumulate hook.  (torch runs it once per backward for every parameter of the graph, also when the
        backward function returned None because a kernel wrote the gradient in place.)  From here on the view is
        not known to be zero, so later backwards of the window accumulate through autograd; a SECOND pass over a
        parameter whose bucket is already on the wire would add to a buffer being reduced."""
        self.fresh.discard(id(p))
        n = self._hook_hits.get(id(p), 0) + 1
        self._hook_hits[id(p)] = n
        if n > 1 and self.exchange.started[self._bucket_of[id(p)]]:
            raise RuntimeError(f"FlatAdam: a gradient ({self._names.get(id(p), '?')}) was accumulated after its bucket's all-reduce had started "
                               "(second backward without zero_grad(), with the overlapped exchange on); set "
                               "opt.sync_grads = False for all but the last backward of the window")

    def grads_ready(self, params) -> None:
        """The kernels that write the FINAL gradients of ``params`` (directly into the flat buffer) have been
        enqueued: once that holds for every parameter of a bucket, its all-reduce starts (overlapping the rest
        of the backward).  Called by the backward functions of layers.py; a no-op on one rank."""
        if not self.sync_grads or self.exchange is None or self._ddp_mode != "overlap":
            return
        for p in params:
            b = self._bucket_of.get(id(p))
            if b is None:
                continue
            self._bucket_ready[b].add(id(p))
            if len(self._bucket_ready[b]) == self._bucket_size[b] and b not in self.ready_order:
                self.ready_order.append(b)
                self.exchange.ready(b)

    def _rebind(self):
        """p.grad must alias flat_g and p.data flat_p (step() and the all-reduce read the flat buffers).
        ``model.zero_grad()`` defaults to set_to_none=True in current torch, after which autograd allocates fresh
        .grad tensors elsewhere: fold such a gradient into its view and point .grad back at it.  A parameter whose
        DATA no longer lives in flat_p (re-assigned behind the optimizer's back) would silently stop training."""
        gbase, pbase = self.flat_g.data_ptr(), self.flat_p.data_ptr()
        for p, off in zip(self.params, self._offsets):
            if p.data.data_ptr() != pbase + 4 * off:
                raise RuntimeError("FlatAdam: a parameter's data no longer aliases the flat parameter buffer "
                                   "(param.data was re-assigned after the optimizer was built)")
            g = p.grad
            if g is not None and g.data_ptr() == gbase + 4 * off:
                continue
            view = self.flat_g[off:off + p.numel()].view(p.shape)
            if g is not None:
                view.add_(g.to(view.dtype))
                self.fresh.discard(id(p))
            p.grad = view

    def zero_grad(self, set_to_none: bool = False):
        """Always keeps the gradients as (zeroed) views of the flat buffer; ``set_to_none`` is accepted for
        torch.optim compatibility and ignored."""
        self.flat_g.zero_()
        for p, off in zip(self.params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
        self.fresh = {id(p) for p in self.params}
        self.exchange.reset()
        self._bucket_ready = [set() for _ in self._bucket_keys]
        self.ready_order = []
        self._hook_hits = {}

    def refresh_lp(self, act: torch.dtype) -> None:
        """One cast launch over the whole flat parameter buffer (always from the current fp32 values, so
        parameters changed behind the optimizer's back -- load_state_dict, manual edits -- are picked up)."""
        if self.flat_lp is None or self.flat_lp.dtype != act:
            self.flat_lp = torch.empty(self.n, dtype=act, device=self.flat_p.device)
        ops.cast(self.flat_p, self.flat_lp)

    def all_reduce_grads(self):
        """Data-parallel exchange: sum of the flat gradient over ranks.  Buckets whose all-reduce was started
        during the backward (grads_ready) are only waited for; the rest is exchanged here."""
        if self._ddp_mode == "none":
            from .parallel import world_size
            return world_size(self.pg)
        return self.exchange.finish()

    def step(self):
        self._rebind()
        from .parallel import world_size
        world = world_size(self.pg)
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])
        scale = 1.0 / (world * self.grad_accum)
        if world > 1 and self._ddp_mode == "overlap":
            # bucket by bucket: the update of a bucket runs as soon as ITS all-reduce has landed, under the
            # transfers of the buckets behind it (the last one -- layer 0 + embedding -- is the only exposed one)
            for b in self.exchange.finish_each():
                lo, hi = self.exchange.buckets[b]
                if hi > lo:
                    ops.adam_step(self.flat_p[lo:hi], self.flat_g[lo:hi], self.m[lo:hi], self.v[lo:hi], None, lr,
                                  self.betas[0], self.betas[1], self.eps, self.step_count, scale)
            return
        self.all_reduce_grads()
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, None, lr, self.betas[0], self.betas[1],
                      self.eps, self.step_count, scale)

    # ---- checkpoint compatibility (MT/train.py:143,151,203: torch.optim.Adam state_dict under 'optimizer') ----
    def _slots(self):
        """(flat offset, numel, shape) of every trainable parameter in ``model.parameters()`` order -- the
        order torch.optim.Adam(model.parameters()) numbers its state entries in."""
        base = self.flat_p.data_ptr()
        return [((p.data_ptr() - base) // 4, p.numel(), tuple(p.shape)) for p in self._model_order]

    def state_dict(self):
        """The dict ``torch.optim.Adam(model.parameters(), ...).state_dict()`` would hold at this point:
        ``state[i] = {step, exp_avg, exp_avg_sq}`` per parameter index (empty before the first step) and one
        param group; a reference checkpoint written from it loads into the reference's optimizer and back."""
        group = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=float(self.param_groups[0]["lr"]),
                                 betas=tuple(self.betas), eps=self.eps).state_dict()["param_groups"][0]
        slots = self._slots()
        group["params"] = list(range(len(slots)))
        state = {}
        if self.step_count > 0:
            for i, (off, n, shape) in enumerate(slots):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.m[off:off + n].view(shape).clone(),
                            "exp_avg_sq": self.v[off:off + n].view(shape).clone()}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts a torch.optim.Adam state_dict (a reference checkpoint's 'optimizer' entry, or our own
        ``state_dict()``) and the flat {'step','m','v','lr'} form earlier versions of this class wrote."""
        if "param_groups" not in sd:
            self.step_count = int(sd["step"])
            self.m.copy_(sd["m"])
            self.v.copy_(sd["v"])
            self.param_groups[0]["lr"] = sd["lr"]
            return
        groups = sd["param_groups"]
        slots = self._slots()
        index = [i for g in groups for i in g["params"]]
        if len(index) != len(slots):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of "
                             "optimizer's group")
        g0 = groups[0]
        self.param_groups[0]["lr"] = g0["lr"]
        self.betas = tuple(g0.get("betas", self.betas))
        self.eps = g0.get("eps", self.eps)
        if g0.get("weight_decay", 0) or g0.get("amsgrad", False):
            raise ValueError("FlatAdam implements plain Adam (weight_decay=0, amsgrad=False), as MT/train.py:143 uses it")
        self.m.zero_()
        self.v.zero_()
        steps = set()
        for pos, key in enumerate(index):
            st = sd["state"].get(key)
            if st is None:
                continue
            off, n, shape = slots[pos]
            if tuple(st["exp_avg"].shape) != shape:
                raise ValueError(f"optimizer state {key}: shape {tuple(st['exp_avg'].shape)} != parameter shape {shape}")
            self.m[off:off + n].view(shape).copy_(st["exp_avg"])
            self.v[off:off + n].view(shape).copy_(st["exp_avg_sq"])
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError("FlatAdam keeps one step counter; the loaded state has per-parameter steps " + str(sorted(steps)))
        self.step_count = steps.pop() if steps else 0
