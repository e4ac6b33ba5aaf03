"""Data-parallel plumbing: one process per GPU with torch.distributed (NCCL over NVLink on the B200
box, gloo in the CPU tests).  This takes the ROLE of the reference's vendored single-process
DataParallelModel / DataParallelCriterion (MT/parallel.py:69-129 -- dead code upstream, the calls at
MT/train.py:232-235 are commented out): identical replicas, each rank draws its own batch, ONE
exchange step per optimizer step -- a sum all-reduce over the flat fp32 gradient buffer -- and an
identical Adam update everywhere.  The exchange is issued in per-layer BUCKETS while the backward is
still running (``BucketedExchange``): a bucket's all-reduce starts on the communicator's own stream as soon
as the kernels that write its gradients have been enqueued, so only the last bucket's transfer is exposed.
Batched sampling shards sequences across ranks with no communication."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (the
    torchrun contract).  Returns (rank, world, local_rank); a no-op for world == 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # the per-layer gradient buckets (a few MB each) run UNDER the backward, whose kernels fill every SM: a
            # communicator that takes fewer SMs hides better than one that moves the bytes a little faster (8 GPUs,
            # config B: 9.82 -> 9.77 ms per step); the user's own setting wins
            os.environ.setdefault("NCCL_MAX_CTAS", "8")
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n items (sequences to sample, batch rows): [lo, hi) of rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_flat_(flat: torch.Tensor, group=None) -> int:
    """In-place SUM all-reduce of a flat gradient buffer; returns the world size (the caller folds
    1/world into the optimizer's gradient scale, so no extra pass over the buffer is made)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return w


class BucketedExchange:
    """Sum all-reduce of a flat gradient buffer in contiguous buckets, each started as soon as it is final.

    ``buckets``: [(lo, hi)] element ranges of ``flat`` (disjoint, covering what must be exchanged).
    ``ready(i)`` enqueues bucket i's all-reduce asynchronously (torch.distributed orders it after the work
    already enqueued on the current stream and runs it on the communicator's stream -- NCCL over NVLink on
    the GPU box -- so it overlaps the rest of the backward); ``finish()`` exchanges every bucket that was not
    started, makes the current stream wait for all of them and returns the world size.  With one rank both
    are no-ops.  A bucket must not be written again between ``ready`` and ``finish``."""

    def __init__(self, flat: torch.Tensor, buckets, group=None):
        self.flat, self.group = flat, group
        self.buckets = [(int(lo), int(hi)) for lo, hi in buckets]
        self.started = [False] * len(self.buckets)
        self.works = []
        self.launch_order = []          # bucket indices in the order their exchange was started (tests / traces)

    def reset(self) -> None:
        if self.works:
            raise RuntimeError("BucketedExchange.reset() with exchanges in flight: call finish() first")
        self.started = [False] * len(self.buckets)
        self.launch_order = []

    def _start(self, i: int) -> None:
        lo, hi = self.buckets[i]
        self.started[i] = True
        self.launch_order.append(i)
        if hi > lo:
            self.works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                              async_op=True))

    def ready(self, i: int) -> None:
        if world_size(self.group) > 1 and not self.started[i]:
            self._start(i)

    def finish_each(self):
        """Starts what was not started and yields every bucket index once ITS exchange has been waited for (in
        launch order): the caller can consume a bucket -- the optimizer update of its range -- while later
        buckets are still on the wire."""
        w = world_size(self.group)
        if w > 1:
            for i in range(len(self.buckets)):
                if not self.started[i]:
                    self._start(i)
            order = [i for i in self.launch_order if self.buckets[i][1] > self.buckets[i][0]]
            for i, wk in zip(order, self.works):
                wk.wait()
                yield i
            for i in self.launch_order:
                if self.buckets[i][1] <= self.buckets[i][0]:
                    yield i
        else:
            for i in range(len(self.buckets)):
                yield i
        self.works = []

    def finish(self) -> int:
        w = world_size(self.group)
        if w > 1:
            for i in range(len(self.buckets)):
                if not self.started[i]:
                    self._start(i)
            for wk in self.works:
                wk.wait()
        self.works = []
        return w


def broadcast_params_(model: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every replica start from rank ``src``'s weights."""
    if world_size(group) > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)


def global_mean_loss(loss_sum: torch.Tensor, n_valid: torch.Tensor, group=None) -> torch.Tensor:
    """Mean loss over ALL ranks' non-pad tokens (the reference divides by the global non-pad count,
    MT/criterion.py:57-59): all-reduce (sum, count) and divide."""
    t = torch.stack([loss_sum.reshape(()).float(), n_valid.reshape(()).float()])
    if world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0] / t[1]
