"""world_size-2 tests of the data-parallel host logic on CPU (gloo): the gradient exchange is a sum
all-reduce over the flat buffer with 1/world folded into the optimizer scale; sampling shards
sequences without communication."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from musicgeneration_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # each rank's local gradient; after the exchange every rank holds the sum
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    ws = parallel.all_reduce_flat_(g)
    assert ws == world
    expect = torch.arange(10, dtype=torch.float32) * sum(k + 1 for k in range(world))
    assert torch.equal(g, expect)
    # replicas start identical
    lin = torch.nn.Linear(3, 2)
    torch.manual_seed(100 + rank)
    with torch.no_grad():
        lin.weight.normal_()
    parallel.broadcast_params_(lin, src=0)
    gathered = [torch.zeros_like(lin.weight) for _ in range(world)]
    dist.all_gather(gathered, lin.weight.data)
    assert all(torch.equal(gathered[0], t) for t in gathered)
    # global mean over non-pad tokens of all ranks
    m = parallel.global_mean_loss(torch.tensor(10.0 * (rank + 1)), torch.tensor(5.0 + rank))
    assert float(m) == pytest.approx((10.0 + 20.0) / (5.0 + 6.0))
    # bucketed exchange: buckets started early (in backward order) and the rest at finish() give the same sums
    flat = torch.arange(12, dtype=torch.float32) * (rank + 1)
    ex = parallel.BucketedExchange(flat, [(0, 5), (5, 9), (9, 12)])
    ex.ready(2)
    ex.ready(1)
    ex.ready(1)                                   # idempotent
    assert ex.finish() == world and ex.launch_order == [2, 1, 0]
    assert torch.equal(flat, torch.arange(12, dtype=torch.float32) * sum(k + 1 for k in range(world)))
    ex.reset()
    assert ex.finish() == world and ex.launch_order == [0, 1, 2]      # nothing started early: all at finish()
    ex.reset()
    flat.copy_(torch.arange(12, dtype=torch.float32) * (rank + 1))
    ex.ready(1)
    assert list(ex.finish_each()) == [1, 0, 2] and not ex.works       # per-bucket completion, launch order
    assert torch.equal(flat, torch.arange(12, dtype=torch.float32) * sum(k + 1 for k in range(world)))
    # sampling shards: disjoint cover of the 7 sequences
    lo, hi = parallel.shard_range(7, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    assert spans == [(0, 4), (4, 7)]
    dist.destroy_process_group()
    out.put(rank)


def test_two_rank_gradient_exchange_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_shard_range_covers_everything():
    from musicgeneration_b200.parallel import shard_range
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
