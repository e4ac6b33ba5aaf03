"""N-rank check of the overlapped gradient exchange on real GPUs (NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py

Every rank runs the config-B model (dropout 0) on its own batch.  (1) local gradients with the exchange off
(sync_grads = False), all-gathered -> the expected sum; (2) the same backward with the bucketed exchange on:
the flat gradient buffer after finish() must equal that sum on every rank (1e-5 of the largest gradient: two
backward passes differ by the order of the fp32 dE reductions), every bucket must have been started DURING the backward in the order
vocabulary, layer n-1 .. 0, embedding, and the Adam step must leave all replicas identical."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import musicgeneration_b200 as mtb  # noqa: E402
from musicgeneration_b200.optim import FlatAdam  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d, V, pad, layers, L, B = 512, 390, 388, 6, 2048, 4
    mtb.config.pad_token = pad
    torch.manual_seed(0)
    m = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.0,
                             precision="bf16").to(dev)
    m.train()
    opt = FlatAdam(m, lr=1e-3)
    crit = mtb.SmoothCrossEntropyLoss(0.1, V, pad)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randint(0, pad, (B, L), generator=g, dtype=torch.int32).to(dev)
    y = torch.randint(0, pad, (B, L), generator=g, dtype=torch.int32).to(dev)
    opt.zero_grad()
    opt.sync_grads = False
    crit(m(x), y).backward()
    local_g = opt.flat_g.clone()
    parts = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(parts, local_g)
    expect = parts[0].clone()
    for p in parts[1:]:
        expect += p
    opt.zero_grad()
    opt.sync_grads = True
    crit(m(x), y).backward()
    started_in_backward = list(opt.exchange.launch_order)
    w = opt.all_reduce_grads()
    torch.cuda.synchronize()
    err = float((opt.flat_g - expect).abs().max())
    scale = float(expect.abs().max())
    ok = (w == world and started_in_backward == [layers + 1] + list(range(layers, -1, -1))        # vocabulary, layers n-1 .. 0, embedding
          and err <= 1e-5 * scale)      # (two backward passes differ by the order of the dE reductions)
    opt.exchange.works = []
    opt.step_count += 1
    from musicgeneration_b200 import ops
    ops.adam_step(opt.flat_p, opt.flat_g, opt.m, opt.v, None, 1e-3, 0.9, 0.98, 1e-9, opt.step_count, 1.0 / world)
    ps = [torch.empty_like(opt.flat_p) for _ in range(world)]
    dist.all_gather(ps, opt.flat_p)
    same = all(torch.equal(ps[0], q) for q in ps)
    flag = torch.tensor([1.0 if (ok and same) else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"ddp_check world={world}: buckets started in backward {started_in_backward}, max |flat_g - sum| = {err:.3e} "
              f"(scale {scale:.3e}), replicas identical after Adam: {same} -> {'OK' if float(flag) == 1.0 else 'FAILED'}")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
