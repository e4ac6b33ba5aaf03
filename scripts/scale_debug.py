"""Where does the N>1 train step go?  Run under torchrun (or alone).  Per rank: device time per step,
host enqueue time per step (how long Python needs to issue one step without waiting for the GPU),
with and without the gradient all-reduce, and the all-reduce alone."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import musicgeneration_b200 as mtb
from musicgeneration_b200.optim import FlatAdam

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
d, V, pad, layers, L, Bg = 512, 390, 388, 6, 2048, 16
mtb.config.pad_token = pad
torch.manual_seed(0)
model = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.2,
                             precision="bf16").to(dev)
model.train()
crit = mtb.SmoothCrossEntropyLoss(0.1, V, pad)
opt = FlatAdam(model, lr=0.0, betas=(0.9, 0.98), eps=1e-9)
sched = mtb.CustomSchedule(d, optimizer=opt)
x = torch.randint(0, pad, (Bg, L), dtype=torch.int32, device=dev)
y = torch.randint(0, pad, (Bg, L), dtype=torch.int32, device=dev)


def step():
    opt.zero_grad()
    loss = crit(model(x), y)
    loss.backward()
    sched.step()


def measure(tag, K=10):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"[rank {rank}] {tag}: device {e0.elapsed_time(e1) / K:.2f} ms/step, host enqueue {t_host / K * 1e3:.2f} ms/step",
          flush=True)


print(f"[rank {rank}] cpu_count {os.cpu_count()} affinity {len(os.sched_getaffinity(0))} torch threads {torch.get_num_threads()} "
      f"dev {torch.cuda.current_device()} {torch.cuda.get_device_name()}", flush=True)
measure("with all-reduce" if world > 1 else "single")
if world > 1:
    real = opt.all_reduce_grads
    opt.all_reduce_grads = lambda: world
    measure("all-reduce skipped")
    opt.all_reduce_grads = real
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(opt.flat_g)
    e1.record()
    torch.cuda.synchronize()
    print(f"[rank {rank}] all-reduce of {opt.flat_g.numel() * 4 / 1e6:.1f} MB alone: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    dist.destroy_process_group()
