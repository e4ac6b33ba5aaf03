"""In-model timing of every C-ABI call of one config-B train step with CUDA events on the launching
stream (no profiler, no cache flush: the numbers include the real cache state and launch gaps).
Aggregated per op and, for GEMMs, per (M, N, K, majors, epilogue).  Also prints the time between ops
(host launch gaps seen by the GPU)."""
import collections
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import musicgeneration_b200 as mtb
from musicgeneration_b200 import ops
from musicgeneration_b200.optim import FlatAdam

dev = torch.device("cuda:0")
d, V, pad, layers, L, Bg = 512, 390, 388, 6, 2048, int(os.environ.get("PB", 16))
mtb.config.pad_token = pad
torch.manual_seed(0)
model = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=float(os.environ.get("PDROP", 0.2)),
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


for _ in range(3):
    step()
torch.cuda.synchronize()

records = []
NAMES = ["embed_pos_fwd", "embed_pos_bwd", "add_ln_fwd", "add_ln_bwd", "gemm", "colsum", "cast", "cast2d",
         "transpose_cast", "rga_fwd", "rga_bwd", "smooth_ce_fwd", "smooth_ce_bwd", "adam_step"]
orig = {n: getattr(ops, n) for n in NAMES}


def wrap(name, fn):
    def w(*a, **k):
        key = name
        if name == "gemm":
            M, N, K = a[3], a[4], a[5]
            tA, tB = a[9], a[10]
            epi = "+".join(t for t, on in (("bias", k.get("bias") is not None), ("add", k.get("addend") is not None),
                                          ("relu", k.get("relu")), ("mask", k.get("relu_mask"))) if on)
            key = f"gemm M={M} N={N} K={K} {'T' if tA else 'N'}{'T' if tB else 'N'} out={str(a[2].dtype)[6:]} {epi}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        records.append((key, e0, e1))
        return r
    return w


for n in NAMES:
    setattr(ops, n, wrap(n, orig[n]))
STEPS = 5
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(STEPS):
    step()
s1.record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
busy = 0.0
for key, e0, e1 in records:
    t = e0.elapsed_time(e1) * 1e3
    agg[key][0] += 1
    agg[key][1] += t
    busy += t
total = s0.elapsed_time(s1) * 1e3
print(f"step {total / STEPS:.0f} us; inside timed ops {busy / STEPS:.0f} us; outside (torch kernels, gaps) {(total - busy) / STEPS:.0f} us")
for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / STEPS:9.1f} us/step  n/step={n / STEPS:5.1f}  avg {t / n:8.1f} us  {key}")
