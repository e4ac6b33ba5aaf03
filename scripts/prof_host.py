"""Host-side (Python) cost of enqueuing one config-B train step: cProfile over a few steps."""
import cProfile
import os
import pstats
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import musicgeneration_b200 as mtb
from musicgeneration_b200.optim import FlatAdam

dev = torch.device("cuda:0")
d, V, pad, layers, L, Bg = 512, 390, 388, 6, 2048, 16
mtb.config.pad_token = pad
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


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
