"""Persistent decode kernel at the bench shape (config B, 32 sequences x 2047 events); MT_DECODE_PROF=1 prints CTA 0's
per-phase times."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import musicgeneration_b200 as mtb

dev = torch.device("cuda:0")
seqs, events = int(os.environ.get("SEQS", 32)), int(os.environ.get("EVENTS", 2047))
mtb.config.pad_token = 388
m = mtb.MusicTransformer(embedding_dim=512, vocab_size=390, num_layer=6, max_seq=2048, dropout=0.0).to(dev)
m.set_precision("bf16")
m.eval()
prior = torch.randint(0, 388, (seqs, 1), dtype=torch.int64).to(dev)
with torch.no_grad():
    m.generate(prior, length=8, temperature=1.0, top_k=32)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.generate(prior, length=events, temperature=1.0, top_k=32)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"{seqs} x {events} events: {dt * 1e3:.1f} ms, {seqs * events / dt:.0f} events/s, {dt / events * 1e6:.1f} us/step")
