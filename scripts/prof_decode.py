"""A few KV-cached decode steps of BASELINE config B (32 sequences, context ~1000) for ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import musicgeneration_b200 as mtb

dev = torch.device("cuda:0")
mtb.config.pad_token = 388
torch.manual_seed(0)
m = mtb.MusicTransformer(embedding_dim=512, vocab_size=390, num_layer=6, max_seq=2048, dropout=0.0).to(dev)
m.set_precision(os.environ.get("PREC", "bf16"))
m.eval()
B, P, n = int(os.environ.get("PB", 32)), int(os.environ.get("PP", 1000)), int(os.environ.get("PN", 8))
prior = torch.randint(0, 388, (B, P), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(2):
    e0.record()
    out = m.generate(prior, length=n, temperature=1.0, top_k=32)
    e1.record()
    torch.cuda.synchronize()
    print("generate ms", e0.elapsed_time(e1), "per step", e0.elapsed_time(e1) / (P + n - 1))
