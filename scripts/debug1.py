import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import musicgeneration_b200 as mtb
from musicgeneration_b200 import engine, ops
from oracle import restate as O
dev = torch.device("cuda:0")

def rel(a, b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-30))

def layerwise(d, V, pad, layers, L, B, prec, with_pad):
    mtb.config.pad_token = pad
    p = O.init_params(d, V, layers, L, seed=0)
    x, y = O.synthetic_ids(B, L, pad)
    if with_pad:
        x[1, L-10:] = pad
    m = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.0).to(dev)
    m.load_state_dict(p, strict=True); m.set_precision(prec); m.train()
    # oracle per-layer
    h = d // 64
    mask = O.look_ahead_mask(x, pad, L)
    emb = p["Decoder.embedding.weight"]
    xo = emb[x.long()] * math.sqrt(d) + torch.from_numpy(O.sinusoid_table(L, d)[None, :L]).float()
    enc = m.Decoder
    cfg = enc.cfg()
    Ws = [l.weights(cfg.act) for l in enc.enc_layers]
    pe = enc.pos_encoding.table(dev)
    ids32 = x.to(dev).to(torch.int32).contiguous()
    _, _, lm = mtb.utils.get_masked_with_pad_tensor(L, ids32, ids32, pad)
    xg = torch.empty((B*L, d), device=dev); xlp = torch.empty((B*L, d), dtype=cfg.act, device=dev) if cfg.act != torch.float32 else None
    ops.embed_pos_fwd(ids32, enc.embedding.weight.data, pe, xg, xlp, 0, math.sqrt(d), 0.0, 0, 0)
    print(f"[{prec} d{d} L{L} B{B} pad{with_pad}] embed rel", rel(xg.cpu().view(B, L, d), xo))
    xl = xlp if xlp is not None else xg
    for li in range(layers):
        xo, _ = O.encoder_layer_forward(xo, p, f"Decoder.enc_layers.{li}.", h, L, mask)
        xg, xl, s, _ = engine.layer_fwd(xg, xl, Ws[li], cfg, B, L, lm, 0, 1, True, False)
        # attention block pieces
        print(f"   layer {li} out rel", rel(xg.cpu().view(B, L, d), xo), " finite", bool(torch.isfinite(xg).all()))
    torch.cuda.synchronize()

layerwise(128, 96, 94, 2, 64, 3, "fp32", True)
layerwise(128, 96, 94, 2, 64, 3, "bf16", True)
layerwise(128, 96, 94, 2, 64, 3, "bf16", False)
layerwise(128, 96, 94, 2, 64, 2, "bf16", True)
layerwise(256, 390, 388, 2, 512, 4, "fp32", False)
layerwise(256, 390, 388, 2, 256, 2, "fp32", False)
layerwise(256, 390, 388, 2, 128, 2, "fp32", False)
layerwise(128, 390, 388, 2, 512, 2, "fp32", False)

# decode
z = np.load("tests/golden/decode_small.npz")
d, V, pad, layers, max_seq, steps, thr = z["meta"].tolist()
p = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p:")}
mtb.config.pad_token = pad
m = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=max_seq, dropout=0.0).to(dev)
m.load_state_dict(p, strict=True); m.eval()
prior = torch.from_numpy(z["prior"]).to(dev)
ids, sl = m.generate(prior, length=steps, greedy=True, return_logits=True)
ref = torch.from_numpy(z["causal_logits"])
for s in range(0, steps, 4):
    print("decode step", s, "logit err", float((sl[s].cpu()-ref[s]).abs().max()), ids[:, 3+s].tolist(), z["causal_ids"][:, 3+s].tolist())
