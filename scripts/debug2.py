import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import musicgeneration_b200 as mtb
from musicgeneration_b200 import engine, ops
from oracle import restate as O
dev = torch.device("cuda:0")
def rel(a, b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-30))

d, V, pad, layers, L, B = 256, 390, 388, 2, 128, 2
mtb.config.pad_token = pad
p = O.init_params(d, V, layers, L, seed=0)
x, y = O.synthetic_ids(B, L, pad)
m = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.0).to(dev)
m.load_state_dict(p, strict=True); m.train()
h = d // 64
mask = O.look_ahead_mask(x, pad, L)
xo = p["Decoder.embedding.weight"][x.long()] * math.sqrt(d) + torch.from_numpy(O.sinusoid_table(L, d)[None, :L]).float()
enc = m.Decoder; cfg = enc.cfg()
Ws = [l.weights(cfg.act) for l in enc.enc_layers]
for li in range(layers):
    pre = f"Decoder.enc_layers.{li}."
    W = Ws[li]
    print("layer", li, "Wqkv vs params", rel(W.Wqkv.cpu(), torch.cat([p[pre+"rga.Wq.weight"], p[pre+"rga.Wk.weight"], p[pre+"rga.Wv.weight"]])),
          "bqkv", rel(W.bqkv.cpu(), torch.cat([p[pre+"rga.Wq.bias"], p[pre+"rga.Wk.bias"], p[pre+"rga.Wv.bias"]])),
          "Wfc", rel(W.Wfc.cpu(), p[pre+"rga.fc.weight"]), "E", rel(W.E.cpu(), p[pre+"rga.E"]),
          "Wpre", rel(W.Wpre.cpu(), p[pre+"FFN_pre.weight"]), "Wsuf", rel(W.Wsuf.cpu(), p[pre+"FFN_suf.weight"]),
          "g1", rel(W.g1.cpu(), p[pre+"layernorm1.weight"]), "b2", float((W.b2.cpu()-p[pre+"layernorm2.bias"]).abs().max()))
    xg = xo.reshape(B*L, d).to(dev).contiguous()
    a, s, _ = engine.rga_block_fwd(xg, xg, xg, W, cfg, B, L, engine.Mask(True, None), False)
    q = O._split_heads(F.linear(xo, p[pre+"rga.Wq.weight"], p[pre+"rga.Wq.bias"]), h)
    k = O._split_heads(F.linear(xo, p[pre+"rga.Wk.weight"], p[pre+"rga.Wk.bias"]), h)
    v = O._split_heads(F.linear(xo, p[pre+"rga.Wv.weight"], p[pre+"rga.Wv.bias"]), h)
    qkv = s["qkv"].cpu().view(B, L, 3, h, 64)
    print("   q", rel(qkv[:, :, 0].permute(0, 2, 1, 3), q), "k", rel(qkv[:, :, 1].permute(0, 2, 1, 3), k), "v", rel(qkv[:, :, 2].permute(0, 2, 1, 3), v))
    o_ref, lse_ref = O.rga_closed_form(q, k, v, p[pre+"rga.E"], L, True)
    Og = s["O"].cpu().view(B, L, h, 64).permute(0, 2, 1, 3)
    print("   O", rel(Og, o_ref), "per head", [rel(Og[:, i], o_ref[:, i]) for i in range(h)], "lse", float((s["lse"].cpu()-lse_ref).abs().max()))
    a_ref, _ = O.rga_forward(xo, p, pre+"rga.", h, L, mask)
    print("   a", rel(a.cpu().view(B, L, d), a_ref))
    xo, _ = O.encoder_layer_forward(xo, p, pre, h, L, mask)
