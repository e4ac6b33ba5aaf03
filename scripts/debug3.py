import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import musicgeneration_b200 as mtb
from oracle import restate as O
dev = torch.device("cuda:0")
def rel(a, b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-30))
z = np.load("tests/golden/train_small.npz")
d, V, pad, layers, L, B = z["meta"].tolist()
p = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p:")}
mtb.config.pad_token = pad
x = torch.from_numpy(z["x"])
mask = O.look_ahead_mask(x, pad, L)
hid_ref, _ = O.encoder_forward(x, p, L, mask)
for how in ("ctor", "after"):
    if how == "ctor":
        m = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.0, precision="bf16").to(dev)
    else:
        m = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.0).to(dev)
        m.set_precision("bf16")
    m.load_state_dict(p, strict=True); m.train()
    print(how, "precisions", m.precision, m.Decoder.precision, m.Decoder.enc_layers[0].precision, m.Decoder.enc_layers[0].rga.precision, m.Decoder.cfg().act)
    _, _, lm = mtb.utils.get_masked_with_pad_tensor(L, x.to(dev), x.to(dev), pad)
    hid, _ = m.Decoder(x.to(dev), mask=lm)
    print(how, "hidden rel", rel(hid.detach().cpu(), hid_ref), "has _mt_lp", hasattr(hid, "_mt_lp"))
    if hasattr(hid, "_mt_lp"):
        print("   lp vs hid", rel(hid._mt_lp.float().cpu().view(B, L, d), hid.detach().cpu()))
    logits = m(x.to(dev))
    print(how, "logits rel", rel(logits.detach().cpu(), torch.from_numpy(z["logits"])))
    from musicgeneration_b200.layers import _LinearFunction
    l2 = _LinearFunction.apply(m.Decoder.cfg(), hid.detach().clone(), m.fc.weight, m.fc.bias)
    print(how, "logits via clone rel", rel(l2.detach().cpu(), torch.from_numpy(z["logits"])))
