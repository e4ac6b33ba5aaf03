"""CPU simulation of the operand-rounding scheme of the bf16 mode (extended to the
benchmarked shapes): fp32 accumulation / LayerNorm / softmax / residual stream, 16-bit rounding of every
MMA operand and of every tensor the kernels store in 16 bits.  Used to choose which operands of layer 0
must carry an 11-bit mantissa (fp16) before the kernels were changed; prints logits / loss error against
the fp32 oracle.

    python scripts/sim_precision.py [d] [L] [B] [layers]
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restate as O  # noqa: E402


def rnd(t, dt):
    return t if dt is None else t.to(dt).to(torch.float32)


def attn(q, k, v, E, max_seq, dt_qk, dt_pv, park_f16=True):
    """closed form with operand rounding; q,k,v [B,h,L,dh] fp32 already rounded by the caller"""
    B, h, L, dh = q.shape
    i = torch.arange(L)[:, None]
    j = torch.arange(L)[None, :]
    idx = (max_seq - 1 - (i - j)).clamp(0, max_seq - 1)
    qe = torch.einsum("bhld,md->bhlm", q, rnd(E, dt_qk))
    if park_f16:
        qe = qe.to(torch.float16).to(torch.float32)          # the skew scratch holds f16 pairs
    srel = torch.gather(qe, 3, idx.expand(B, h, L, L)) * (j <= i).to(q.dtype)
    s = (torch.matmul(q, k.transpose(-1, -2)) + srel) / math.sqrt(dh)
    s = s.masked_fill((j > i)[None, None], float("-inf"))
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    o = torch.matmul(rnd(p, dt_pv), v) / l                   # unnormalised P is what the MMA sees
    return o


def forward(ids, p, max_seq, scheme):
    """scheme: dict(l0=dtype for layer-0 attention operands, rest=dtype elsewhere)"""
    emb = p["Decoder.embedding.weight"]
    d = emb.size(1)
    h = d // 64
    x = emb[ids.long()] * math.sqrt(d)
    pe = torch.from_numpy(O.sinusoid_table(max_seq, d)[None, :ids.size(1), :]).to(torch.float32)
    x = x + pe
    n = O.num_layers_of(p)
    for l in range(n):
        pre = f"Decoder.enc_layers.{l}."
        dt = scheme["l0"] if l == 0 else scheme["rest"]
        dt_fc = scheme.get("l0_fc", dt) if l == 0 else dt
        rest = scheme["rest"]
        xl = rnd(x, dt)
        def lin(name, inp, dti, dto):
            return rnd(F.linear(inp, rnd(p[pre + name + ".weight"], dti), p[pre + name + ".bias"]), dto)
        sp = lambda t: t.reshape(t.shape[0], t.shape[1], h, d // h).permute(0, 2, 1, 3)
        q = sp(lin("rga.Wq", xl, dt, dt))
        k = sp(lin("rga.Wk", xl, dt, dt))
        v = sp(lin("rga.Wv", xl, dt, dt))
        o = attn(q, k, v, p[pre + "rga.E"], max_seq, dt, dt)
        o = rnd(o.permute(0, 2, 1, 3).reshape(x.shape), dt)
        a = lin("rga.fc", rnd(o, dt_fc), dt_fc, rest)
        o1 = F.layer_norm(a + x, (d,), p[pre + "layernorm1.weight"], p[pre + "layernorm1.bias"], 1e-6)
        hm = F.relu(lin("FFN_pre", rnd(o1, rest), rest, rest))
        f = lin("FFN_suf", hm, rest, rest)
        x = F.layer_norm(o1 + f, (d,), p[pre + "layernorm2.weight"], p[pre + "layernorm2.bias"], 1e-6)
    rest = scheme["rest"]
    return F.linear(rnd(x, rest), rnd(p["fc.weight"], rest), p["fc.bias"])


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    layers = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    V, pad = 390, 388
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.init_params(d, V, layers, L, seed=0)
    x, y = O.synthetic_ids(B, L, pad)
    with torch.no_grad():
        ref = O.model_forward(x, p, L, pad)
        ref_loss = float(O.smooth_ce(ref, y, 0.1, V, pad))
        bf, hf = torch.bfloat16, torch.float16
        for name, sch in (("all bf16", dict(l0=bf, rest=bf)),
                          ("layer-0 attention block fp16 (x, Wqkv, q/k/v, E, P, O, Wfc), rest bf16", dict(l0=hf, rest=bf)),
                          ("same, but Wfc / O operand of layer 0 in bf16", dict(l0=hf, l0_fc=bf, rest=bf)),
                          ("all fp16", dict(l0=hf, rest=hf))):
            out = forward(x, p, L, sch)
            err = float((out - ref).norm() / ref.norm())
            loss = float(O.smooth_ce(out, y, 0.1, V, pad))
            agree = float((out.argmax(-1) == ref.argmax(-1)).float().mean())
            print(f"d{d} L{L} B{B} {layers}L | {name}: logits rel {err:.3e}  loss rel {abs(loss - ref_loss) / ref_loss:.2e}  argmax agree {agree:.4f}",
                  flush=True)


if __name__ == "__main__":
    main()
