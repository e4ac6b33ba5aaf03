"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/*.h declares,
and the host logic (masks, tables, packing, schedule, state_dict layout) behaves like the
reference.  No compute call is made without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import restate as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_lib():
    from musicgeneration_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "mt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in mt_b200.h but not exported"
    from musicgeneration_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert _lib.load().mt_version() >= 100


def test_ops_refuse_cpu_tensors(built_lib):
    import musicgeneration_b200 as mtb
    m = mtb.MusicTransformer(embedding_dim=64, vocab_size=20, num_layer=1, max_seq=8, dropout=0.0)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m(torch.zeros(1, 8, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mtb.SmoothCrossEntropyLoss(0.1, 20, 18)(torch.zeros(1, 8, 20), torch.zeros(1, 8, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m.generate(torch.zeros(1, 2, dtype=torch.long), 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "musicgeneration_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn
            assert "/root/reference" not in src, fn


def test_state_dict_layout_and_param_count():
    import musicgeneration_b200 as mtb
    z = np.load(os.path.join(GOLD, "train_small.npz"))
    d, V, pad, layers, L, B = z["meta"].tolist()
    m = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L)
    ref = [(k[2:], tuple(z[k].shape)) for k in z.files if k.startswith("p:")]
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == ref
    big = mtb.MusicTransformer()      # reference defaults: d256, V390, 6 layers, max_seq 2048
    assert sum(p.numel() for p in big.parameters()) == 2967174
    assert big.Decoder.enc_layers[0].rga.h == 4 and big.Decoder.enc_layers[0].rga.dh == 64
    assert big.infer is False and big.max_seq == 2048 and big.vocab_size == 390
    with pytest.raises(AttributeError):
        mtb.MusicTransformer(loader_path="x")      # same failure as MT/network.py:19-20


def test_qkv_packing_keeps_parameters_and_values():
    import musicgeneration_b200 as mtb
    rga = mtb.RelativeGlobalAttention(h=2, d=128, max_seq=16)
    before = {k: v.clone() for k, v in rga.state_dict().items()}
    w, b = rga.packed()
    assert w.shape == (384, 128) and b.shape == (384,)
    assert torch.equal(w[128:256], before["Wk.weight"]) and torch.equal(b[256:], before["Wv.bias"])
    assert rga.Wk.weight.data_ptr() == w.data_ptr() + 128 * 128 * 4
    w2, _ = rga.packed()                       # second call: already adjacent, no copy
    assert w2.data_ptr() == w.data_ptr()
    for k, v in rga.state_dict().items():
        assert torch.equal(v, before[k])
    rga.Wq.weight.data.add_(1.0)               # an optimizer-style in-place update is seen
    assert torch.equal(rga.packed()[0][:128], before["Wq.weight"] + 1.0)
    rga.load_state_dict(before)
    assert torch.equal(rga.packed()[0][:128], before["Wq.weight"])


def test_masks_tables_schedule():
    import musicgeneration_b200 as mtb
    from musicgeneration_b200 import utils
    x = torch.tensor([[3, 9, 7, 9], [1, 2, 3, 4]])
    _, _, m = utils.get_masked_with_pad_tensor(4, x, x, 9)
    assert m.causal and m.pad_keys.tolist() == [[0, 1, 0, 1], [0, 0, 0, 0]]
    assert torch.equal(utils.materialize(m, 4), O.look_ahead_mask(x, 9, 4))
    with pytest.raises(RuntimeError, match="must match the size"):
        utils.get_masked_with_pad_tensor(8, x, x, 9)
    assert torch.equal(utils.sequence_mask(torch.tensor([1, 3]), 4),
                       torch.tensor([[True, False, False, False], [True, True, True, False]]))
    # dense reference-style masks are decomposed
    dense = O.look_ahead_mask(x, 9, 4)
    mm = mtb.layers.as_mask(dense, 2, 4)
    assert mm.causal and mm.pad_keys.tolist() == [[0, 1, 0, 1], [0, 0, 0, 0]]
    with pytest.raises(NotImplementedError):
        mtb.layers.as_mask(torch.rand(2, 1, 4, 4) > 0.5, 2, 4)
    z = np.load(os.path.join(GOLD, "misc.npz"))
    assert (mtb.layers.sinusoid_table(40, 64)[0].astype(np.float32) == z["pe"].astype(np.float32)).all()
    sched = mtb.CustomSchedule(256, optimizer=None)
    for s, r in zip(z["rate_steps"].tolist(), z["rates"].tolist()):
        assert sched.rate(s) == pytest.approx(r, rel=1e-12)
    assert mtb.config.vocab_size == 309 and mtb.config.pad_token == 308 and mtb.config.threshold_len == 500


def test_custom_schedule_drives_optimizer():
    import musicgeneration_b200 as mtb
    p = torch.nn.Parameter(torch.ones(3))
    opt = torch.optim.Adam([p], lr=0, betas=(0.9, 0.98), eps=1e-9)
    sched = mtb.CustomSchedule(256, warmup_steps=10, optimizer=opt)
    p.grad = torch.ones(3)
    sched.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(O.noam_rate(1, 256, 10))
    assert float(p[0]) < 1.0
