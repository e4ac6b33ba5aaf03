"""The oracle (oracle/restate.py) against the fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py) -- and against the live reference when /root/reference exists."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as O
from oracle.ref_import import reference_available, reference_dir

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: z[k] for k in z.files}


def params_of(z, prefix="p:"):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in z.items() if k.startswith(prefix)}


def test_train_small_forward_loss_grads():
    z = load("train_small.npz")
    d, V, pad, layers, L, B = z["meta"].tolist()
    p = {k: v.clone().requires_grad_(True) for k, v in params_of(z).items()}
    x, y = torch.from_numpy(z["x"]), torch.from_numpy(z["y"])
    logits = O.model_forward(x, p, L, pad)
    np.testing.assert_allclose(logits.detach().numpy(), z["logits"], rtol=0, atol=2e-5)
    loss = O.smooth_ce(logits, y, 0.1, V, pad)
    assert abs(float(loss) - float(z["loss"])) < 2e-6
    loss.backward()
    for k, v in p.items():
        g = z["g:" + k]
        err = np.abs(v.grad.numpy() - g).max()
        assert err <= 1e-6 + 1e-4 * np.abs(g).max(), (k, err)
    assert float(O.categorical_accuracy(logits, y)) == pytest.approx(float(z["acc"]))
    assert (O.logits_bucket(logits).numpy() == z["bucket"]).all()


def test_train_small_eval_weights():
    z = load("train_small.npz")
    d, V, pad, layers, L, B = z["meta"].tolist()
    p = params_of(z)
    with torch.no_grad():
        _, ws = O.model_forward(torch.from_numpy(z["x"]), p, L, pad, return_weights=True)
    np.testing.assert_allclose(ws[0].numpy(), z["w0"], atol=2e-6)
    np.testing.assert_allclose(ws[1].numpy(), z["w1"], atol=2e-6)


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
@pytest.mark.parametrize("tag", ["none", "causal"])
def test_rga_module(case, tag):
    z = load("rga_small.npz")
    h, d, max_seq, L = z["meta"]["abcd".index(case)].tolist()
    p = {"rga." + k[len(case) + 3:]: torch.from_numpy(v).clone().requires_grad_(True)
         for k, v in z.items() if k.startswith(case + ":p:")}
    x = torch.from_numpy(z[case + ":x"]).clone().requires_grad_(True)
    ar = torch.arange(L)
    mask = (ar[None, :] > ar[:, None])[None, None] if tag == "causal" else None
    out, w = O.rga_forward(x, p, "rga.", h, max_seq, mask)
    np.testing.assert_allclose(out.detach().numpy(), z[f"{case}:{tag}:out"], atol=3e-5)
    if case == "a":
        np.testing.assert_allclose(w.detach().numpy(), z[f"{case}:{tag}:w"], atol=2e-6)
    wgt = torch.cos(torch.arange(out.numel(), dtype=torch.float32)).reshape(out.shape)
    (out * wgt).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), z[f"{case}:{tag}:dx"], atol=5e-4, rtol=1e-4)
    for k, v in p.items():
        g = z[f"{case}:{tag}:g:{k[4:]}"]
        assert np.abs(v.grad.numpy() - g).max() <= 1e-5 + 2e-4 * np.abs(g).max(), k
    # the closed form the kernels implement == the literal skew
    with torch.no_grad():
        q = O._split_heads(torch.nn.functional.linear(x, p["rga.Wq.weight"], p["rga.Wq.bias"]), h)
        k_ = O._split_heads(torch.nn.functional.linear(x, p["rga.Wk.weight"], p["rga.Wk.bias"]), h)
        v_ = O._split_heads(torch.nn.functional.linear(x, p["rga.Wv.weight"], p["rga.Wv.bias"]), h)
        o_cf, lse = O.rga_closed_form(q, k_, v_, p["rga.E"], max_seq, causal=(tag == "causal"))
        o_lit = torch.matmul(w, v_)
        assert (o_cf - o_lit).abs().max() < 2e-5
        s_lit = O.rga_scores(q, k_, p["rga.E"], max_seq)
        if mask is not None:
            s_lit = s_lit.masked_fill(mask, float("-inf"))
        assert (torch.logsumexp(s_lit, -1) - lse).abs().max() < 2e-5


def test_decode_small():
    z = load("decode_small.npz")
    d, V, pad, layers, max_seq, steps, thr = z["meta"].tolist()
    p = params_of(z)
    prior = torch.from_numpy(z["prior"])
    with torch.no_grad():
        ids, zs = O.generate_causal_greedy(prior, steps, p, max_seq, pad)
        lit = O.generate_literal_greedy(prior, steps, p, max_seq, thr)
    assert (ids.numpy() == z["causal_ids"]).all()
    np.testing.assert_allclose(zs.numpy(), z["causal_logits"], atol=3e-5)
    assert (lit.numpy() == z["literal_ids"]).all()


def test_decode_small_pad_tokens_in_prior():
    """Pad tokens inside the prior (and a generated pad token): the causal oracle masks them as keys exactly
    as the reference's recompute does (MT/utils.py:73)."""
    z = load("decode_small.npz")
    d, V, pad, layers, max_seq, steps, thr = z["meta"].tolist()
    with torch.no_grad():
        ids, zs = O.generate_causal_greedy(torch.from_numpy(z["prior_pad"]), steps, params_of(z), max_seq, pad)
    assert (ids.numpy() == z["causal_ids_pad"]).all()
    np.testing.assert_allclose(zs.numpy(), z["causal_logits_pad"], atol=2e-5)


def _seeded_drop_in_params(d, layers, L):
    """state_dict of the drop-in module built under torch.manual_seed(0): the reference's own seed-0 weights
    (same constructors in the same order; test_live_reference_matches_oracle re-checks the equality live)."""
    import musicgeneration_b200 as mtb
    torch.manual_seed(0)
    m = mtb.MusicTransformer(embedding_dim=d, vocab_size=390, num_layer=layers, max_seq=L, dropout=0.0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def test_config_a_and_b_summaries_pin_the_oracle_at_the_benchmarked_shapes():
    """BASELINE configs A (d256) and B (d512) at L = 2048: the oracle on the seed-0 weights reproduces the loss
    and logits the unmodified reference recorded (oracle/make_golden.py --config-a / --config-b)."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    za, zb = load("config_a_summary.npz"), load("config_b_summary.npz")
    x, y = O.synthetic_ids(2, 2048, 388)
    with torch.no_grad():
        la = O.model_forward(x, _seeded_drop_in_params(256, 6, 2048), 2048, 388)
        np.testing.assert_allclose(la[:, ::256, ::39].numpy(), za["logits_slice"], atol=3e-5)
        assert abs(float(O.smooth_ce(la, y, 0.1, 390, 388)) - float(za["loss"])) < 1e-5
        pb = _seeded_drop_in_params(512, 6, 2048)
        lb = O.model_forward(x[:1], pb, 2048, 388)
        np.testing.assert_allclose(lb[:, ::16].numpy(), zb["logits_rows"][:1], atol=5e-5)
        # first decode steps of the config-B fixture (pad token inside the second prior)
        ids, zs = O.generate_causal_greedy(torch.from_numpy(zb["prior"]), 6, pb, 2048, 388)
    assert (ids.numpy() == zb["causal_ids"][:, :ids.shape[1]]).all()
    np.testing.assert_allclose(zs.numpy(), zb["causal_logits"][:6], atol=5e-5)


def test_pe_and_schedule():
    z = load("misc.npz")
    assert (O.sinusoid_table(40, 64).astype(np.float32) == z["pe"].astype(np.float32)).all()
    np.testing.assert_allclose(O.sinusoid_table(40, 64), z["pe"], rtol=0, atol=1e-15)
    for s, r in zip(z["rate_steps"].tolist(), z["rates"].tolist()):
        assert O.noam_rate(s, 256) == pytest.approx(r, rel=1e-12)


def test_mask_semantics():
    x = torch.tensor([[3, 9, 7, 9], [1, 2, 3, 4]])
    m = O.look_ahead_mask(x, 9, 4)
    assert m.shape == (2, 1, 4, 4)
    assert m[0, 0, 2].tolist() == [False, True, False, True]
    assert m[1, 0, 0].tolist() == [False, True, True, True]


def test_sampler_semantics():
    z = torch.tensor([[0.0, 3.0, 1.0, 2.0, -1.0]])
    assert O.sample_topk_from_uniform(z, torch.tensor([0.0]), 1.0, 2).item() == 1
    assert O.sample_topk_from_uniform(z, torch.tensor([0.999]), 1.0, 2).item() == 3
    torch.manual_seed(0)
    u = torch.rand(4000)
    ids = O.sample_topk_from_uniform(z.expand(4000, -1), u, 0.7, 3)
    assert set(ids.tolist()) <= {1, 2, 3}
    pr = torch.softmax(z[0, [1, 2, 3]] / 0.7, 0)
    freq = torch.stack([(ids == i).float().mean() for i in (1, 2, 3)])
    assert (freq - pr).abs().max() < 0.03


def test_staged_reference_is_the_reference(tmp_path):
    """oracle/make_ref.py byte-compiles the reference's own files; the staged modules load WITHOUT the source
    tree (a fresh interpreter with the source candidates removed) and compute what the sources compute."""
    import subprocess
    import sys
    from oracle import make_ref
    from oracle.ref_import import staged_dir
    if reference_dir() is not None:
        assert make_ref.build(verbose=False) is not None
    if staged_dir() is None:
        pytest.skip("no staged reference (and no sources to stage it from)")
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "import oracle.ref_import as ri\n"
        "ri._CANDIDATES[:] = []\n"
        "R = ri.load_reference(); assert R.compiled\n"
        "R.config.pad_token = 70\n"
        "torch.manual_seed(11)\n"
        "m = R.network.MusicTransformer(embedding_dim=192, vocab_size=72, num_layer=2, max_seq=32, dropout=0.0)\n"
        "from oracle import restate as O\n"
        "x, y = O.synthetic_ids(2, 32, 70, seed=5)\n"
        "m.train(); ref = m(x)\n"
        "ours = O.model_forward(x, {k: v.detach() for k, v in m.state_dict().items()}, 32, 70)\n"
        "assert (ref - ours).abs().max() < 2e-5\n"
        "print('staged ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "staged ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not reference_available(), reason="neither the reference tree nor its staged modules are present")
def test_live_reference_matches_oracle():
    """Re-run the live comparison (different seed/shape from the fixture) when possible: in the build container
    against the sources, on the GPU box against the compiled modules oracle/make_ref.py staged."""
    from oracle.ref_import import load_reference
    R = load_reference()
    R.config.pad_token = 70
    torch.manual_seed(11)
    m = R.network.MusicTransformer(embedding_dim=192, vocab_size=72, num_layer=2, max_seq=32,
                                   dropout=0.0)
    x, y = O.synthetic_ids(2, 32, 70, seed=5)
    x[0, 20:] = 70
    m.train()
    ref = m(x)
    p = {k: v.detach() for k, v in m.state_dict().items()}
    ours = O.model_forward(x, p, 32, 70)
    assert (ref - ours).abs().max() < 2e-5
    l_ref = R.criterion.SmoothCrossEntropyLoss(0.1, 72, 70)(ref, y)
    assert abs(float(l_ref) - float(O.smooth_ce(ours, y, 0.1, 72, 70))) < 2e-6
    # the drop-in module built under the same seed has the reference's initial weights bit for bit
    import musicgeneration_b200 as mtb
    torch.manual_seed(11)
    mine = mtb.MusicTransformer(embedding_dim=192, vocab_size=72, num_layer=2, max_seq=32, dropout=0.0)
    sd = mine.state_dict()
    assert list(sd.keys()) == list(p.keys()) and all(torch.equal(sd[k], p[k]) for k in p)


# ---- data feed (MT/data.py) ---------------------------------------------------------------
def _replay(D):
    from oracle.make_golden import data_feed_script
    out = {}

    def record(name, *arrs):
        for i, a in enumerate(arrs):
            out[f"{name}:{i}"] = np.asarray(a)

    data_feed_script(D, record)
    return out


def test_data_oracle_matches_reference_golden(tmp_path):
    """DataOracle replays the call script the UNMODIFIED MT/data.py ran in the build container: same
    split, same windows, same failed draws, bit for bit."""
    z = load("data_feed.npz")
    names = O.write_token_corpus(str(tmp_path))
    assert sorted(os.path.relpath(n, tmp_path) for n in names) == sorted(z["files"].tolist())
    D = O.DataOracle(str(tmp_path), 30, files=[os.path.join(tmp_path, f) for f in z["files"].tolist()])
    for k in ("train", "valid", "test"):
        assert [os.path.relpath(f, tmp_path) for f in D.file_dict[k]] == z["split:" + k].tolist()
    got = _replay(D)
    for k, v in got.items():
        assert v.dtype == z[k].dtype and v.shape == z[k].shape, k
        assert (v == z[k]).all(), k


@pytest.mark.skipif(reference_dir() is None, reason="reference tree not present")
def test_data_oracle_matches_live_reference(tmp_path, monkeypatch):
    import functools
    from oracle.ref_import import load_reference
    R = load_reference()
    O.write_token_corpus(str(tmp_path), n_files=31, seed=9)
    monkeypatch.setattr(torch, "load", functools.partial(torch.load, weights_only=False))
    ref = R.data.Data(str(tmp_path), 30)
    got_ref = _replay(ref)
    got = _replay(O.DataOracle(str(tmp_path), 30, files=ref.files))
    assert got.keys() == got_ref.keys()
    for k in got:
        assert (got[k] == got_ref[k]).all(), k
