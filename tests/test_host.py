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


def test_size_queries_of_the_training_pair_and_the_persistent_decode(built_lib):
    """The size queries are plain arithmetic (no device): the P stash is B h nT (nT + 1) / 2 tiles of 32 KB plus 2 x 128 fp32
    row references per tile; shapes the tcgen05 training pair does not take report 0 (the engine then keeps no stash)."""
    from musicgeneration_b200 import _lib
    lib = _lib.load()
    for B, h, L in ((16, 8, 2048), (4, 12, 4096), (2, 2, 64), (1, 3, 200)):
        nT = (L + 127) // 128
        tiles = B * h * nT * (nT + 1) // 2
        assert lib.mt_rga_stash_bytes(B, h, L, 64, _lib.MT_BF16) == tiles * (32768 + 1024)
        assert lib.mt_rga_stash_bytes(B, h, L, 64, _lib.MT_F16_BF16) == tiles * (32768 + 1024)
        assert lib.mt_rga_bwd_workspace_bytes(B, h, L, 64, _lib.MT_BF16) == tiles * 32768
    assert lib.mt_rga_stash_bytes(2, 2, 64, 128, _lib.MT_BF16) == 0           # head dim 128: fp32-math kernels, no stash
    assert lib.mt_rga_stash_bytes(2, 2, 64, 64, _lib.MT_F32) == 0
    ws = lib.mt_decode_run_workspace_bytes(32, 512, 390)
    assert ws >= 256 + 4 * 32 * 512 * 4 + 32 * 390 * 4 + 32 * (3 * 512 + 512 + 256) * 2 and ws % 256 == 0
    assert lib.mt_decode_run_workspace_bytes(0, 512, 390) == 0


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


# ---- events -> notes -> MIDI file ------------------------------------------------
def _parse_smf(data: bytes):
    """Minimal Standard MIDI File reader for the round-trip check: returns (division, tempo_us, program,
    note-ons [(pitch, velocity, tick)], note-offs [(pitch, tick)]) -- overlapping notes of one pitch (which the
    reference's decoder produces: a re-struck open note keeps its default length) make on/off PAIRING ambiguous in
    a MIDI stream, so the two event multisets are compared instead."""
    import struct
    assert data[:4] == b"MThd"
    hlen, fmt, ntrk, div = struct.unpack(">IHHH", data[4:14])
    pos, tracks = 8 + hlen, []
    for _ in range(ntrk):
        assert data[pos:pos + 4] == b"MTrk"
        n = struct.unpack(">I", data[pos + 4:pos + 8])[0]
        tracks.append(data[pos + 8:pos + 8 + n])
        pos += 8 + n
    assert pos == len(data) and fmt == 1 and ntrk == 2
    tempo, program, ons, offs = None, None, [], []
    for trk in tracks:
        i, t = 0, 0
        while i < len(trk):
            d = 0
            while True:
                b = trk[i]; i += 1
                d = (d << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            t += d
            st = trk[i]
            if st == 0xFF:
                typ, ln = trk[i + 1], trk[i + 2]
                if typ == 0x51:
                    tempo = int.from_bytes(trk[i + 3:i + 3 + ln], "big")
                i += 3 + ln
            elif st & 0xF0 == 0xC0:
                program = trk[i + 1]; i += 2
            elif st & 0xF0 == 0x90:
                pitch, vel = trk[i + 1], trk[i + 2]; i += 3
                if vel:
                    ons.append((pitch, vel, t))
                else:
                    offs.append((pitch, t))
            else:
                raise AssertionError(hex(st))
    return div, tempo, program, ons, offs


def test_events_to_notes_match_reference_golden():
    """musicgeneration_b200.sequence.events_to_notes == EventSeq.from_array(ids).to_note_seq() + the velocity
    rescaling of MT/utils.py:25-31, run UNMODIFIED in the build container (tests/golden/notes.npz): velocities
    and pitches equal, start / end times bit-identical (same float64 accumulation order)."""
    from musicgeneration_b200 import sequence as S
    z = np.load(os.path.join(ROOT, "tests", "golden", "notes.npz"))
    for name in ("random", "dense", "phrase"):
        got = S.events_to_notes(z["ids:" + name], velocity_scale=0.8)
        ref = z["notes:" + name]
        assert len(got) == len(ref) > 0, name
        arr = np.array([[n.velocity, n.pitch, n.start, n.end] for n in got], dtype=np.float64)
        assert (arr == ref).all(), name
    assert S.EVENT_DIM == 308 and S.events_to_notes([388, 389, 400, -1]) == []


def test_midi_file_round_trip(tmp_path):
    from musicgeneration_b200 import sequence as S
    z = np.load(os.path.join(ROOT, "tests", "golden", "notes.npz"))
    path = tmp_path / "out.mid"
    n = S.event_indeces_to_midi_file(z["ids:dense"], str(path))
    notes = S.events_to_notes(z["ids:dense"], 0.8)
    assert n == len(notes) == len(z["notes:dense"])
    div, tempo, program, ons, offs = _parse_smf(path.read_bytes())
    assert (div, tempo, program) == (220, 500000, 1)
    tick = lambda t: int(round(t * 220 * 2))
    assert sorted(ons) == sorted((m.pitch, m.velocity, tick(m.start)) for m in notes)
    assert sorted(offs) == sorted((m.pitch, tick(m.end)) for m in notes)
    assert S.decode_midi is S.event_indeces_to_midi_file
