"""GPU parity of the fused relative global attention (K1/K2) through the C ABI against the CPU
oracle's closed form (fp64) and against the reference-generated fixtures."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restate as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PATHS = {"simt": 1, "tc": 2}


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run_case(B, h, L, dh, max_seq, causal, pad, dtype, path, seed=0, scale=1.0, spill=True, io_dtype=None, do_scale=1.0,
             stash=False):
    """io_dtype: type of O / dO / dq / dk / dv when it differs from the q / k / v / E type (the mixed mode).
    stash: the training pair (the forward keeps its P tiles, the backward reads them)."""
    from musicgeneration_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    d = h * dh
    qkv = (torch.randn(B, L, 3, h, dh, generator=g) * scale).to(dtype)
    E = torch.randn(max_seq, dh, generator=g).to(dtype)
    io_dtype = io_dtype or dtype
    dO = (torch.randn(B, L, h, dh, generator=g) * do_scale).to(io_dtype)
    pad_keys = None
    if pad:
        pad_keys = torch.zeros(B, L, dtype=torch.bool)
        pad_keys[0, L // 2:L // 2 + 3] = True
        pad_keys[B - 1, L - 5:] = True
    # ---- oracle in fp64 on the (rounded) inputs
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).double().requires_grad_(True) for i in range(3)]
    E64 = E.double().requires_grad_(True)
    o_ref, lse_ref = O.rga_closed_form(q, k, v, E64, max_seq, causal, pad_keys)
    (o_ref * dO.permute(0, 2, 1, 3).double()).sum().backward()
    # ---- CUDA
    qkv_d = qkv.to(dev)
    Ed = E.to(dev)
    Od = torch.empty(B, L, h, dh, dtype=io_dtype, device=dev)
    lse = torch.empty(B, h, L, device=dev)
    strides = (L * 3 * d, 3 * d, dh)
    ostr = (L * d, d, dh)
    qd, kd, vd = qkv_d[:, :, 0], qkv_d[:, :, 1], qkv_d[:, :, 2]
    pk = pad_keys.to(torch.uint8).to(dev) if pad_keys is not None else None
    st = None
    if stash:
        st = ops.rga_stash_new(qd, Ed, Od, B, h, L, dh)
        assert st is not None
        st.fill_(0xFF)           # (NaN patterns: a tile the forward fails to write shows up)
    ops.rga_fwd(qd, kd, vd, strides, Ed, pk, Od, ostr, lse, B, h, L, dh, max_seq, causal, path=path, stash=st)
    dqkv = torch.zeros(B, L, 3, h, dh, dtype=io_dtype, device=dev)
    dE = torch.zeros(max_seq, dh, device=dev)
    delta = torch.empty(B, h, L, device=dev)
    ops.rga_bwd(qd, kd, vd, strides, Ed, pk, Od, dO.to(dev), ostr, lse, delta, dqkv[:, :, 0],
                dqkv[:, :, 1], dqkv[:, :, 2], dE, B, h, L, dh, max_seq, causal, path=path, spill=spill, stash=st)
    res = dict(
        o=rel(Od.float().cpu().permute(0, 2, 1, 3), o_ref.detach()),
        lse=float((lse.cpu().double() - lse_ref.detach()).abs().max()),
        dq=rel(dqkv[:, :, 0].float().cpu().permute(0, 2, 1, 3), q.grad),
        dk=rel(dqkv[:, :, 1].float().cpu().permute(0, 2, 1, 3), k.grad),
        dv=rel(dqkv[:, :, 2].float().cpu().permute(0, 2, 1, 3), v.grad),
        dE=rel(dE.cpu(), E64.grad))
    # weights kernel (eval-mode return)
    P = torch.empty(B, h, L, L, device=dev)
    ops.rga_weights(qd, kd, strides, Ed, pk, lse, P, B, h, L, dh, max_seq, causal)
    with torch.no_grad():
        i = torch.arange(L)[:, None]
        j = torch.arange(L)[None, :]
        qe = torch.einsum("bhld,md->bhlm", q, E64)
        idx = (max_seq - 1 - (i - j)).clamp(0, max_seq - 1)
        srel = torch.gather(qe, 3, idx.expand(B, h, L, L)) * (j <= i)
        s = (q @ k.transpose(-1, -2) + srel) / dh ** 0.5
        dead = torch.zeros(B, 1, L, L, dtype=torch.bool)
        if causal:
            dead = dead | (j > i)
        if pad_keys is not None:
            dead = dead | pad_keys[:, None, None, :]
        p_ref = torch.softmax(s.masked_fill(dead, float("-inf")), -1)
    res["P"] = float((P.cpu().double() - p_ref).abs().max())
    return res


@pytest.mark.parametrize("B,h,L,dh,max_seq,causal,pad", [
    (2, 2, 64, 64, 64, True, False),
    (2, 2, 48, 64, 64, True, False),        # L < max_seq, L not a tile multiple
    (1, 4, 200, 32, 256, True, True),
    (2, 1, 130, 128, 130, True, False),
    (2, 2, 96, 64, 96, False, False),       # generate(): mask=None, rel term only for j<=i
    (1, 2, 77, 64, 128, False, True),
    (1, 4, 512, 64, 512, True, False),
])
def test_rga_simt_fp32(B, h, L, dh, max_seq, causal, pad):
    r = run_case(B, h, L, dh, max_seq, causal, pad, torch.float32, PATHS["simt"], seed=L)
    assert r["o"] < 2e-6 and r["lse"] < 2e-5 and r["P"] < 2e-6, r
    assert max(r["dq"], r["dk"], r["dv"], r["dE"]) < 1e-5, r


def test_rga_simt_fp32_large_logits():
    # layer-0-like statistics: |q|,|k| ~ 13 -> logits in the hundreds
    r = run_case(1, 2, 256, 64, 256, True, False, torch.float32, PATHS["simt"], seed=5, scale=6.0)
    assert r["o"] < 1e-5 and r["P"] < 1e-4, r


@pytest.mark.parametrize("B,h,L,dh,max_seq,causal", [(2, 2, 128, 64, 128, True), (1, 2, 96, 64, 128, False)])
def test_rga_simt_bf16_io(B, h, L, dh, max_seq, causal):
    r = run_case(B, h, L, dh, max_seq, causal, False, torch.bfloat16, PATHS["simt"], seed=1)
    assert r["o"] < 4e-3 and max(r["dq"], r["dk"], r["dv"], r["dE"]) < 8e-3, r


def run_fwd_tc(B, h, L, dh, max_seq, causal, pad, dtype, seed=0, scale=1.0):
    """tcgen05 forward (path=2) vs the fp64 oracle on the same 16-bit-rounded inputs."""
    from musicgeneration_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    d = h * dh
    qkv = (torch.randn(B, L, 3, h, dh, generator=g) * scale).to(dtype)
    E = torch.randn(max_seq, dh, generator=g).to(dtype)
    pad_keys = None
    if pad:
        pad_keys = torch.zeros(B, L, dtype=torch.bool)
        pad_keys[0, L // 2:L // 2 + 3] = True
        pad_keys[B - 1, L - 5:] = True
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3)]
    o_ref, lse_ref = O.rga_closed_form(q, k, v, E.double(), max_seq, causal, pad_keys)
    qkv_d, Ed = qkv.to(dev), E.to(dev)
    Od = torch.zeros(B, L, h, dh, dtype=dtype, device=dev)
    lse = torch.zeros(B, h, L, device=dev)
    pk = pad_keys.to(torch.uint8).to(dev) if pad_keys is not None else None
    ops.rga_fwd(qkv_d[:, :, 0], qkv_d[:, :, 1], qkv_d[:, :, 2], (L * 3 * d, 3 * d, dh), Ed, pk, Od,
                (L * d, d, dh), lse, B, h, L, dh, max_seq, causal, path=2)
    torch.cuda.synchronize()
    o = Od.float().cpu().permute(0, 2, 1, 3)
    return dict(o=rel(o, o_ref), lse=float((lse.cpu().double() - lse_ref).abs().max()),
                omax=float((o.double() - o_ref).abs().max()))


@pytest.mark.parametrize("B,h,L,max_seq,causal,pad", [
    (1, 1, 128, 128, True, False),
    (2, 2, 256, 256, True, False),
    (2, 4, 512, 512, True, False),
    (1, 2, 200, 256, True, False),          # ragged L, L < max_seq
    (1, 2, 384, 512, True, True),           # max_seq - L not a tile multiple of the band, pads
    (2, 2, 256, 256, False, False),         # generate(): no mask
    (1, 2, 300, 300, False, True),
    (1, 8, 2048, 2048, True, False),        # BASELINE shape (one sequence)
])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_rga_fwd_tcgen05(B, h, L, max_seq, causal, pad, dtype):
    r = run_fwd_tc(B, h, L, 64, max_seq, causal, pad, dtype, seed=L + h)
    # operands are exact; P is rounded to 16 bits before P.V and the output is 16-bit
    tol = 6e-3 if dtype == torch.bfloat16 else 1.5e-3
    assert r["o"] < tol and r["lse"] < 2e-3, r


@pytest.mark.parametrize("B,h,L,max_seq,pad", [
    (1, 1, 128, 128, False),
    (2, 2, 64, 64, False),             # less than one tile (the parity fixtures' shape)
    (2, 2, 256, 256, False),
    (1, 2, 200, 256, False),           # ragged L < max_seq
    (1, 2, 384, 512, True),
    (2, 4, 512, 512, False),
    (1, 8, 1024, 2048, True),
])
@pytest.mark.parametrize("spill", [True, False, "stash"])
def test_rga_bwd_tcgen05(B, h, L, max_seq, pad, spill):
    """forward + backward both on the tcgen05 path; bf16 operands, fp64 oracle on the same inputs.
    spill: dS tiles written once by the dK/dV kernel and consumed by the dQ / dE kernels
    against every kernel recomputing them; "stash": the training pair -- the forward keeps its P tiles and the
    backward reads them (what the model's train step runs)."""
    r = run_case(B, h, L, 64, max_seq, True, pad, torch.bfloat16, PATHS["tc"], seed=L + 3 * h, spill=spill is not False,
                 stash=spill == "stash")
    assert r["o"] < 6e-3, r
    # P and dS are rounded to bf16 before the gradient GEMMs
    assert max(r["dq"], r["dk"], r["dv"]) < 1.2e-2 and r["dE"] < 1.2e-2, r


@pytest.mark.parametrize("B,h,L,max_seq,pad,scale", [
    (1, 1, 128, 128, False, 1.0),
    (2, 2, 64, 64, False, 1.0),
    (2, 2, 256, 256, False, 1.0),
    (1, 2, 200, 256, False, 1.0),           # ragged L < max_seq
    (1, 2, 384, 512, True, 1.0),
    (2, 4, 512, 512, False, 6.0),           # layer-0-like statistics: logits in the hundreds
    (1, 8, 1024, 2048, True, 3.0),
])
@pytest.mark.parametrize("stash", [False, True])
def test_rga_tcgen05_mixed_f16_qkv_bf16_io(B, h, L, max_seq, pad, scale, stash):
    """The first encoder layer's mode (MT_F16_BF16): f16 q / k / v / E operands, bf16 O / dO / dq / dk / dv.  The
    tensor cores take one operand format per product, so the backward runs on f16(2^12 dO), an f16 P and an f16
    2^12 dS (rga_tc_bwd.cu); dO has the magnitude of real activation gradients (the scaled copy overflows f16
    beyond |dO| |v| ~ 16).  fp64 oracle on the same inputs."""
    r = run_case(B, h, L, 64, max_seq, True, pad, torch.float16, PATHS["tc"], seed=L + 3 * h, scale=scale,
                 io_dtype=torch.bfloat16, do_scale=1e-3, stash=stash)
    # O is rounded to bf16 on the way out (as in the bf16 mode); the logits see f16 operands and an f16 P
    assert r["o"] < 6e-3 and r["lse"] < 2e-3 * max(1.0, scale * scale), r
    assert max(r["dq"], r["dk"], r["dv"]) < 1.2e-2 and r["dE"] < 1.2e-2, r


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_rga_bwd_stash_with_spilled_ds_variant(monkeypatch, dtype):
    """MT_RGA_STASH_SPILL=1: the intermediate training variant kept for comparison (dK/dV from the P stash, dS spilled
    once, dQ / dE by the fused consumer of rga_tc_bwd3.cu) against the fp64 closed form, bf16 and the f16 layer-0 mode."""
    monkeypatch.setenv("MT_RGA_STASH_SPILL", "1")
    io = torch.bfloat16 if dtype == torch.float16 else None
    r = run_case(2, 4, 512, 64, 512, True, False, dtype, PATHS["tc"], seed=9, io_dtype=io,
                 do_scale=1e-3 if io is not None else 1.0, stash=True)
    assert r["o"] < 6e-3, r
    assert max(r["dq"], r["dk"], r["dv"]) < 1.2e-2 and r["dE"] < 1.2e-2, r


def test_rga_mixed_mode_needs_the_tensor_core_path_and_workspace():
    from musicgeneration_b200 import ops
    dev = torch.device("cuda:0")
    B, h, L, dh = 1, 1, 128, 64
    qkv = torch.zeros(B, L, 3, h, dh, dtype=torch.float16, device=dev)
    E = torch.zeros(L, dh, dtype=torch.float16, device=dev)
    Od = torch.zeros(B, L, h, dh, dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(B, h, L, device=dev)
    with pytest.raises(RuntimeError, match="tcgen05 path only"):
        ops.rga_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], (L * 192, 192, 64), E, None, Od, (L * 64, 64, 64), lse,
                    B, h, L, dh, L, True, path=1)
    dq = torch.zeros_like(qkv, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="workspace"):
        ops.rga_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], (L * 192, 192, 64), E, None, Od, Od, (L * 64, 64, 64),
                    lse, torch.zeros(B, h, L, device=dev), dq[:, :, 0], dq[:, :, 1], dq[:, :, 2],
                    torch.zeros(L, dh, device=dev), B, h, L, dh, L, True, path=2, spill=False)


def test_rga_bwd_tcgen05_matches_simt_backward():
    """Same bf16 inputs through both backward implementations (SIMT fp32 math vs tensor cores)."""
    from musicgeneration_b200 import ops
    dev = torch.device("cuda:0")
    B, h, L, dh, max_seq = 2, 4, 640, 64, 1024
    d = h * dh
    g = torch.Generator().manual_seed(77)
    qkv = torch.randn(B, L, 3, h, dh, generator=g).to(torch.bfloat16).to(dev)
    E = torch.randn(max_seq, dh, generator=g).to(torch.bfloat16).to(dev)
    dO = torch.randn(B, L, h, dh, generator=g).to(torch.bfloat16).to(dev)
    strides, ostr = (L * 3 * d, 3 * d, dh), (L * d, d, dh)
    outs = {}
    for name, path, spill in (("simt", PATHS["simt"], False), ("tc", PATHS["tc"], True), ("tc_recompute", PATHS["tc"], False),
                              ("tc_stash", PATHS["tc"], True)):
        Od = torch.empty(B, L, h, dh, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(B, h, L, device=dev)
        st = ops.rga_stash_new(qkv[:, :, 0], E, Od, B, h, L, dh) if name == "tc_stash" else None
        ops.rga_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, ostr, lse, B, h, L, dh,
                    max_seq, True, path=path, stash=st)
        dqkv = torch.zeros(B, L, 3, h, dh, dtype=torch.bfloat16, device=dev)
        dE = torch.zeros(max_seq, dh, device=dev)
        delta = torch.empty(B, h, L, device=dev)
        ops.rga_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, dO, ostr, lse, delta,
                    dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], dE, B, h, L, dh, max_seq, True, path=path, spill=spill,
                    stash=st)
        outs[name] = (dqkv.float().cpu(), dE.cpu())
    assert rel(outs["tc_stash"][0], outs["simt"][0]) < 1.2e-2
    assert rel(outs["tc_stash"][1], outs["simt"][1]) < 1.2e-2
    assert rel(outs["tc"][0], outs["simt"][0]) < 1.2e-2
    assert rel(outs["tc"][1], outs["simt"][1]) < 1.2e-2
    assert float(outs["tc"][1][:max_seq - L].abs().max()) == 0.0      # only rows max_seq-L.. get gradient
    # the two tensor-core variants see the same bf16 dS tiles: equal up to fp32 summation order / bf16 rounding of dq
    assert rel(outs["tc"][0], outs["tc_recompute"][0]) < 2e-3
    assert rel(outs["tc"][1], outs["tc_recompute"][1]) < 1e-4


def test_rga_tcgen05_long_context_against_oracle():
    """BASELINE config C's context (L = max_seq = 4096, 32 key tiles per row) on one head against the
    fp64 closed form; ragged last tile included (L = 4000 < max_seq)."""
    r = run_case(1, 1, 4000, 64, 4096, True, False, torch.bfloat16, PATHS["tc"], seed=11)
    assert r["o"] < 6e-3 and r["lse"] < 2e-3, r
    assert max(r["dq"], r["dk"], r["dv"]) < 1.2e-2 and r["dE"] < 1.2e-2, r


@pytest.mark.parametrize("B,h,L", [(1, 12, 4096), (1, 2, 8192), (16, 8, 2048), (16, 6, 2048), (16, 8, 1024)])
def test_rga_tcgen05_long_context_matches_simt(B, h, L):
    """Config C's attention shape (12 heads, L = 4096), the top of the microbench sweep (L = 8192) and config B's
    full per-layer shape (16 x 8 heads x 2048): the last three are large enough for the launchers to put several
    heads on one CTA (4 + 4, 4 + 2 with a remainder, 2 per CTA) -- the small cases elsewhere never do.
    the tensor-core kernels (dS-spill backward) against the fp32-math SIMT kernels on the same bf16
    inputs (the fp64 oracle would need several L x L fp64 matrices per head on the host)."""
    from musicgeneration_b200 import ops
    dev = torch.device("cuda:0")
    dh, max_seq = 64, L
    d = h * dh
    g = torch.Generator().manual_seed(L + h)
    qkv = torch.randn(B, L, 3, h, dh, generator=g).to(torch.bfloat16).to(dev)
    E = torch.randn(max_seq, dh, generator=g).to(torch.bfloat16).to(dev)
    dO = torch.randn(B, L, h, dh, generator=g).to(torch.bfloat16).to(dev)
    strides, ostr = (L * 3 * d, 3 * d, dh), (L * d, d, dh)
    outs = {}
    for name, path in list(PATHS.items()) + [("tc_stash", PATHS["tc"])]:
        Od = torch.empty(B, L, h, dh, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(B, h, L, device=dev)
        st = ops.rga_stash_new(qkv[:, :, 0], E, Od, B, h, L, dh) if name == "tc_stash" else None
        ops.rga_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, ostr, lse, B, h, L, dh,
                    max_seq, True, path=path, stash=st)
        dqkv = torch.zeros(B, L, 3, h, dh, dtype=torch.bfloat16, device=dev)
        dE = torch.zeros(max_seq, dh, device=dev)
        delta = torch.empty(B, h, L, device=dev)
        ops.rga_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, dO, ostr, lse, delta,
                    dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], dE, B, h, L, dh, max_seq, True, path=path, stash=st)
        outs[name] = (Od.float().cpu(), lse.cpu(), dqkv.float().cpu(), dE.cpu())
        del st
    # the training pair (P stash) against the fp32-math kernels
    assert rel(outs["tc_stash"][0], outs["simt"][0]) < 6e-3
    assert rel(outs["tc_stash"][2], outs["simt"][2]) < 1.2e-2
    assert rel(outs["tc_stash"][3], outs["simt"][3]) < 1.2e-2
    assert rel(outs["tc"][0], outs["simt"][0]) < 6e-3
    assert float((outs["tc"][1] - outs["simt"][1]).abs().max()) < 2e-3
    assert rel(outs["tc"][2], outs["simt"][2]) < 1.2e-2
    assert rel(outs["tc"][3], outs["simt"][3]) < 1.2e-2


def test_rga_fwd_tcgen05_large_logits():
    r = run_fwd_tc(1, 2, 256, 64, 256, True, False, torch.bfloat16, seed=5, scale=6.0)
    assert r["o"] < 8e-3 and r["lse"] < 2e-2, r


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
@pytest.mark.parametrize("tag", ["none", "causal"])
def test_rga_module_against_reference_fixture(case, tag):
    import musicgeneration_b200 as mtb
    z = np.load(os.path.join(GOLD, "rga_small.npz"))
    h, d, max_seq, L = z["meta"]["abcd".index(case)].tolist()
    dev = torch.device("cuda:0")
    rga = mtb.RelativeGlobalAttention(h=h, d=d, max_seq=max_seq).to(dev)
    sd = {k[len(case) + 3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(case + ":p:")}
    rga.load_state_dict(sd, strict=True)
    rga.need_weights = True
    x = torch.from_numpy(z[case + ":x"]).to(dev).requires_grad_(True)
    ar = torch.arange(L, device=dev)
    mask = (ar[None, :] > ar[:, None])[None, None] if tag == "causal" else None
    out, w = rga([x, x, x], mask)
    np.testing.assert_allclose(out.detach().cpu().numpy(), z[f"{case}:{tag}:out"], atol=5e-5)
    if case == "a":
        np.testing.assert_allclose(w.cpu().numpy(), z[f"{case}:{tag}:w"], atol=2e-5)
    wgt = torch.cos(torch.arange(out.numel(), dtype=torch.float32)).reshape(out.shape).to(dev)
    (out * wgt).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), z[f"{case}:{tag}:dx"], atol=1e-3, rtol=2e-4)
    for k, p in rga.named_parameters():
        g = z[f"{case}:{tag}:g:{k}"]
        assert np.abs(p.grad.cpu().numpy() - g).max() <= 2e-5 + 3e-4 * np.abs(g).max(), k


@pytest.mark.parametrize("path,dtype", [("simt", torch.float32), ("tc", torch.bfloat16)])
def test_rga_fully_masked_rows_are_zero_by_definition(path, dtype):
    """The documented deviation: a query row whose every visible key is a pad
    token -- only possible when a sequence STARTS with pads -- gets output 0 and LSE 0 here.  The reference adds -1e9 to
    every logit of such a row (MT/layers.py:93-95), fp32 absorbs the logits, and softmax returns the uniform
    distribution over ALL L keys, future ones included: the mean of V.  Every other row of the same call must still agree
    with the closed form."""
    from musicgeneration_b200 import ops
    dev = torch.device("cuda:0")
    B, h, L, dh, max_seq, npad = 2, 2, 160, 64, 160, 3
    d = h * dh
    g = torch.Generator().manual_seed(21)
    qkv = torch.randn(B, L, 3, h, dh, generator=g).to(dtype)
    E = torch.randn(max_seq, dh, generator=g).to(dtype)
    pad_keys = torch.zeros(B, L, dtype=torch.bool)
    pad_keys[0, :npad] = True                      # sequence 0 starts with three pad tokens
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3)]
    o_ref, lse_ref = O.rga_closed_form(q, k, v, E.double(), max_seq, True, pad_keys)      # NaN on the dead rows
    qkv_d, Ed = qkv.to(dev), E.to(dev)
    Od = torch.full((B, L, h, dh), 7.0, dtype=dtype, device=dev)
    lse = torch.full((B, h, L), 7.0, device=dev)
    ops.rga_fwd(qkv_d[:, :, 0], qkv_d[:, :, 1], qkv_d[:, :, 2], (L * 3 * d, 3 * d, dh), Ed, pad_keys.to(torch.uint8).to(dev), Od,
                (L * d, d, dh), lse, B, h, L, dh, max_seq, True, path=PATHS[path])
    o = Od.float().cpu().permute(0, 2, 1, 3)
    assert float(o[0, :, :npad].abs().max()) == 0.0 and float(lse.cpu()[0, :, :npad].abs().max()) == 0.0
    live = torch.ones(B, h, L, dtype=torch.bool)
    live[0, :, :npad] = False
    tol = 2e-6 if dtype == torch.float32 else 6e-3
    assert rel(o[live], o_ref[live]) < tol
    # what the reference returns on those rows: softmax of fp32(s - 1e9) = uniform over all keys -> mean of V
    s_row = torch.full((L,), -1e9, dtype=torch.float32) + torch.randn(L, generator=g)
    assert torch.allclose(torch.softmax(s_row, -1), torch.full((L,), 1.0 / L), atol=1e-6)


def test_rga_argument_errors():
    from musicgeneration_b200 import ops
    dev = torch.device("cuda:0")
    q = torch.zeros(1, 8, 3, 1, 48, device=dev)
    E = torch.zeros(8, 48, device=dev)
    O_ = torch.zeros(1, 8, 1, 48, device=dev)
    lse = torch.zeros(1, 1, 8, device=dev)
    with pytest.raises(RuntimeError, match="head dim"):
        ops.rga_fwd(q[:, :, 0], q[:, :, 1], q[:, :, 2], (8 * 144, 144, 48), E, None, O_, (8 * 48, 48, 48),
                    lse, 1, 1, 8, 48, 8, True, path=1)
    with pytest.raises(RuntimeError, match="bad shape"):
        ops.rga_fwd(q[:, :, 0], q[:, :, 1], q[:, :, 2], (8 * 144, 144, 48), E, None, O_, (8 * 48, 48, 48),
                    lse, 1, 1, 8, 64, 4, True, path=1)
