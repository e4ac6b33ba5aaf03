"""GPU parity of the individual C-ABI ops against CPU restatements (oracle / plain torch fp32
or fp64 on the host) on seeded inputs.  Everything here calls libmt_b200 through ctypes."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restate as O  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _ops():
    from musicgeneration_b200 import ops
    return ops


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 390, 256), (77, 50, 33), (4096, 96, 128),
                                   (64, 64, 4096)])
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_simt(dev, M, N, K, tA, tB, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(dtype)
    B = torch.randn((N, K) if tB else (K, N), generator=g).to(dtype)
    bias = torch.randn(N, generator=g)
    add = torch.randn(M, N, generator=g)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    Ad, Bd = A.to(dev), B.to(dev)
    C = torch.empty(M, N, dtype=torch.float32, device=dev)
    ops.gemm(Ad, Bd, C, M, N, K, Ad.stride(0), Bd.stride(0), N, tA, tB, path=1)
    tol = 2e-6 if dtype == torch.float32 else 2e-6   # inputs are exactly representable either way
    assert rel(C.cpu(), ref) < tol
    # bias + addend + relu, bf16 output when inputs are bf16
    C2 = torch.empty(M, N, dtype=dtype, device=dev)
    ops.gemm(Ad, Bd, C2, M, N, K, Ad.stride(0), Bd.stride(0), N, tA, tB, bias=bias.to(dev),
             addend=add.to(dev), relu=True, path=1)
    ref2 = torch.relu(ref + bias.double() + add.double())
    assert rel(C2.float().cpu(), ref2) < (1e-5 if dtype == torch.float32 else 4e-3)
    # relu mask epilogue
    aux = torch.randn(M, N, generator=g).to(dtype)
    C3 = torch.empty(M, N, dtype=torch.float32, device=dev)
    ops.gemm(Ad, Bd, C3, M, N, K, Ad.stride(0), Bd.stride(0), N, tA, tB, aux=aux.to(dev), relu_mask=True,
             path=1)
    ref3 = ref * (aux.double() > 0)
    assert rel(C3.cpu(), ref3) < 2e-6


def test_gemm_argument_errors(dev):
    ops = _ops()
    A = torch.zeros(8, 8, device=dev)
    with pytest.raises(RuntimeError, match="leading dimension"):
        ops.gemm(A, A, A, 8, 8, 8, 4, 8, 8, False, True, path=1)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.gemm(A.cpu(), A, A, 8, 8, 8, 8, 8, 8, False, True, path=1)


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [64, 128, 192, 256, 512, 768])
@pytest.mark.parametrize("lp", [False, True])
def test_add_ln_fwd_bwd(dev, d, lp):
    ops = _ops()
    T = 131
    g = torch.Generator().manual_seed(d)
    a = torch.randn(T, d, generator=g) * 2
    r = torch.randn(T, d, generator=g) * 3 + 0.5
    gamma = torch.randn(d, generator=g)
    beta = torch.randn(d, generator=g)
    dy = torch.randn(T, d, generator=g)
    a_in = a.to(torch.bfloat16) if lp else a
    a64 = a_in.double().requires_grad_(True)
    r64 = r.double().requires_grad_(True)
    g64 = gamma.double().requires_grad_(True)
    b64 = beta.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(a64 + r64, (d,), g64, b64, 1e-6)
    ref.backward(dy.double())
    out = torch.empty(T, d, device=dev)
    out_lp = torch.empty(T, d, dtype=torch.bfloat16, device=dev) if lp else None
    mean = torch.empty(T, device=dev)
    rstd = torch.empty(T, device=dev)
    ad, rd, gd, bd = a_in.to(dev), r.to(dev), gamma.to(dev), beta.to(dev)
    ops.add_ln_fwd(ad, rd, gd, bd, out, out_lp, mean, rstd, 1e-6, 0.0, 0, 0)
    assert rel(out.cpu(), ref.detach()) < 2e-6
    if lp:
        assert rel(out_lp.float().cpu(), ref.detach()) < 4e-3
    dyd = dy.to(dev)
    dz = torch.empty(T, d, device=dev)
    da = torch.empty(T, d, dtype=torch.bfloat16 if lp else torch.float32, device=dev)
    dgam = torch.empty(d, device=dev)
    dbet = torch.empty(d, device=dev)
    dbias = torch.empty(d, device=dev)
    ops.add_ln_bwd(dyd, ad, rd, gd, mean, rstd, dz, da, dgam, dbet, 0.0, 0, 0, dbias=dbias)
    # colsum(da) = bias gradient of the linear that produced `a` (sum of the stored, rounded values)
    assert rel(dbias.cpu(), da.float().sum(0).cpu()) < 2e-5
    assert rel(dz.cpu(), r64.grad) < 5e-6
    assert rel(da.float().cpu(), a64.grad) < (4e-3 if lp else 5e-6)
    assert rel(dgam.cpu(), g64.grad) < 5e-6
    assert rel(dbet.cpu(), b64.grad) < 5e-6
    # in-place form used by the engine (dz aliases dout)
    dy2 = dyd.clone()
    ops.add_ln_bwd(dy2, ad, rd, gd, mean, rstd, dy2, da, dgam, dbet, 0.0, 0, 0)
    assert torch.equal(dy2, dz)


def test_dropout_is_consistent_between_fwd_and_bwd(dev):
    ops = _ops()
    T, d, p = 257, 256, 0.2
    a = torch.ones(T, d, device=dev)
    r = torch.zeros(T, d, device=dev)
    gamma = torch.ones(d, device=dev)
    beta = torch.zeros(d, device=dev)
    out = torch.empty(T, d, device=dev)
    mean = torch.empty(T, device=dev)
    rstd = torch.empty(T, device=dev)
    ops.add_ln_fwd(a, r, gamma, beta, out, None, mean, rstd, 1e-6, p, 1234, 5)
    # z = mask/(1-p): LN output sign tells kept (positive) from dropped (negative)
    kept_fwd = out > 0
    frac = kept_fwd.float().mean().item()
    assert abs(frac - (1 - p)) < 0.01
    dz = torch.empty(T, d, device=dev)
    da = torch.empty(T, d, device=dev)
    dg = torch.empty(d, device=dev)
    db = torch.empty(d, device=dev)
    dy = torch.randn(T, d, device=dev)
    ops.add_ln_bwd(dy, a, r, gamma, mean, rstd, dz, da, dg, db, p, 1234, 5)
    kept_bwd = da != 0
    assert (kept_bwd == kept_fwd).float().mean().item() > 0.9999
    assert torch.allclose(da[kept_fwd], dz[kept_fwd] / (1 - p), rtol=1e-6, atol=1e-7)
    # a different site or seed gives a different mask
    out2 = torch.empty(T, d, device=dev)
    ops.add_ln_fwd(a, r, gamma, beta, out2, None, mean, rstd, 1e-6, p, 1234, 6)
    assert ((out2 > 0) != kept_fwd).float().mean().item() > 0.2


# ---------------------------------------------------------------------------------------
def test_embed_pos_fwd_bwd(dev):
    ops = _ops()
    from musicgeneration_b200.layers import sinusoid_table
    B, L, d, V = 3, 40, 128, 50
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, V, (B, L), generator=g, dtype=torch.int32)
    emb = torch.randn(V, d, generator=g)
    pe = torch.from_numpy(sinusoid_table(64, d)[0].astype(np.float32))
    assert (pe.numpy() == O.sinusoid_table(64, d).astype(np.float32)).all()
    out = torch.empty(B * L, d, device=dev)
    out_lp = torch.empty(B * L, d, dtype=torch.bfloat16, device=dev)
    ops.embed_pos_fwd(ids.to(dev), emb.to(dev), pe.to(dev), out, out_lp, 0, math.sqrt(d), 0.0, 0, 0)
    ref = emb[ids.long()] * math.sqrt(d) + pe[None, :L]
    assert torch.equal(out.cpu().view(B, L, d), ref)
    assert rel(out_lp.float().cpu().view(B, L, d), ref) < 4e-3
    # position offset (decode)
    out1 = torch.empty(B, d, device=dev)
    ops.embed_pos_fwd(ids[:, 7:8].contiguous().to(dev), emb.to(dev), pe.to(dev), out1, None, 7,
                      math.sqrt(d), 0.0, 0, 0)
    assert torch.equal(out1.cpu(), ref[:, 7])
    dy = torch.randn(B * L, d, generator=g)
    demb = torch.zeros(V, d, device=dev)
    ops.embed_pos_bwd(ids.to(dev), dy.to(dev), demb, math.sqrt(d), 0.0, 0, 0)
    refg = torch.zeros(V, d, dtype=torch.float64)
    refg.index_add_(0, ids.reshape(-1).long(), dy.double() * math.sqrt(d))
    assert rel(demb.cpu(), refg) < 1e-6


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("V", [96, 390, 337])
def test_smooth_ce(dev, V):
    ops = _ops()
    T, pad, eps = 203, V - 2, 0.1
    g = torch.Generator().manual_seed(V)
    z = torch.randn(T, V, generator=g) * 3
    y = torch.randint(0, pad, (T,), generator=g, dtype=torch.int32)
    y[5:20] = pad
    z64 = z.double().requires_grad_(True)
    ref = O.smooth_ce(z64, y, eps, V, pad)
    (ref * 0.7).backward()
    zd, yd = z.to(dev), y.to(dev)
    row_ws = torch.empty(3, T, device=dev)
    am = torch.empty(T, dtype=torch.int32, device=dev)
    sums = torch.empty(4, device=dev)
    ops.smooth_ce_fwd(zd, yd, row_ws, am, sums, eps, pad)
    s = sums.cpu()
    assert abs(float(s[3]) - float(ref)) < 2e-6 * max(1, abs(float(ref)))
    assert int(s[1]) == int((y != pad).sum())
    assert (am.cpu() == z.argmax(-1).int()).all()
    assert int(s[2]) == int((z.argmax(-1).int() == y).sum())
    dz = torch.empty(T, V, device=dev)
    ops.smooth_ce_bwd(zd, yd, row_ws, sums, torch.tensor([0.7], device=dev), dz, eps, pad)
    assert rel(dz.cpu(), z64.grad) < 5e-6
    assert float(dz[5:20].abs().max()) == 0.0


def test_colsum_cast_transpose(dev):
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    X = torch.randn(5000, 200, generator=g)
    for dt in (torch.float32, torch.bfloat16):
        Xd = X.to(dt).to(dev)
        out = torch.empty(150, device=dev)
        ops.colsum(Xd[:, 20:170], out, 5000, 150, 200)
        assert rel(out.cpu(), X.to(dt).double()[:, 20:170].sum(0)) < 2e-6
    small = torch.empty(200, device=dev)
    ops.colsum(X[:100].contiguous().to(dev), small, 100, 200, 200)
    assert rel(small.cpu(), X[:100].double().sum(0)) < 2e-6
    y = torch.empty(5000 * 200 - 3, dtype=torch.bfloat16, device=dev)
    ops.cast(X.to(dev), y, y.numel())
    assert torch.equal(y.cpu(), X.reshape(-1)[:-3].to(torch.bfloat16))
    t = torch.empty(200, 5000, dtype=torch.bfloat16, device=dev)
    ops.transpose_cast(X.to(dev), t, 5000, 200)
    assert torch.equal(t.cpu(), X.t().to(torch.bfloat16))


def test_adam_matches_torch(dev):
    ops = _ops()
    n = 10007
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=0.0, betas=(0.9, 0.98), eps=1e-9)
    p = p0.to(dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    shadow = torch.empty(n, dtype=torch.bfloat16, device=dev)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g)
        lr = O.noam_rate(step, 256)
        for grp in opt.param_groups:
            grp["lr"] = lr
        ref_p.grad = grad.clone()
        opt.step()
        ops.adam_step(p, (grad * 4).to(dev), m, v, shadow, lr, 0.9, 0.98, 1e-9, step, 0.25)
    assert rel(p.cpu(), ref_p.detach()) < 2e-6
    assert torch.equal(shadow.cpu(), p.cpu().to(torch.bfloat16))


def test_sampler(dev):
    ops = _ops()
    B, V = 64, 390
    g = torch.Generator().manual_seed(4)
    z = torch.randn(B, V, generator=g) * 2
    u = torch.rand(B, generator=g)
    out = torch.empty(B, dtype=torch.int32, device=dev)
    ops.sample(z.to(dev), None, out, 1.0, 0, True)
    assert (out.cpu() == z.argmax(-1).int()).all()
    for T_, k in ((1.0, 0), (0.8, 32), (1.3, 5), (1.0, 1)):
        ops.sample(z.to(dev), u.to(dev), out, T_, k, False)
        ref = O.sample_topk_from_uniform(z, u, T_, k)
        agree = (out.cpu().long() == ref).float().mean().item()
        assert agree >= 0.97, (T_, k, agree)      # a draw within 1 ulp of a CDF step may differ
        if k:
            topk = torch.topk(z, k, -1).indices
            assert all(int(out[b]) in topk[b].tolist() for b in range(B))
    # statistics: frequencies follow softmax over the kept ids
    zz = torch.tensor([[0.0, 3.0, 1.0, 2.0, -1.0]]).expand(20000, -1).contiguous()
    uu = torch.rand(20000, generator=g)
    o2 = torch.empty(20000, dtype=torch.int32, device=dev)
    ops.sample(zz.to(dev), uu.to(dev), o2, 0.7, 3, False)
    pr = torch.softmax(zz[0, [1, 2, 3]] / 0.7, 0)
    freq = torch.stack([(o2.cpu() == i).float().mean() for i in (1, 2, 3)])
    assert set(o2.cpu().tolist()) <= {1, 2, 3}
    assert (freq - pr).abs().max() < 0.015


# ---------------------------------------------------------------------------------------
# tcgen05 GEMM (path=2): same contract as the SIMT kernel, bf16 operands
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (384, 512, 512), (300, 390, 512),
                                   (4096, 1536, 512), (200, 136, 200), (1000, 256, 72)])
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False)])
def test_gemm_tc(dev, M, N, K, tA, tB):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N * 3 + K * 5 + tA * 11 + tB * 13)

    def pad8(n):
        return (n + 7) // 8 * 8
    # leading dimensions padded to a multiple of 8 elements (TMA needs 16-byte row pitch)
    a_shape = (K, M) if tA else (M, K)
    b_shape = (N, K) if tB else (K, N)
    A_full = torch.randn(a_shape[0], pad8(a_shape[1]), generator=g).to(torch.bfloat16)
    B_full = torch.randn(b_shape[0], pad8(b_shape[1]), generator=g).to(torch.bfloat16)
    A = A_full[:, :a_shape[1]]
    B = B_full[:, :b_shape[1]]
    bias = torch.randn(N, generator=g)
    ldc = pad8(N)
    add = torch.randn(M, ldc, generator=g)
    aux = torch.randn(M, ldc, generator=g).to(torch.bfloat16)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    Ad, Bd = A_full.to(dev), B_full.to(dev)
    C = torch.full((M, ldc), 7.0, dtype=torch.float32, device=dev)
    ops.gemm(Ad, Bd, C, M, N, K, Ad.stride(0), Bd.stride(0), ldc, tA, tB, path=2)
    assert rel(C[:, :N].cpu(), ref) < 2e-6, rel(C[:, :N].cpu(), ref)
    if ldc > N:
        assert float((C[:, N:] - 7.0).abs().max()) == 0.0      # nothing written past N
    C2 = torch.zeros((M, ldc), dtype=torch.bfloat16, device=dev)
    ops.gemm(Ad, Bd, C2, M, N, K, Ad.stride(0), Bd.stride(0), ldc, tA, tB, bias=bias.to(dev),
             addend=add.to(dev), relu=True, path=2)
    ref2 = torch.relu(ref + bias.double() + add[:, :N].double())
    assert rel(C2[:, :N].float().cpu(), ref2) < 4e-3
    C3 = torch.zeros((M, ldc), dtype=torch.float32, device=dev)
    ops.gemm(Ad, Bd, C3, M, N, K, Ad.stride(0), Bd.stride(0), ldc, tA, tB, aux=aux.to(dev), relu_mask=True,
             path=2)
    assert rel(C3[:, :N].cpu(), ref * (aux[:, :N].double() > 0)) < 2e-6
    # unpadded ldc (scalar store path), e.g. the logits [T, 390]
    C4 = torch.empty((M, N), dtype=torch.float32, device=dev)
    ops.gemm(Ad, Bd, C4, M, N, K, Ad.stride(0), Bd.stride(0), N, tA, tB, bias=bias.to(dev), path=2)
    assert rel(C4.cpu(), ref + bias.double()) < 2e-6


@pytest.mark.parametrize("M,N,K", [(256, 1536, 512), (300, 392, 200)])
def test_gemm_tc_f16_operands_and_output(dev, M, N, K):
    """f16 x f16 -> f16 (+bias): the first encoder layer's QKV projection in the bf16 mode."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 8).to(torch.float16)
    W = torch.randn(N, K, generator=g).to(torch.float16) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref = A.double() @ W.double().t() + bias.double()
    C = torch.zeros((M, N), dtype=torch.float16, device=dev)
    ops.gemm(A.to(dev), W.to(dev), C, M, N, K, K, K, N, False, True, bias=bias.to(dev), path=2)
    assert rel(C.float().cpu(), ref) < 6e-4            # 11-bit output rounding
    C32 = torch.zeros((M, N), dtype=torch.float32, device=dev)
    ops.gemm(A.to(dev), W.to(dev), C32, M, N, K, K, K, N, False, True, bias=bias.to(dev), path=2)
    assert rel(C32.cpu(), ref) < 2e-6


def test_gemm_tc_split_k_weight_gradient_shape(dev):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    T, N, K = 8192, 512, 256                    # dW[N,K] = dY[T,N]^T . X[T,K]
    dY = torch.randn(T, N, generator=g).to(torch.bfloat16)
    X = torch.randn(T, K, generator=g).to(torch.bfloat16)
    ref = dY.double().t() @ X.double()
    dW = torch.empty(N, K, dtype=torch.float32, device=dev)
    ops.gemm(dY.to(dev), X.to(dev), dW, N, K, T, N, K, K, True, False, path=2)
    assert rel(dW.cpu(), ref) < 3e-6
    dW2 = torch.empty(N, K, dtype=torch.float32, device=dev)
    ops.gemm(dY.to(dev), X.to(dev), dW2, N, K, T, N, K, K, True, False, path=2)
    assert torch.equal(dW, dW2)                 # deterministic split-K


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(32, 1536, 512), (1, 390, 512), (17, 512, 256), (64, 256, 512), (33, 390, 64),
                                   (48, 8, 1024)])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_gemm_skinny_decode_shapes(dev, M, N, K, out_bf16):
    """Decode-sized x . W^T (+bias, ReLU): the auto path takes the mma.sync strip kernel; compared with
    an fp64 product of the same bf16 operands."""
    ops = _ops()
    g = torch.Generator().manual_seed(M * 5 + N + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = torch.randn(N, K, generator=g).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ W.double().t() + bias.double()
    Ad, Wd = A.to(dev), W.to(dev)
    odt = torch.bfloat16 if out_bf16 else torch.float32
    C = torch.full((M, N), 7.0, dtype=odt, device=dev)
    ops.gemm(Ad, Wd, C, M, N, K, K, K, N, False, True, bias=bias.to(dev))
    assert rel(C.float().cpu(), ref) < (4e-3 if out_bf16 else 2e-6)
    C2 = torch.empty((M, N), dtype=odt, device=dev)
    ops.gemm(Ad, Wd, C2, M, N, K, K, K, N, False, True, bias=bias.to(dev), relu=True)
    assert rel(C2.float().cpu(), torch.relu(ref)) < (4e-3 if out_bf16 else 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,odt", [(32, 1536, 512, torch.float16), (32, 512, 512, torch.float32), (3, 768, 256, torch.float16)])
def test_gemm_skinny_f16(dev, M, N, K, odt):
    """f16 operands (and output) of the decode step's first layer."""
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, generator=g) * 8).to(torch.float16)
    W = torch.randn(N, K, generator=g).to(torch.float16) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref = A.double() @ W.double().t() + bias.double()
    C = torch.full((M, N), 7.0, dtype=odt, device=dev)
    ops.gemm(A.to(dev), W.to(dev), C, M, N, K, K, K, N, False, True, bias=bias.to(dev))
    assert rel(C.float().cpu(), ref) < (6e-4 if odt == torch.float16 else 2e-6)


@pytest.mark.gpu
def test_scalar_readback_order_and_values():
    """utils.ScalarReadback returns every pushed value, in order, lagging by at most depth - 1."""
    from musicgeneration_b200.utils import ScalarReadback
    rb = ScalarReadback(depth=2)
    got = []
    for i in range(7):
        rb.push(torch.full((), float(i) * 1.5, device="cuda"))
        if len(rb) == 2:
            got.append(rb.pop())
    got += rb.drain()
    assert got == [i * 1.5 for i in range(7)]
    with pytest.raises(RuntimeError):
        rb.pop()
    rb.push(torch.zeros((), device="cuda"))
    rb.push(torch.zeros((), device="cuda"))
    with pytest.raises(RuntimeError):
        rb.push(torch.zeros((), device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,ldm", [(1536, 512, 32768, 1536), (256, 512, 4096, 256), (390, 512, 2048, 392),
                                       (128, 128, 512, 128), (512, 256, 8192, 512), (200, 384, 3000, 208)])
def test_wgrad_bias_one_pass(dev, M, N, K, ldm):
    """mt_wgrad_bias: dW = dy^T x and db = colsum(dy) from one pass (extra N = 16 product against a ones tile).
    Reference: fp32 matmul / column sum of the same bf16-rounded operands; fp32 accumulation, so 1e-5 of the
    largest magnitude (the summation order differs)."""
    from musicgeneration_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    dy = torch.zeros(K, ldm, dtype=torch.bfloat16)
    dy[:, :M] = torch.randn(K, M, generator=g).to(torch.bfloat16)
    x = torch.randn(K, N, generator=g).to(torch.bfloat16)
    dyd, xd = dy.to(dev), x.to(dev)
    assert ops.wgrad_bias_supported(dyd, xd, M, N)
    dW = torch.full((M, N), float("nan"), device=dev)
    db = torch.full((M,), float("nan"), device=dev)
    ops.wgrad_bias(dyd, xd, dW, db, M, N, K, ldm, N, N)
    ref_W = dy[:, :M].double().t() @ x.double()
    ref_b = dy[:, :M].double().sum(0)
    assert (dW.double().cpu() - ref_W).abs().max() <= 1e-5 * ref_W.abs().max() + 1e-4
    assert (db.double().cpu() - ref_b).abs().max() <= 1e-5 * ref_b.abs().max() + 1e-4
    # and against the two-launch path it replaces
    dW2 = torch.empty((M, N), device=dev)
    db2 = torch.empty((M,), device=dev)
    ops.gemm(dyd, xd, dW2, M, N, K, ldm, N, N, True, False)
    ops.colsum(dyd, db2, K, M, ldm)
    assert torch.equal(dW, dW2)
    assert (db - db2).abs().max() <= 1e-5 * db2.abs().max() + 1e-4
