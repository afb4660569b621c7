"""Kernel-level parity: every C-ABI op against the oracle's restatement (or plain fp32 math) on the
same seeded inputs.  GPU only; calls go through the C ABI (ctypes)."""
import numpy as np
import pytest
import torch

from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from mmvae_b200 import ops as _ops
    _ops.lib()
    return _ops


def dev(x):
    return torch.as_tensor(x).cuda()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def clustered_csr(B, G, seed, background=0.02):
    """CSR batch whose rows hold 9..64 entries inside some 64-gene windows (real expression panels cluster:
    the kernels keep 8 entries per (cell, window) in registers and stream the rest) on top of a uniform
    background; every 7th row is empty, one row is completely dense"""
    rng = np.random.default_rng(seed)
    NW = (G + 63) // 64
    crow, cols, vals = [0], [], []
    for b in range(B):
        if b % 7 == 3:
            c = np.zeros(0, dtype=np.int64)
        elif b == 5:
            c = np.arange(G)
        else:
            c = [rng.choice(G, size=max(1, int(background * G)), replace=False)]
            for w in rng.choice(NW, size=3, replace=False):
                lo, hi = w * 64, min(G, w * 64 + 64)
                c.append(lo + rng.choice(hi - lo, size=min(hi - lo, int(rng.integers(9, 65))), replace=False))
            c = np.unique(np.concatenate(c))
        cols.append(c.astype(np.int32))
        vals.append(rng.uniform(0.5, 7.0, size=c.size).astype(np.float32))
        crow.append(crow[-1] + c.size)
    return np.asarray(crow, dtype=np.int32), np.concatenate(cols), np.concatenate(vals)


def make_csr(B, G, density, seed):
    return clustered_csr(B, G, seed) if density == "clustered" else O.synth_csr(B, G, density, seed=seed)


@pytest.mark.parametrize("B,G,H,density", [(16, 200, 64, 0.1), (33, 1000, 1024, 0.05), (8, 300, 50, 0.2)])
@pytest.mark.parametrize("wdtype", [torch.float32, torch.bfloat16])
def test_csr_linear_fwd(ops, B, G, H, density, wdtype):
    crow, col, val = O.synth_csr(B, G, density, seed=1)
    g = torch.Generator().manual_seed(0)
    W = torch.randn(H, G, generator=g) * 0.05
    b = torch.randn(H, generator=g) * 0.1
    Wt = W.t().contiguous().to(wdtype)
    ref = O.csr_linear(crow, col, val, Wt.float().t(), b)
    y = ops.csr_linear_fwd(dev(crow), dev(col), dev(val), G, Wt.cuda(), b.cuda())
    assert rel(y, ref) < 2e-6


def test_csr_linear_empty_rows(ops):
    crow = np.array([0, 0, 3, 3, 5], dtype=np.int32)
    col = np.array([1, 4, 7, 0, 9], dtype=np.int32)
    val = np.array([1, 2, 3, 4, 5], dtype=np.float32)
    W = torch.randn(8, 10)
    b = torch.randn(8)
    ref = O.csr_linear(crow, col, val, W, b)
    y = ops.csr_linear_fwd(dev(crow), dev(col), dev(val), 10, W.t().contiguous().cuda(), b.cuda())
    assert torch.allclose(y.cpu(), ref, atol=1e-6)


@pytest.mark.parametrize("B,G,H", [(24, 264, 64), (64, 2000, 1024), (5, 77, 10)])
def test_csr_transpose_and_bwd_w(ops, B, G, H):
    crow, col, val = O.synth_csr(B, G, 0.1, seed=3)
    nnz = int(crow[-1])
    cptr, ridx, cval = ops.csr_transpose(dev(crow), dev(col), dev(val), G, nnz)
    cptr_c, ridx_c, cval_c = cptr.cpu().numpy(), ridx.cpu().numpy(), cval.cpu().numpy()
    dense = O.csr_to_dense(crow, col, val, G).numpy()
    # bit-exact: CSC rebuilt to dense equals CSR rebuilt to dense; column pointers are exact counts
    assert cptr_c[0] == 0 and cptr_c[-1] == nnz
    assert np.array_equal(np.diff(cptr_c), (dense != 0).sum(0))
    rebuilt = np.zeros_like(dense)
    cols = np.repeat(np.arange(G), np.diff(cptr_c))
    rebuilt[ridx_c[:nnz], cols] = cval_c[:nnz]
    assert np.array_equal(rebuilt, dense)
    dY = torch.randn(B, H, generator=torch.Generator().manual_seed(1))
    out = torch.full((G, H), 7.0).cuda()
    ops.csr_linear_bwd_w(cptr, ridx, cval, B, G, dY.cuda(), out)
    ref = torch.from_numpy(dense).t() @ dY
    assert rel(out, ref) < 2e-6
    absent = (dense != 0).sum(0) == 0
    assert absent.any()
    assert torch.all(out.cpu()[torch.from_numpy(absent)] == 0)


@pytest.mark.parametrize("B,H", [(24, 64), (300, 1000), (1024, 512)])
@pytest.mark.parametrize("relu,p", [(1, 0.0), (1, 0.25), (0, 0.0)])
def test_bn_act_drop(ops, B, H, relu, p):
    g = torch.Generator().manual_seed(5)
    Y = (torch.randn(B, H, generator=g) * 2 + 3).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(H, generator=g)).requires_grad_(True)
    rm, rv = torch.randn(H, generator=g) * 0.1, torch.rand(H, generator=g) + 0.5
    mask = (torch.rand(B, H, generator=g) >= p)
    out, nrm, nrv = O.batch_norm_train(Y, gamma, beta, rm, rv)
    if relu:
        out = torch.relu(out)
    if p > 0:
        out = out * mask / (1 - p)
    dOut = torch.randn(B, H, generator=g)
    out.backward(dOut)

    Yc = Y.detach().cuda()
    mean, rstd = torch.empty(H).cuda(), torch.empty(H).cuda()
    rmc, rvc = rm.cuda().clone(), rv.cuda().clone()
    ops.bn_stats(Yc, 1e-3, 0.01, mean, rstd, rmc, rvc)
    assert rel(rmc, nrm) < 1e-6 and rel(rvc, nrv) < 1e-6
    o32 = torch.empty(B, H).cuda()
    o16 = torch.empty(B, H, dtype=torch.bfloat16).cuda()
    m8 = mask.to(torch.uint8).cuda() if p > 0 else None
    ops.bn_act_drop_fwd(Yc, mean, rstd, gamma.detach().cuda(), beta.detach().cuda(), relu, p, 0, m8, o32, o16)
    assert rel(o32, out.detach()) < 1e-5
    assert rel(o16.float(), out.detach()) < 1e-2
    dY, dg, db, dbias = torch.empty(B, H).cuda(), torch.empty(H).cuda(), torch.empty(H).cuda(), torch.empty(H).cuda()
    ops.bn_act_drop_bwd(dOut.cuda(), Yc, o32, mean, rstd, gamma.detach().cuda(), relu, p, 0, m8, dY, None, dg, db,
                        dbias)
    assert rel(dY, Y.grad) < 2e-4
    assert rel(dg, gamma.grad) < 1e-4 and rel(db, beta.grad) < 1e-4
    assert float(dbias.abs().max()) < 1e-2 * float(dY.abs().max()) * B ** 0.5


def test_dropout_hash_statistics(ops):
    B, H, p = 512, 1024, 0.1
    Y = torch.ones(B, H).cuda()
    o = torch.empty(B, H).cuda()
    ops.bn_act_drop_fwd(Y, None, None, None, None, 0, p, 1234, None, o, None)
    keep = (o > 0).float().mean().item()
    assert abs(keep - 0.9) < 5e-3
    assert torch.allclose(o[o > 0], torch.tensor(1 / 0.9).cuda())
    # backward regenerates the same mask from the seed
    d = torch.empty(B, H).cuda()
    ops.bn_act_drop_bwd(torch.ones(B, H).cuda(), None, None, None, None, None, 0, p, 1234, None, d, None, None, None,
                        None)
    assert torch.equal(d > 0, o > 0)
    o2 = torch.empty(B, H).cuda()
    ops.bn_act_drop_fwd(Y, None, None, None, None, 0, p, 99, None, o2, None)
    assert not torch.equal(o2 > 0, o > 0)


SHAPES = [(128, 128, 64), (24, 64, 264), (300, 200, 96), (1024, 512, 1024), (257, 1000, 520), (64, 16, 32)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_f32(ops, M, N, K, tA, tB):
    g = torch.Generator().manual_seed(M + N + K)
    A, Bm = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    ref = torch.relu(A.double() @ Bm.double().t() + bias.double()).float()
    Ad = (A.t().contiguous() if tA else A).cuda()
    Bd = (Bm.t().contiguous() if tB else Bm).cuda()
    C = torch.empty(M, N).cuda()
    ops.gemm(Ad, tA, Bd, tB, M, N, K, bias=bias.cuda(), relu=True, C32=C)
    assert rel(C, ref) < 1e-5


def _pad8(t):
    """row-major 2-D bf16 with leading dimension rounded up to 8 elements (TMA pitch rule)"""
    R, C = t.shape
    ld = (C + 7) // 8 * 8
    buf = torch.zeros(R, ld, dtype=t.dtype, device=t.device)
    buf[:, :C] = t
    return buf[:, :C]


@pytest.mark.parametrize("M,N,K", SHAPES + [(1024, 2048, 128), (2048, 8192, 256), (130, 60530, 64),
                                             (256, 1024, 8192), (1024, 1024, 60530)])  # last two: split-K
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_bf16_tc(ops, M, N, K, tA, tB):
    g = torch.Generator().manual_seed(M * 3 + N + K)
    A = torch.randn(M, K, generator=g).bfloat16()
    Bm = torch.randn(N, K, generator=g).bfloat16()
    bias = torch.randn(N, generator=g)
    ref = (A.cuda().float() @ Bm.cuda().float().t() + bias.cuda())
    Ad = _pad8((A.t().contiguous() if tA else A).cuda())
    Bd = _pad8((Bm.t().contiguous() if tB else Bm).cuda())
    C = _pad8(torch.empty(M, N).cuda())
    C16 = None
    ops.gemm(Ad, tA, Bd, tB, M, N, K, bias=bias.cuda(), relu=False, C32=C)
    assert rel(C, ref) < (2e-5 if K < 8192 else 2e-4), (tA, tB)   # long K: fp32 summation-order noise
    if K >= 8192:
        return  # the long-K shapes exist for the split-K path
    # relu + accumulate + bf16 output
    C2 = _pad8(torch.ones(M, N).cuda())
    ld = C2.stride(0)
    C16 = torch.zeros(M, ld, dtype=torch.bfloat16).cuda()[:, :N]
    ops.gemm(Ad, tA, Bd, tB, M, N, K, bias=bias.cuda(), relu=True, accumulate=True, C32=C2, C16=C16)
    ref2 = torch.relu(ref + 1.0)
    assert rel(C2, ref2) < (2e-5 if K < 8192 else 2e-4)
    assert rel(C16.float(), ref2) < 1e-2


def _pad4(t):
    R, C = t.shape
    ld = (C + 3) // 4 * 4
    buf = torch.zeros(R, ld, dtype=t.dtype, device=t.device)
    buf[:, :C] = t
    return buf[:, :C]


def _tf32(t):
    """operand as the tensor core reads it: fp32 ROUNDED to TF32's 10-bit mantissa (the TMA unit rounds on the way
    into shared memory, tensor-map data type TFLOAT32; ties may differ from this half-away emulation)"""
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K", [(24, 64, 32), (130, 72, 100), (1024, 512, 1024), (1024, 256, 512), (300, 1024, 36),
                                   (512, 280, 64), (256, 128, 4096)])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_tf32_tc(ops, M, N, K, tA, tB):
    """tcgen05 kind::tf32 on fp32 operands (K-major and MN-major, edge tiles): exact against the product of the
    TF32-truncated operands; within TF32 rounding (2^-11) of the fp32 product -- 8x finer than bf16 operands"""
    g = torch.Generator().manual_seed(M * 5 + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    Bm = torch.randn(N, K, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref_t = _tf32(A).double() @ _tf32(Bm).double().t() + bias.double()
    ref = A.double() @ Bm.double().t() + bias.double()
    Ad = _pad4(A.t().contiguous() if tA else A)
    Bd = _pad4(Bm.t().contiguous() if tB else Bm)
    C = _pad8(torch.empty(M, N).cuda())
    ops.gemm(Ad, tA, Bd, tB, M, N, K, bias=bias, C32=C, tf32=True)
    assert rel(C, ref_t) < 5e-5, (tA, tB, rel(C, ref_t), rel(C, ref))
    assert rel(C, ref) < 5e-4        # rounding, not truncation: no systematic shrink of the product
    assert abs(float((C.double() - ref).sum() / ref.abs().sum())) < 2e-5
    C2 = _pad8(torch.ones(M, N).cuda())
    C16 = torch.zeros(M, C2.stride(0), dtype=torch.bfloat16).cuda()[:, :N]
    ops.gemm(Ad, tA, Bd, tB, M, N, K, bias=bias, relu=True, accumulate=True, C32=C2, C16=C16, tf32=True)
    assert rel(C2, torch.relu(ref + 1.0)) < 1e-3
    assert rel(C16.float(), torch.relu(ref + 1.0)) < 1e-2


def test_reparam_kl(ops):
    B, Z = 48, 32
    g = torch.Generator().manual_seed(2)
    ML = (torch.randn(B, 2 * Z, generator=g) * 0.7).requires_grad_(True)
    eps = torch.randn(B, Z, generator=g)
    mu, lv = ML[:, :Z], ML[:, Z:]
    sigma = torch.sqrt(torch.exp(lv) + 1e-4)
    z = mu + eps * sigma
    kl = O.kl_normal_std(mu, sigma)
    dz = torch.randn(B, Z, generator=g)
    klw = 0.37
    (klw * kl + (z * dz).sum()).backward()
    z32 = torch.empty(B, Z).cuda()
    z16 = torch.empty(B, Z, dtype=torch.bfloat16).cuda()
    sums = torch.empty(3, dtype=torch.float64).cuda()
    ops.reparam_kl_fwd(ML.detach().cuda(), eps.cuda(), Z, 1e-4, z32, z16, sums)
    assert rel(z32, z.detach()) < 1e-6
    s = sums.cpu()
    assert abs(s[0].item() / B - kl.item()) < 1e-5 * abs(kl.item())
    assert abs(s[1].item() / (B * Z) - mu.mean().item()) < 1e-6
    assert abs(s[2].item() / (B * Z) - (sigma ** 2).mean().item()) < 1e-5
    dML = torch.empty(B, 2 * Z).cuda()
    ops.reparam_kl_bwd(ML.detach().cuda(), eps.cuda(), dz.cuda(), Z, 1e-4, klw / B, dML, None)
    assert rel(dML, ML.grad) < 1e-5


@pytest.mark.parametrize("C", [5, 11, 272, 1000])
def test_softmax_ce(ops, C):
    B = 37
    g = torch.Generator().manual_seed(C)
    logits = (torch.randn(B, C, generator=g) * 3).requires_grad_(True)
    labels = torch.randint(0, C, (B,), generator=g)
    loss = O.cross_entropy_sum(logits, labels)
    (loss * 2.5).backward()
    dl = torch.empty(B, C).cuda()
    ls = torch.zeros(1, dtype=torch.float64).cuda()
    ops.softmax_ce_sum(logits.detach().cuda(), C, labels.cuda(), 2.5, dl, ls)
    assert abs(ls.item() - loss.item()) < 1e-5 * abs(loss.item())
    assert rel(dl, logits.grad) < 1e-5


@pytest.mark.parametrize("B,G", [(24, 264), (64, 6053)])
def test_mse_relu_csr(ops, B, G):
    crow, col, val = O.synth_csr(B, G, 0.08, seed=9)
    x = O.csr_to_dense(crow, col, val, G)
    logits = (torch.randn(B, G, generator=torch.Generator().manual_seed(4)) * 2).requires_grad_(True)
    loss = ((torch.relu(logits) - x) ** 2).sum()
    loss.backward()
    ld = (G + 63) // 64 * 64
    lg = logits.detach().cuda().clone()
    dl32 = torch.zeros(B, G).cuda()
    dl16 = torch.zeros(B, ld, dtype=torch.bfloat16).cuda()
    ls = torch.empty(1, dtype=torch.float64).cuda()
    ops.mse_relu_csr(lg, G, dev(crow), dev(col), dev(val), True, dl32, None, ls)
    assert abs(ls.item() - loss.item()) < 1e-5 * loss.item()
    assert rel(dl32, logits.grad) < 1e-6
    assert torch.equal(lg.cpu(), torch.relu(logits.detach()))
    lg2 = logits.detach().cuda().clone()
    ops.mse_relu_csr(lg2, G, dev(crow), dev(col), dev(val), False, None, dl16, ls)
    assert rel(dl16[:, :G].float(), logits.grad) < 1e-2
    assert torch.all(dl16[:, G:] == 0)


def test_sumsq_and_clip_adam(ops):
    n = 100003
    g_ = torch.Generator().manual_seed(8)
    p, g = torch.randn(n, generator=g_), torch.randn(n, generator=g_) * 3
    m, v = torch.randn(n, generator=g_) * 0.1, torch.rand(n, generator=g_) * 0.1
    nb = (n + 3) // 4 * 4

    def padded(t):
        b = torch.zeros(nb)
        b[:n] = t
        return b.cuda()

    pc, gc, mc, vc = padded(p), padded(g), padded(m), padded(v)
    ns = torch.zeros(1, dtype=torch.float64).cuda()
    ops.sumsq(gc[:n], ns)
    assert abs(ns.item() ** 0.5 - g.double().norm().item()) < 1e-9 * g.double().norm().item() + 1e-9
    gl, total = O.clip_by_norm([g], 10.0)
    rp, rm_, rv_ = O.adam_update(p, gl[0], m, v, 3, 5e-3, 1e-6, (0.9, 0.999), 1e-8)
    p16 = torch.zeros(nb, dtype=torch.bfloat16).cuda()
    ops.clip_adam(pc[:n], gc[:n], mc[:n], vc[:n], p16[:n], ns, 10.0, 1.0, 5e-3, 0.9, 0.999, 1e-8, 1e-6, 3)
    assert rel(pc[:n], rp) < 1e-6 and rel(mc[:n], rm_) < 1e-6 and rel(vc[:n], rv_) < 1e-6
    assert torch.equal(p16[:n].cpu(), pc[:n].cpu().bfloat16())


def test_transpose_cast_axpy(ops):
    a = torch.randn(70, 130).cuda()
    t = torch.empty(130, 70).cuda()
    ops.transpose(a, t)
    assert torch.equal(t, a.t().contiguous())
    b16 = torch.empty(70, 130, dtype=torch.bfloat16).cuda()
    ops.cast_bf16(a, b16)
    assert torch.equal(b16, a.bfloat16())
    c = torch.randn(70, 130).cuda()
    ref = a + 0.5 * c
    ops.axpy(a, c, 0.5)
    assert torch.allclose(a, ref)


def test_cpu_tensor_rejected(ops):
    with pytest.raises(RuntimeError):
        ops.cast_bf16(torch.randn(4), torch.empty(4, dtype=torch.bfloat16))


@pytest.mark.parametrize("B,G,H,density", [(24, 264, 64, 0.1), (130, 1000, 128, 0.05), (256, 6053, 1024, 0.05),
                                            (100, 3000, 256, 0.3), (200, 2000, 256, "clustered"),
                                            (1024, 60530, 1024, 0.05)],
                         ids=["tiny", "1000", "6053", "dense30", "clustered-windows", "BASELINE-config2-shape"])
def test_decoder_mse_fused(ops, B, G, H, density):
    """fused tcgen05 GEMM + ReLU + sum-MSE-vs-CSR epilogue against the dense restatement -- up to the shape
    bench.py times (1024 cells x 60 530 genes x 1024: 946-window pointer table, 1896 tiles over 148 CTAs) and
    with rows holding more than 8 entries per 64-gene window (the spill loop)"""
    crow, col, val = make_csr(B, G, density, 11)
    x = O.csr_to_dense(crow, col, val, G).cuda()
    g = torch.Generator().manual_seed(6)
    h = torch.relu(torch.randn(B, H, generator=g)).bfloat16().cuda()
    W = (torch.randn(G, H, generator=g) * (2.0 / H) ** 0.5).bfloat16().cuda()
    bout = (torch.randn(G, generator=g) * 0.1).cuda()
    logits = (h.float() @ W.float().t() + bout).requires_grad_(True)
    loss = ((torch.relu(logits) - x) ** 2).sum()
    loss.backward()
    ldd = (G + 63) // 64 * 64
    dl16 = torch.full((B, ldd), 3.0, dtype=torch.bfloat16).cuda()
    ls = torch.empty(1, dtype=torch.float64).cuda()
    ops.decoder_mse_fused(_pad8(h), _pad8(W), bout, G, dev(crow), dev(col), dev(val), dl16, ls)
    torch.cuda.synchronize()
    assert abs(ls.item() - loss.item()) < 2e-5 * loss.item()
    assert rel(dl16[:, :G].float(), logits.grad) < 6e-3   # bf16 output rounding
    assert torch.all(dl16[:, G:(G + 7) // 8 * 8] == 0)
    # exact gene masking: dlogits is exactly 0 wherever logits <= 0 and x == 0
    dead = (logits.detach() < -1e-3) & (x == 0)
    assert torch.all(dl16[:, :G][dead] == 0)


@pytest.mark.parametrize("B,G,H,density", [(24, 264, 64, 0.1), (300, 3000, 512, 0.05), (130, 1000, 1024, 0.2),
                                            (1024, 6053, 1024, 0.05), (64, 777, 256, 0.5),
                                            (200, 2000, 256, "clustered"), (1024, 60530, 1024, 0.05)],
                         ids=["tiny", "3000", "dense20", "6053", "dense50", "clustered-windows",
                              "BASELINE-config2-shape"])
def test_csr_linear_tc_fwd_bwd(ops, B, G, H, density):
    """tensor-pipe SpMM (tile densified in smem): forward and weight gradient against the oracle with
    x rounded to bf16 (the staged operand precision); exact zeros for absent genes; pointer table exact.
    Covers the benched shape (148-way split-K over 946 gene windows) and rows with more than 8 entries per
    64-gene window (entries beyond the register-resident 8 are streamed and must be un-scattered again)."""
    crow, col, val = make_csr(B, G, density, 21)
    g = torch.Generator().manual_seed(2)
    Wt16 = (torch.randn(G, H, generator=g) * 0.05).bfloat16()
    b = torch.randn(H, generator=g) * 0.1
    crow_d, col_d, val_d = dev(crow), dev(col), dev(val)
    nnz = int(crow[-1])
    tp, packed = ops.csr_tile_ptr(crow_d, col_d, val_d, G, nnz)
    ntp = (G + 63) // 64 + 1
    tp_c = tp.cpu().numpy()[:ntp * B].reshape(ntp, B)
    for r in (0, B // 2, B - 1):   # bit-exact pointer table (lower bounds of 64-gene windows)
        cols = col[crow[r]:crow[r + 1]]
        want = crow[r] + np.searchsorted(cols, np.arange(ntp) * 64, side="left")
        assert np.array_equal(tp_c[:, r], want)
    pk = packed.cpu().numpy().view(np.uint32)[:nnz]   # bit-exact gene ids, bf16-rounded values
    assert np.array_equal(pk & 0xFFFF, col.astype(np.uint32))
    assert np.array_equal((pk >> 16).astype(np.uint16),
                          torch.from_numpy(val).bfloat16().view(torch.int16).numpy().view(np.uint16))
    val16 = torch.from_numpy(val).bfloat16().float().numpy()
    O.FAST_CSR = nnz * H > 1 << 28    # the explicit per-non-zero restatement materialises [nnz, H] floats
    try:
        ref = O.csr_linear(crow, col, val16, Wt16.float().t(), b)
    finally:
        O.FAST_CSR = False
    y = ops.csr_linear_fwd_tc(packed, tp, B, G, Wt16.cuda(), b.cuda())
    assert rel(y, ref) < 2e-5
    dY16 = torch.randn(B, H, generator=g).bfloat16()
    dense16 = O.csr_to_dense(crow, col, val16, G)
    want_dw = dense16.t() @ dY16.float()
    dWt = torch.full((G, H), 5.0).cuda()
    ops.csr_linear_bwd_w_tc(packed, tp, B, G, dY16.cuda(), dWt)
    assert rel(dWt, want_dw) < 2e-5
    absent = (dense16 != 0).sum(0) == 0
    if absent.any():
        assert torch.all(dWt.cpu()[absent] == 0)


def test_csr_linear_bwd_w_tc_shard_equals_sum_of_full_gradients(ops):
    """data-parallel first layer (2 simulated ranks on one GPU): each rank's gene shard computed from the
    concatenated inputs == the same rows of the sum of the per-rank full gradients.  The inputs are laid out
    exactly as StepEngine._dp_gather_csr leaves them (packed records back to back at a fixed capacity, window
    pointers rebased per source rank, shard windows + closing row)."""
    from mmvae_b200 import dp
    world, B, G, H = 2, 96, 1000, 64
    per = dp.shard_rows(G, world)          # 512
    G_pad, WS = world * per, per // 64
    gen = torch.Generator(device="cuda").manual_seed(3)
    batches, dYs, fulls, tps, packs = [], [], [], [], []
    for r in range(world):
        crow, col, val = O.synth_csr(B, G, 0.06, seed=70 + r)
        crow, col, val = (torch.from_numpy(a).cuda() for a in (crow, col, val))
        nnz = int(col.numel())
        dY = torch.randn(B, H, device="cuda", generator=gen).bfloat16()
        tp, packed = ops.csr_tile_ptr(crow, col, val, G_pad, nnz)
        full = torch.zeros(G_pad, H, device="cuda")
        ops.csr_linear_bwd_w_tc(packed, tp, B, G_pad, dY, full)
        batches.append((crow, col, val)); dYs.append(dY); fulls.append(full); tps.append(tp.view(-1, B)); packs.append(packed)
    cap = max(int(p.numel()) for p in packs) + 64
    packed_all = torch.zeros(world * cap, dtype=torch.int32, device="cuda")
    for r in range(world):
        packed_all[r * cap:r * cap + packs[r].numel()] = packs[r]
    dY_all = torch.cat(dYs).contiguous()
    total = fulls[0] + fulls[1]
    for r in range(world):
        tp_shard = torch.cat([tps[s][r * WS:r * WS + WS + 1] + s * cap for s in range(world)], dim=1).contiguous()
        out = torch.full((per, H), float("nan"), device="cuda")
        ops.csr_linear_bwd_w_tc_shard(packed_all, tp_shard, world * B, G_pad, dY_all, out, r * per, (r + 1) * per)
        ref = total[r * per:(r + 1) * per]
        assert torch.isfinite(out).all()
        assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-6
        assert torch.equal(out[max(0, G - r * per):], torch.zeros_like(out[max(0, G - r * per):]))   # padding rows


def test_torch_custom_ops_reach_the_kernels(ops):
    """torch.ops.cmmvae.* (mmvae_b200/torch_ops.py) call the same C entry points as the ctypes wrappers"""
    import mmvae_b200.torch_ops  # noqa: F401
    crow, col, val = O.synth_csr(64, 500, 0.1, seed=9)
    g = torch.Generator().manual_seed(1)
    Wt = (torch.randn(500, 96, generator=g) * 0.05).cuda()
    b = torch.randn(96, generator=g).cuda()
    y = torch.ops.cmmvae.csr_linear_fwd(dev(crow), dev(col), dev(val), 500, Wt, b)
    assert torch.equal(y, ops.csr_linear_fwd(dev(crow), dev(col), dev(val), 500, Wt, b))
    dY = torch.randn(64, 96, generator=g).cuda()
    dWt = torch.ops.cmmvae.csr_linear_bwd_w(dev(crow), dev(col), dev(val), 500, dY)
    assert rel(dWt, O.csr_to_dense(crow, col, val, 500).t() @ dY.cpu()) < 1e-5
    A, Bm = torch.randn(130, 64, generator=g).cuda(), torch.randn(72, 64, generator=g).cuda()
    Cc = torch.empty(130, 72).cuda()
    torch.ops.cmmvae.gemm(A, False, Bm, False, 130, 72, 64, None, False, Cc, None, True)
    assert rel(Cc, A.double() @ Bm.double().t()) < 1e-3
    ML, eps = torch.randn(32, 16, generator=g).cuda(), torch.randn(32, 8, generator=g).cuda()
    z, sums = torch.ops.cmmvae.reparam_kl_fwd(ML, eps, 1e-4)
    assert rel(z, ML[:, :8] + eps * torch.sqrt(torch.exp(ML[:, 8:]) + 1e-4)) < 1e-6 and sums.shape == (3,)


@pytest.mark.parametrize("parallel,relu,layer_norm,Z", [(True, False, True, 64), (False, True, True, 64),
                                                         (True, True, False, 64), (True, False, True, 160),
                                                         (False, False, True, 36)])
def test_conditional_bank_matches_torch_modules(tmp_path, parallel, relu, layer_norm, Z):
    """mmvae_b200.conditional.CondBank + csrc/conditional.cu against the module route (ConditionalLayers.forward,
    components.py:581-631, under autograd) at a batch where values own several 32-row tiles: outputs, input
    gradient, the gradients of the values present (others untouched), and per-value Adam steps over 3 batches
    against torch.optim.Adam on the same modules"""
    import random
    import pandas as pd
    from helpers import rel_l2
    from mmvae_b200.conditional import CondBank
    from mmvae_b200.modules.base import FCBlockConfig
    from mmvae_b200.modules.base.components import ConditionalLayers
    import copy, os
    from mmvae_b200 import layers as L
    torch.manual_seed(5)
    B = 200                      # (Z = 160: blocks wider than one 128-column pass of the tile kernels; 36: narrower)
    L.set_precision("fp32")      # (the module route's Linear layers follow the precision policy)
    try:
        _conditional_bank_case(tmp_path, parallel, relu, layer_norm, Z, B)
    finally:
        L.set_precision("bf16")


def _conditional_bank_case(tmp_path, parallel, relu, layer_norm, Z, B):
    import copy, os, random
    import pandas as pd
    from helpers import rel_l2
    from mmvae_b200.conditional import CondBank
    from mmvae_b200.modules.base import FCBlockConfig
    from mmvae_b200.modules.base.components import ConditionalLayers
    os.makedirs(tmp_path / "shared"); os.makedirs(tmp_path / "human"); os.makedirs(tmp_path / "mouse")
    pd.DataFrame([f"a.{i}" for i in range(3)]).to_csv(tmp_path / "shared" / "unique_expression_assay.csv", header=False, index=False)
    for sp in ("human", "mouse"):
        pd.DataFrame([f"d_{i}" for i in range(40)]).to_csv(tmp_path / sp / "unique_expression_donor.csv", header=False, index=False)
    cfg = FCBlockConfig(layers=[Z], use_layer_norm=layer_norm, activation_fn=torch.nn.ReLU if relu else None)
    ref = ConditionalLayers(str(tmp_path), ["assay", "donor", "species"], cfg,
                            selection_order=["parallel"] if parallel else ["donor", "species", "assay"]).cuda()
    mine = copy.deepcopy(ref)
    bank = CondBank(mine, "cuda", lr=5e-3, weight_decay=1e-6)
    opt = torch.optim.Adam(ref.parameters(), lr=5e-3, weight_decay=1e-6)
    ws = {}
    def wsf(name, shape, dtype=torch.float32):
        return ws.setdefault((name, tuple(shape), dtype), torch.empty(shape, dtype=dtype, device="cuda"))
    rng = np.random.default_rng(0)
    for t in range(3):
        meta = pd.DataFrame({"assay": [f"a.{i}" for i in rng.integers(0, 3, B)],
                             "donor": [f"d_{i}" for i in rng.integers(0, 12 + 10 * t, B)]})
        z = torch.randn(B, Z, device="cuda", requires_grad=True)
        random.seed(10 + t)
        opt.zero_grad(set_to_none=True)
        y_ref = ref(z, meta, "human")
        wgt = torch.randn_like(y_ref)
        (y_ref * wgt).sum().backward()
        random.seed(10 + t)
        bank.make_plan(meta, "human", B)
        y, y16 = bank.forward(z.detach(), wsf, True)
        assert torch.allclose(y, y_ref, rtol=1e-3, atol=2e-4), float((y - y_ref).abs().max())
        assert torch.allclose(y16.float(), y_ref, rtol=1e-2, atol=1e-2)
        dz = bank.backward(wgt.contiguous(), wsf)
        assert torch.allclose(dz, z.grad, rtol=1e-3, atol=1e-4), float((dz - z.grad).abs().max())
        present = set(bank.plan["present"].cpu().tolist())
        for s_, ((n_, p_ref), p_mine) in enumerate(zip(ref.named_parameters(), bank.params)):
            if p_ref.grad is None:
                assert s_ // 2 not in present, n_
            else:
                assert s_ // 2 in present, n_
                assert rel_l2(p_mine.grad.cpu().numpy(), p_ref.grad.cpu().numpy()) < 2e-4 or \
                    float((p_mine.grad - p_ref.grad).abs().max()) < 1e-4, n_
        ns = torch.zeros(1, dtype=torch.float64, device="cuda")
        bank.add_norm_sq(ns)
        want = torch.nn.utils.clip_grad_norm_(ref.parameters(), 10.0)
        assert float(ns.sqrt()) == pytest.approx(float(want), rel=1e-4)
        opt.step()
        bank.clip_adam(ns, 10.0)
    for (n_, p_ref), p_mine in zip(ref.named_parameters(), bank.params):
        if n_.endswith("lin.bias") and layer_norm:
            continue      # shift-invariant under LayerNorm: its gradient is rounding noise that Adam normalises
        assert rel_l2(p_mine.detach().cpu().numpy(), p_ref.detach().cpu().numpy()) < 1e-4, n_
    steps = bank.steps.cpu().tolist()
    for s_, (n_, p_ref) in enumerate(list(ref.named_parameters())[::2]):
        assert steps[s_] == int(opt.state[p_ref]["step"]) if p_ref in opt.state else steps[s_] == 0, n_
