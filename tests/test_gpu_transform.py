"""(a3) node transforms: fp32 CUDA-core kernel and bf16 tcgen05 kernel vs float64 / fp32 torch math on
the same inputs; (a2+a3) fused input layer vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import make_graph_arrays
from oracle import port

pytestmark = pytest.mark.gpu


def _ref(a1, w1, a2, w2, b, relu):
    y = a1.double() @ w1.double().T
    if a2 is not None:
        y = y + a2.double() @ w2.double().T
    if b is not None:
        y = y + b.double()
    return y.relu() if relu else y


@pytest.mark.parametrize("M,N,K1,K2", [(1, 64, 64, 0), (129, 64, 64, 64), (1000, 128, 128, 128),
                                        (777, 1024, 1024, 1024), (300, 72, 20, 8), (4096, 256, 512, 0)])
@pytest.mark.parametrize("relu", [False, True])
def test_transform_fp32(cuda, M, N, K1, K2, relu):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a1 = torch.randn(M, K1, device=cuda, generator=g)
    w1 = torch.randn(N, K1, device=cuda, generator=g) / K1 ** 0.5
    a2 = torch.randn(M, K2, device=cuda, generator=g) if K2 else None
    w2 = torch.randn(N, K2, device=cuda, generator=g) / K2 ** 0.5 if K2 else None
    b = torch.randn(N, device=cuda, generator=g)
    y = ops.node_transform(a1, w1, a2, w2, b, relu=relu)
    e = _ref(a1, w1, a2, w2, b, relu)
    # fp32 accumulate over K <= 2048: relative error ~ sqrt(K) * 2^-24 of the row scale
    assert float((y.double() - e).abs().max()) < 2e-5 * max(1.0, float(e.abs().max()))


@pytest.mark.parametrize("M,N,K1,K2", [(128, 256, 64, 0), (128, 256, 128, 0), (128, 256, 256, 256),
                                        (1, 64, 64, 64), (129, 64, 64, 64), (1000, 128, 128, 128),
                                        (5000, 1024, 1024, 1024), (50_000, 1024, 1024, 1024), (333, 512, 192, 64)])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("bf", [torch.bfloat16, torch.float16])
def test_transform_bf16_tcgen05(cuda, M, N, K1, K2, relu, bf):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a1 = torch.randn(M, K1, device=cuda, generator=g).to(bf)
    w1 = (torch.randn(N, K1, device=cuda, generator=g) / K1 ** 0.5).to(bf)
    a2 = torch.randn(M, K2, device=cuda, generator=g).to(bf) if K2 else None
    w2 = (torch.randn(N, K2, device=cuda, generator=g) / K2 ** 0.5).to(bf) if K2 else None
    b = torch.randn(N, device=cuda, generator=g)
    y = ops.node_transform(a1, w1, a2, w2, b, relu=relu)
    torch.cuda.synchronize()
    assert y.dtype == bf and y.shape == (M, N)
    e = _ref(a1, w1, a2, w2, b, relu)           # exact products of the bf16 inputs, float64 accumulate
    err = (y.double() - e).abs()
    scale = max(1.0, float(e.abs().max()))
    # inputs are identical 16-bit values; the only errors are fp32 accumulation and the final rounding
    # (2^-9 relative for bf16, 2^-12 for IEEE half)
    k = 1.0 if bf == torch.bfloat16 else 0.125
    assert float(err.max()) < 8e-3 * k * scale, f"max err {float(err.max())} scale {scale}"
    assert float(err.mean()) < 2e-3 * k


@pytest.mark.parametrize("hids,out_dtype", [(64, torch.float32), (128, torch.float32), (1024, torch.float32),
                                             (1024, torch.bfloat16), (96, torch.float32)])
def test_conv_in_fused_vs_oracle(cuda, hids, out_dtype):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    m, n, z = 700, 1300, 6000
    row, col, val = make_graph_arrays(m, n, z, 5)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda)
    ref = port.graph_from_coo(row, col, val, m, n)
    rng = np.random.default_rng(hids)
    x_s = rng.standard_normal((m, 8)).astype(np.float32)
    x_t = rng.standard_normal((n, 8)).astype(np.float32)
    w_rel = (rng.standard_normal((hids, 8)) / 3).astype(np.float32)
    w_root = (rng.standard_normal((hids, 8)) / 3).astype(np.float32)
    b = rng.standard_normal(hids).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    csr, csc = g.views()
    # variables side: dst = vars, src = cons, orientation = CSC
    out, zc = ops.conv_in_fused(csc, t(x_s), t(x_t), t(w_rel), t(b), t(w_root), out_dtype, relu=True)
    agg_e = port.spmm_sequential(ref.colptr, ref.row_csc, ref.val_csc, x_s)
    np.testing.assert_allclose(zc[:, :8].cpu().numpy(), agg_e, rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(zc[:, 8:16].cpu().numpy(), x_t)
    e = np.maximum(agg_e.astype(np.float64) @ w_rel.T.astype(np.float64) + b + x_t.astype(np.float64) @ w_root.T, 0)
    tol = 1e-4 if out_dtype == torch.float32 else 3e-2
    np.testing.assert_allclose(out.float().cpu().numpy(), e, rtol=tol, atol=tol)
    # constraints side, no relu
    out2, _ = ops.conv_in_fused(csr, t(x_t), t(x_s), t(w_rel), t(b), t(w_root), out_dtype, relu=False)
    agg2 = port.spmm_sequential(ref.rowptr, ref.col, ref.val, x_t)
    e2 = agg2.astype(np.float64) @ w_rel.T.astype(np.float64) + b + x_s.astype(np.float64) @ w_root.T
    np.testing.assert_allclose(out2.float().cpu().numpy(), e2, rtol=tol, atol=tol)


@pytest.mark.parametrize("hids", [32, 96, 1024])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,z", [(700, 1300, 6000), (1, 3, 2), (129, 255, 0), (5000, 9000, 60_000)])
def test_conv_in_16_one_kernel_input_layer(cuda, hids, dt, m, n, z):
    """lpgnn_conv_in_16 (aggregate + 16-wide MMA + bias + ReLU + 16-bit store in one kernel) vs float64 on the operands
    as the kernel rounds them (aggregate in fp32 CSR order -> 16-bit, weights -> 16-bit): only fp32 accumulation and the
    output rounding remain.  Also the optional weight-gradient operand z16 and the two-kernel path it replaces."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    row, col, val = make_graph_arrays(m, n, z, 5)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda)
    ref = port.graph_from_coo(row, col, val, m, n)
    rng = np.random.default_rng(hids + m)
    x_s = rng.standard_normal((m, 8)).astype(np.float32)
    x_t = rng.standard_normal((n, 8)).astype(np.float32)
    w_rel = (rng.standard_normal((hids, 8)) / 3).astype(np.float32)
    w_root = (rng.standard_normal((hids, 8)) / 3).astype(np.float32)
    b = rng.standard_normal(hids).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    rd = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dt).double().numpy()   # 16-bit rounding
    csr, csc = g.views()
    for view, xs, xd, ptr_, idx_, val_, relu in ((csc, x_s, x_t, ref.colptr, ref.row_csc, ref.val_csc, True),
                                                 (csr, x_t, x_s, ref.rowptr, ref.col, ref.val, False)):
        out, z16 = ops.conv_in_16(view, t(xs), t(xd), t(w_rel), t(b), t(w_root), dt, relu=relu, want_z16=True)
        assert out.dtype == dt and out.shape == (xd.shape[0], hids)
        agg = port.spmm_sequential(ptr_, idx_, val_, xs)
        e = rd(agg) @ rd(w_rel).T + b.astype(np.float64) + rd(xd) @ rd(w_root).T
        e_scale = np.abs(rd(agg)) @ np.abs(rd(w_rel)).T + np.abs(b).astype(np.float64) + np.abs(rd(xd)) @ np.abs(rd(w_root)).T
        if relu:
            e = np.maximum(e, 0)
        eps = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
        err = np.abs(out.double().cpu().numpy() - e)
        # output rounding (eps of the value) + a one-ulp flip of an aggregate entry (eps of the terms' magnitude)
        assert float((err / (e_scale + 1e-6)).max()) < 1.5 * eps, float((err / (e_scale + 1e-6)).max())
        zz = z16.float().cpu().numpy()
        # (the kernel accumulates with fused multiply-adds, the oracle with separate roundings: at most one 16-bit ulp apart)
        np.testing.assert_allclose(zz[:, :8], rd(agg).astype(np.float32), rtol=2 * eps, atol=1e-6)
        np.testing.assert_array_equal(zz[:, 8:16], rd(xd).astype(np.float32))
        assert (zz[:, 16] == 1).all() and (zz[:, 17:] == 0).all()
        if hids % 64 == 0:   # the gather + one-K-block tcgen05 path computes the same function
            _, zb = ops.gather_cat(view, t(xs), t(xd), want_f32=False, want_bf16=True, dtype16=dt)
            assert torch.equal(zb, z16)


@pytest.mark.parametrize("hids", [32, 256, 512, 768, 1024])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,z", [(700, 1300, 6000), (1, 3, 2), (129, 255, 0), (3, 40_000, 90_000), (60_000, 110_000, 500_000),
                                   (37_999, 5, 40_000)])
def test_conv_in_16_pair_equals_two_single_launches(cuda, hids, dt, m, n, z):
    """lpgnn_conv_in_16_pair (both directions in one launch, blocks split between the sides) == two lpgnn_conv_in_16
    launches, bit for bit, outputs and weight-gradient operands; sizes below and above the co-resident grid, lopsided sides."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    row, col, val = make_graph_arrays(m, n, z, 11)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda)
    rng = np.random.default_rng(hids + n)
    t = lambda *shape: torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(cuda)
    x_s, x_t = t(m, 8), t(n, 8)
    l2r, r2l = (t(hids, 8), t(hids), t(hids, 8)), (t(hids, 8), t(hids), t(hids, 8))
    csr, csc = g.views()
    for want_z in (False, True):
        right, z_t = ops.conv_in_16(csc, x_s, x_t, *l2r, dt, relu=True, want_z16=want_z)
        left, z_s = ops.conv_in_16(csr, x_t, x_s, *r2l, dt, relu=True, want_z16=want_z)
        pl, pr, pz_s, pz_t = ops.conv_in_16_pair(csr, csc, x_s, x_t, l2r, r2l, dt, relu=True, want_z16=want_z)
        assert torch.equal(pl, left) and torch.equal(pr, right)
        if want_z:
            assert torch.equal(pz_s, z_s) and torch.equal(pz_t, z_t)
        # large inputs with N in {256, ..., 1024} take the warp-specialised register-B kernel: same bits as the other one
        from lpgnn_b200 import _lib
        prev = _lib.load().lpgnn_set_conv_in_regb(0)
        try:
            ql, qr, qz_s, qz_t = ops.conv_in_16_pair(csr, csc, x_s, x_t, l2r, r2l, dt, relu=True, want_z16=want_z)
        finally:
            _lib.load().lpgnn_set_conv_in_regb(prev)
        assert torch.equal(ql, pl) and torch.equal(qr, pr)
        if want_z:
            assert torch.equal(qz_s, pz_s) and torch.equal(qz_t, pz_t)


@pytest.mark.parametrize("M,N,K", [(1000, 1024, 1024), (129, 128, 128), (5000, 64, 64), (300, 512, 256)])
@pytest.mark.parametrize("want_out", [False, True])
def test_transform_with_fused_head(cuda, M, N, K, want_out):
    """Last-layer transform with the basis-status head + knowledge masking fused into the epilogue vs the
    unfused kernels (same bf16 inputs; the fused path keeps the ReLU'd activation in fp32)."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    bf = torch.bfloat16
    a1 = torch.randn(M, K, device=cuda, generator=g).to(bf)
    a2 = torch.randn(M, K, device=cuda, generator=g).to(bf)
    w1 = (torch.randn(N, K, device=cuda, generator=g) / K ** 0.5).to(bf)
    w2 = (torch.randn(N, K, device=cuda, generator=g) / K ** 0.5).to(bf)
    b = torch.randn(N, device=cuda, generator=g)
    hw = torch.randn(3, N, device=cuda, generator=g) / N ** 0.5
    hb = torch.randn(3, device=cuda, generator=g)
    feas = torch.randint(-1, 2, (M, 8), device=cuda, generator=g).float()
    logits, out = ops.node_transform_head(a1, w1, a2, w2, b, hw, hb, feas, relu=True, want_out=want_out)
    act = (a1.double() @ w1.double().T + a2.double() @ w2.double().T + b.double()).relu()
    raw = act @ hw.double().T + hb.double()
    exp = torch.nn.functional.normalize(raw) * 10
    exp[:, 0] -= 10 * (feas[:, 5] != 0)
    exp[:, 2] -= 10 * (feas[:, 7] != 0)
    err = (logits.double() - exp).abs()
    assert float(err.max()) < 5e-3, float(err.max())              # fp32 accumulate of exact bf16 products
    assert (out is None) == (not want_out)
    if want_out:
        ref = ops.node_transform(a1, w1, a2, w2, b, relu=True)
        assert torch.equal(out, ref)


@pytest.mark.parametrize("M,N,K", [(3000, 1024, 1024), (129, 128, 128), (5000, 64, 64), (40_000, 256, 256)])
@pytest.mark.parametrize("dp", [0.0, 0.3])
def test_training_transform_with_fused_head_sees_the_stored_activation(cuda, M, N, K, dp):
    """Training forward of the last hidden layer: ReLU + dropout + bf16 store with the head accumulated in the same
    epilogue.  The stored activation is bit-identical to the transform without the head (same keep decisions), the
    un-normalised logits are the head of exactly that activation (up to its bf16 rounding: the epilogue feeds the fp32
    values), the final logits are add_knowledge of them, and the drop rate is the requested one."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    bf = torch.bfloat16
    a1 = torch.randn(M, K, device=cuda, generator=g).to(bf)
    a2 = torch.randn(M, K, device=cuda, generator=g).to(bf)
    w1 = (torch.randn(N, K, device=cuda, generator=g) / K ** 0.5).to(bf)
    w2 = (torch.randn(N, K, device=cuda, generator=g) / K ** 0.5).to(bf)
    b = torch.randn(N, device=cuda, generator=g)
    hw = torch.randn(3, N, device=cuda, generator=g) / N ** 0.5
    hb = torch.randn(3, device=cuda, generator=g)
    feas = torch.randint(-1, 2, (M, 8), device=cuda, generator=g).float()
    drop = (dp, 1234) if dp > 0 else None
    out, logits, raw = ops.node_transform_head_train(a1, w1, a2, w2, b, hw, hb, feas, relu=True, dropout=drop)
    ref = ops.node_transform(a1, w1, a2, w2, b, relu=True, dropout=drop)
    assert torch.equal(out, ref)
    if dp > 0:
        pre = ops.node_transform(a1, w1, a2, w2, b, relu=True)
        kept = (out != 0).sum().item() / max((pre != 0).sum().item(), 1)
        assert abs(kept - (1 - dp)) < 0.01, kept
    exp_raw = out.double() @ hw.double().T + hb.double()
    scale = (out.double().abs() @ hw.double().abs().T + hb.double().abs())
    assert float(((raw.double() - exp_raw).abs() / scale).max()) < 2.0 ** -8        # bf16 rounding of the stored values
    exp = torch.nn.functional.normalize(raw.double()) * 10
    exp[:, 0] -= 10 * (feas[:, 5] != 0)
    exp[:, 2] -= 10 * (feas[:, 7] != 0)
    assert float((logits.double() - exp).abs().max()) < 1e-4


@pytest.mark.parametrize("M,N,K1,K2", [(1000, 1024, 1024, 1024), (129, 64, 64, 64), (5000, 128, 128, 0), (333, 512, 192, 64)])
@pytest.mark.parametrize("parts,tol", [(2, 4e-5), (3, 2e-5)])
def test_transform_fp32_via_split_bf16_tensor_core_passes(cuda, M, N, K1, K2, parts, tol):
    """fp32 operands split into 2 / 3 bf16 parts, 3 / 6 tensor-core passes.  The split itself is exact to 2^-17 / 2^-25;
    the result error is then bounded by the tensor core's own fp32 accumulation (truncating adds: ~1e-6 at K=128,
    ~1.5e-5 at K=2048 measured), which is why parts=3 is only ~2x better than parts=2 at the largest K."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K1)
    a1 = torch.randn(M, K1, device=cuda, generator=g)
    w1 = torch.randn(N, K1, device=cuda, generator=g) / K1 ** 0.5
    a2 = torch.randn(M, K2, device=cuda, generator=g) if K2 else None
    w2 = torch.randn(N, K2, device=cuda, generator=g) / K2 ** 0.5 if K2 else None
    b = torch.randn(N, device=cuda, generator=g)
    ps = ops.split_bf16(a1, parts)
    resid = float((sum(p.float() for p in ps) - a1).abs().max()) / float(a1.abs().max())
    assert resid < (2 ** -16 if parts == 2 else 2 ** -23)
    sp = lambda x: ops.split_bf16(x, parts)
    y = ops.node_transform_split(sp(a1), sp(w1), sp(a2) if K2 else None, sp(w2) if K2 else None, b, relu=True)
    e = _ref(a1, w1, a2, w2, b, True)
    err = float((y.double() - e).abs().max()) / max(1.0, float(e.abs().max()))
    assert err < tol, err
    y_simt = ops.node_transform(a1, w1, a2, w2, b, relu=True)
    assert float((y - y_simt).abs().max()) < 2e-4 * max(1.0, float(e.abs().max()))


def test_split_x2_is_a_22_bit_power_of_two_scaled_representation(cuda):
    """x = scale * (hi + 2^-11 lo): scale is a power of two per row, shared by the two tensors of a call; the
    reconstruction error is 2^-22 of the element (huge and tiny rows alike, no overflow), zero rows stay zero."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    x1 = torch.randn(300, 192, device=cuda, generator=g)
    x2 = torch.randn(300, 64, device=cuda, generator=g) * 3
    x1[0] *= 1e30; x2[0] *= 1e28          # far above the half range
    x1[1] *= 1e-30; x2[1] *= 1e-33        # far below it
    x1[2] = 0; x2[2] = 0
    x1[3, 5] = 7e4                        # one element above 65504 in an otherwise ordinary row
    (h1, l1), (h2, l2), sc = ops.split_x2(x1, x2)
    assert h1.dtype == torch.float16 and sc.shape == (300,)
    assert torch.isfinite(h1.float()).all() and torch.isfinite(l1.float()).all() and torch.isfinite(h2.float()).all()
    mant = torch.frexp(sc)[0]
    assert torch.equal(mant, torch.full_like(mant, 0.5))          # exact powers of two
    assert float(sc[2]) == 1.0
    for x, h, l in ((x1, h1, l1), (x2, h2, l2)):
        rec = sc.double()[:, None] * (h.double() + l.double() / 2048)
        rowmax = torch.maximum(x1.abs().amax(1), x2.abs().amax(1)).double()[:, None].clamp_min(1e-300)
        err = (rec - x.double()).abs()
        assert float((err / x.abs().double().clamp_min(1e-300))[x != 0].max()) < 2 ** -21 or float((err / rowmax).max()) < 2 ** -33
        assert float((err / rowmax).max()) < 2 ** -22
    # scaled row maximum sits in [2^12, 2^13)
    top = torch.maximum(h1.float().abs().amax(1), h2.float().abs().amax(1))
    nz = torch.ones(300, dtype=torch.bool, device=cuda); nz[2] = False
    assert float(top[nz].min()) >= 4096 and float(top[nz].max()) <= 8192


@pytest.mark.parametrize("M,N,K1,K2", [(1, 64, 64, 0), (129, 64, 64, 64), (1000, 128, 128, 128), (333, 512, 192, 64),
                                        (2047, 256, 256, 256), (2048, 1024, 1024, 1024), (2049, 1024, 1024, 1024),
                                        (5000, 1024, 1024, 1024), (50_000, 1024, 1024, 1024)])
@pytest.mark.parametrize("relu", [False, True])
def test_transform_fp32_on_tensor_cores_x2(cuda, M, N, K1, K2, relu):
    """lpgnn_node_transform_x2 (three half x half passes, chunked accumulation) vs float64 on the same fp32 inputs:
    as accurate as the CUDA-core SGEMM-style kernel, rows of very different magnitudes included."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K1)
    rowmag = torch.exp(torch.randn(M, 1, device=cuda, generator=g) * 3)
    a1 = torch.randn(M, K1, device=cuda, generator=g).relu() * rowmag
    w1 = torch.randn(N, K1, device=cuda, generator=g) / K1 ** 0.5
    a2 = torch.randn(M, K2, device=cuda, generator=g) * rowmag if K2 else None
    w2 = torch.randn(N, K2, device=cuda, generator=g) / K2 ** 0.5 if K2 else None
    b = torch.randn(N, device=cuda, generator=g)
    pa1, pa2, rs = ops.split_x2(a1, a2)
    pw1, pw2, cs = ops.split_x2(w1, w2)
    y = ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, b, relu=relu)
    torch.cuda.synchronize()
    e = _ref(a1, w1, a2, w2, b, relu)
    # error relative to the row's own scale (the row normalisation of add_knowledge makes exactly this matter)
    rscale = e.abs().amax(1, keepdim=True).clamp_min(1.0)
    err = float(((y.double() - e).abs() / rscale).max())
    y_simt = ops.node_transform(a1, w1, a2, w2, b, relu=relu)
    err_simt = float(((y_simt.double() - e).abs() / rscale).max())
    print(f"x2 M={M} K={K1}+{K2}: err {err:.2e}  (CUDA-core fp32 kernel: {err_simt:.2e})")
    assert err < 2e-6, (err, err_simt)


def test_transform_x2_chunk_length_bounds_the_accumulation_drift(cuda):
    """The reason for the chunked accumulation: one TMEM chunk over the whole reduction drifts (truncating adds);
    short chunks summed in registers do not."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 4096, 1024, 1024
    a1, a2 = torch.randn(M, K, device=cuda, generator=g).relu(), torch.randn(M, K, device=cuda, generator=g).relu()
    w1, w2 = torch.randn(N, K, device=cuda, generator=g) / 32, torch.randn(N, K, device=cuda, generator=g) / 32
    pa1, pa2, rs = ops.split_x2(a1, a2)
    pw1, pw2, cs = ops.split_x2(w1, w2)
    e = _ref(a1, w1, a2, w2, None, False)
    errs = {}
    prev = ops.set_x2_chunk(4)
    try:
        for ck in (1, 2, 4, 8, 16, 64):
            ops.set_x2_chunk(ck)
            y = ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, None)
            errs[ck] = float((y.double() - e).norm() / e.norm())
    finally:
        ops.set_x2_chunk(prev)
    print("x2 relative Frobenius error by chunk length (K-blocks of 64):", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs[4] < 5e-7
    assert errs[1] <= errs[64]


@pytest.mark.parametrize("M,N", [(3000, 1024), (700, 128), (129, 64)])
def test_transform_x2_fused_head_equals_head_on_the_written_activation(cuda, M, N):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M)
    a1, a2 = torch.randn(M, N, device=cuda, generator=g), torch.randn(M, N, device=cuda, generator=g)
    w1, w2 = torch.randn(N, N, device=cuda, generator=g) / N ** 0.5, torch.randn(N, N, device=cuda, generator=g) / N ** 0.5
    b = torch.randn(N, device=cuda, generator=g)
    hw, hb = torch.randn(3, N, device=cuda, generator=g) / N ** 0.5, torch.randn(3, device=cuda, generator=g)
    feas = torch.randn(M, 8, device=cuda, generator=g)
    feas[:, 5] = (feas[:, 5] > 0.5).float(); feas[:, 7] = (feas[:, 7] < -0.5).float()
    pa1, pa2, rs = ops.split_x2(a1, a2)
    pw1, pw2, cs = ops.split_x2(w1, w2)
    out, logits = ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, b, relu=True, head=(hw, hb, feas))
    none, logits2 = ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, b, relu=True, head=(hw, hb, feas), want_out=False)
    assert none is None and torch.equal(logits, logits2)
    assert torch.equal(out, ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, b, relu=True))
    exp, _ = ops.head_mask(out, hw, hb, feas)
    assert float((logits - exp).abs().max()) < 2e-4


@pytest.mark.parametrize("M", [2048, 2049, 2304, 5000, 33_333])
def test_cta_pair_transform_is_bit_identical_to_single_cta(cuda, M):
    """The cta_group::2 form of the wide bf16 transform (two CTAs, one 256-row MMA, half of W per CTA) against the
    one-CTA-per-tile kernel: same MMAs per output element -> same bits; odd row-block counts exercise the pair whose
    second CTA lies beyond M; with ReLU, bias, the fused head, and the training epilogues (mask, dropout)."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M)
    bf = torch.bfloat16
    a1 = torch.randn(M, 1024, device=cuda, generator=g).to(bf)
    a2 = torch.randn(M, 512, device=cuda, generator=g).to(bf)
    w1 = (torch.randn(1024, 1024, device=cuda, generator=g) / 32).to(bf)
    w2 = (torch.randn(1024, 512, device=cuda, generator=g) / 32).to(bf)
    b = torch.randn(1024, device=cuda, generator=g)
    hw, hb = torch.randn(3, 1024, device=cuda, generator=g), torch.randn(3, device=cuda, generator=g)
    x = torch.randn(M, 8, device=cuda, generator=g)
    act = torch.randn(M, 1024, device=cuda, generator=g).to(bf)
    outs = {}
    try:
        for mode in (0, 1):
            lib.lpgnn_set_gemm_cluster(mode)
            head = ops.node_transform_head(a1, w1, a2, w2, b, hw, hb, x)
            outs[mode] = (ops.node_transform(a1, w1, a2, w2, b, relu=True),
                          ops.node_transform(a1, w1, None, None, None, relu=False),
                          head[0] if isinstance(head, tuple) else head,
                          ops.node_transform(a1, w1, a2, w2, b, relu=True, dropout=(0.1, 1234)),
                          ops.node_transform(a1, w1, a2, w2, None, relu=False, mask=(act, 1.0 / 0.9)))
    finally:
        lib.lpgnn_set_gemm_cluster(1)
    for u, v in zip(outs[0], outs[1]):
        assert torch.equal(u, v)
    ref = torch.relu(a1.float() @ w1.float().t() + a2.float() @ w2.float().t() + b)
    assert float((outs[1][0].float() - ref).abs().max()) < 0.05 * float(ref.abs().max())


def test_cta_pair_wgrad_is_bit_identical_to_single_cta(cuda):
    """Weight gradient dW = dY^T X (MN-major operands, split-K) in the pair form vs the one-CTA form and vs fp32 torch."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(7)
    Mn = 30_000
    dy = torch.randn(Mn, 1024, device=cuda, generator=g).to(torch.bfloat16)
    xx = torch.randn(Mn, 1024, device=cuda, generator=g).to(torch.bfloat16)
    try:
        lib.lpgnn_set_gemm_cluster(0)
        d0 = ops.wgrad(dy, xx)
        lib.lpgnn_set_gemm_cluster(1)
        d1 = ops.wgrad(dy, xx)
    finally:
        lib.lpgnn_set_gemm_cluster(1)
    assert torch.equal(d0, d1)
    ref = dy.float().t() @ xx.float()
    assert float((d1 - ref).abs().max()) < 2e-3 * float(ref.abs().max())


def test_conv_in_fused_above_the_old_grid_y_limit(cuda):
    """fp32 input layer on more than 65535 * 64 rows (row blocks used to sit on gridDim.y): LPs below edge_num_thresh =
    1.2e7 edges can have more nodes than that.  A diagonal-like graph keeps the expected values trivial."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    rows = 65535 * 64 + 70
    ptr = torch.arange(rows + 1, dtype=torch.int32, device=cuda)            # one neighbour per row: node (i mod 1000)
    idx = (torch.arange(rows, device=cuda) % 1000).to(torch.int32)
    val = torch.full((rows,), 2.0, device=cuda)
    g = torch.Generator(device="cuda").manual_seed(0)
    x_src = torch.randn(1000, 8, device=cuda, generator=g)
    x_dst = torch.randn(rows, 8, device=cuda, generator=g)
    w_rel, w_root = torch.randn(64, 8, device=cuda, generator=g), torch.randn(64, 8, device=cuda, generator=g)
    b = torch.randn(64, device=cuda, generator=g)
    out, _ = ops.conv_in_fused((ptr, idx, val, rows), x_src, x_dst, w_rel, b, w_root, torch.float32, relu=True)
    for r in (0, 1, 65535 * 64 - 1, 65535 * 64, rows - 1):
        e = torch.relu((2.0 * x_src[r % 1000]) @ w_rel.T + b + x_dst[r] @ w_root.T)
        assert torch.allclose(out[r], e, rtol=1e-5, atol=1e-5), r


@pytest.mark.parametrize("m,n,z,H", [(700, 1300, 6000, 64), (20_000, 40_000, 200_000, 1024), (129, 255, 0, 128)])
def test_x2_producers_write_the_operands_directly(cuda, m, n, z, H):
    """lpgnn_conv_in_fused_x2 / lpgnn_spmm_x2: the fp32 input layer and the aggregation emit x2 operands themselves (row
    scale from an a-priori bound instead of the row maximum).  scale * (hi + 2^-11 lo) must reproduce the fp32 results to
    22 bits relative to the row's magnitude, the scales are powers of two that bound their rows, and the banded-sweep and
    two-step forms of spmm_x2 agree with the plain fp32 aggregation."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    row, col, val = make_graph_arrays(m, n, z, 5, sort=True)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda, is_sorted=True)
    csr, csc = g.views()
    gen = torch.Generator(device="cuda").manual_seed(H + m)
    x_s, x_t = torch.randn(m, 8, device=cuda, generator=gen), torch.randn(n, 8, device=cuda, generator=gen) * 5
    w_rel, w_root = torch.randn(H, 8, device=cuda, generator=gen) / 3, torch.randn(H, 8, device=cuda, generator=gen) / 3
    b = torch.randn(H, device=cuda, generator=gen)
    out, (hi, lo), sc = ops.conv_in_fused_x2(csc, x_s, x_t, w_rel, b, w_root, relu=True)
    ref, _ = ops.conv_in_fused(csc, x_s, x_t, w_rel, b, w_root, torch.float32, relu=True)
    assert torch.equal(out, ref)
    mant = torch.frexp(sc)[0]
    assert torch.equal(mant, torch.full_like(mant, 0.5))
    assert bool((out.abs().amax(1) <= sc * 4096).all())                       # the guarantee spmm_x2 builds on
    rec = sc.double()[:, None] * (hi.double() + lo.double() / 2048)
    rowmax = out.abs().amax(1, keepdim=True).double().clamp_min(1e-30)
    assert float(((rec - out.double()).abs() / rowmax).max()) < 2.0 ** -21
    # aggregation of that output (constraint side: A . right) straight into x2 operands
    (ahi, alo), asc = ops.spmm_x2(csr, out, sc)
    agg = ops.spmm(csr, out)
    arec = asc.double()[:, None] * (ahi.double() + alo.double() / 2048)
    amax = agg.abs().amax(1, keepdim=True).double().clamp_min(1e-30)
    assert torch.isfinite(ahi.float()).all() and torch.isfinite(alo.float()).all()
    assert float(((arec - agg.double()).abs() / amax).max()) < 2.0 ** -19      # bound-based scale: a few bits of head room spent
    assert bool((agg.abs().amax(1) <= asc * 8192).all())        # (two-step form: scale from the row maximum, [2^12, 2^13))
