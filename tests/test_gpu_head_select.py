"""(a4-a6) head + knowledge masking (bit-exact mask) and the basis decision vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _feas(rows, rng):
    f = rng.standard_normal((rows, 8)).astype(np.float32)
    f[:, 5] = rng.choice([0.0, 1.0, -1.0], size=rows, p=[0.5, 0.25, 0.25])
    f[:, 7] = rng.choice([0.0, 1.0, -1.0], size=rows, p=[0.4, 0.5, 0.1])
    return f


@pytest.mark.parametrize("H,dtype", [(64, torch.float32), (128, torch.float32), (1024, torch.float32),
                                      (1024, torch.bfloat16), (64, torch.bfloat16), (1024, torch.float16),
                                      (128, torch.float16), (2048, torch.float16)])
def test_head_mask_vs_oracle(cuda, H, dtype):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    rows = 3001
    rng = np.random.default_rng(H)
    h = torch.from_numpy(rng.standard_normal((rows, H)).astype(np.float32)).to(dtype)
    w = (rng.standard_normal((3, H)) / H ** 0.5).astype(np.float32)
    b = rng.standard_normal(3).astype(np.float32)
    feas = _feas(rows, rng)
    logits, raw = ops.head_mask(h.to(cuda), torch.from_numpy(w).to(cuda), torch.from_numpy(b).to(cuda),
                                torch.from_numpy(feas).to(cuda), want_raw=True)
    raw_e = h.float().numpy().astype(np.float64) @ w.T.astype(np.float64) + b
    np.testing.assert_allclose(raw.cpu().numpy(), raw_e, rtol=1e-4, atol=1e-4)
    e, _ = port.add_knowledge_np(raw_e.astype(np.float32), raw_e.astype(np.float32)[:1], feas, feas[:1])
    got = logits.cpu().numpy()
    np.testing.assert_allclose(got, e, rtol=1e-4, atol=2e-4)
    # the mask itself is exact: masked entries sit in [-20, 0], unmasked ones in [-10, 10], row norm 10 before the shift
    unshift = got.copy()
    unshift[feas[:, 5] != 0, 0] += 10
    unshift[feas[:, 7] != 0, 2] += 10
    np.testing.assert_allclose(np.linalg.norm(unshift, axis=1), 10.0, rtol=1e-5)


def test_add_knowledge_bit_exact_with_torch_ops(cuda):
    """Same fp32 operations as F.normalize(x)*10 and the masked subtraction (reference arch.py:129-141)."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    rng = np.random.default_rng(0)
    rows = 10_000
    x = torch.from_numpy(rng.standard_normal((rows, 3)).astype(np.float32))
    x[0] = 0.0                                   # zero row: eps clamp
    feas = torch.from_numpy(_feas(rows, rng))
    got = ops.add_knowledge_kernel(x.to(cuda), feas.to(cuda)).cpu()
    exp, _ = port.add_knowledge_t(x, x[:1], feas, feas[:1])
    assert float((got - exp).abs().max()) <= 2e-6      # <= 1-2 ulp at magnitude 10 (sum order of 3 squares)
    masked0, masked2 = feas[:, 5] != 0, feas[:, 7] != 0
    assert bool((got[masked0, 0] <= 0).all()) and bool((got[masked2, 2] <= 0).all())


@pytest.mark.parametrize("m,n", [(1, 1), (5, 7), (1000, 2000), (50_000, 100_000), (2049, 4097)])
def test_basis_select_vs_oracle(cuda, m, n):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    rng = np.random.default_rng(m)
    lc = (rng.standard_normal((m, 3)) * 4).astype(np.float32)
    lv = (rng.standard_normal((n, 3)) * 4).astype(np.float32)
    status, counts = ops.basis_select(torch.from_numpy(lc).to(cuda), torch.from_numpy(lv).to(cuda), want_counts=True)
    status = status.cpu().numpy()
    exp = port.inference_gnn_np(np.concatenate([lc, lv]), m)
    agree = float(np.mean(status == exp))
    assert agree >= 0.999, agree
    # invariants asserted by the reference (val.py:118-122)
    assert int((status == 1).sum()) == m
    assert int((status[m:] == 1).sum()) == int(((status[:m] == 0) | (status[:m] == 2)).sum())
    c = counts.cpu().numpy()
    assert c[1] == m and c[0] + c[1] + c[2] == m + n and c[3] == int((status[m:] == 1).sum())
    # and against the reference's own torch ops
    exp_t = port.inference_gnn_t(torch.from_numpy(np.concatenate([lc, lv])), m).numpy()
    assert float(np.mean(status == exp_t)) >= 0.999


def test_basis_select_ties_nan_and_uint8(cuda):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    m, n = 40, 60
    lc = np.zeros((m, 3), dtype=np.float32)                 # every node ties at p1 = 1/3
    lv = np.zeros((n, 3), dtype=np.float32)
    lv[7] = np.nan                                          # NaN row -> probabilities 0 -> never basic, status 0
    lv[8] = [0.0, 5.0, 0.0]                                 # clear winner
    s = ops.basis_select(torch.from_numpy(lc).to(cuda), torch.from_numpy(lv).to(cuda), int64=False).cpu().numpy()
    assert s.dtype == np.uint8 and int((s == 1).sum()) == m
    assert s[m + 8] == 1 and s[m + 7] == 0
    # ties go to the lowest node indices: the first m-1 tied nodes (all constraints but the last one... ) are basic
    exp = port.inference_gnn_np(np.concatenate([lc, lv]), m)
    np.testing.assert_array_equal(s, exp)


def test_basis_select_k_zero_and_all(cuda):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    rng = np.random.default_rng(3)
    lc = torch.from_numpy(rng.standard_normal((10, 3)).astype(np.float32)).to(cuda)
    lv = torch.from_numpy(rng.standard_normal((20, 3)).astype(np.float32)).to(cuda)
    assert int((ops.basis_select(lc, lv, k_basic=0) == 1).sum()) == 0
    assert int((ops.basis_select(lc, lv, k_basic=30) == 1).sum()) == 30


@pytest.mark.parametrize("m,n,quant", [(1, 1, 0), (5, 7, 0), (1000, 2000, 0), (50_000, 100_000, 0), (50_000, 100_000, 8),
                                        (300_000, 500_000, 64), (700_000, 1_500_000, 0), (2049, 4097, 2)])
@pytest.mark.parametrize("int64", [False, True])
def test_one_launch_select_equals_the_launch_chain(cuda, m, n, quant, int64):
    """The cooperative one-launch kernel (keys in registers, 11+11+10-bit radix select with grid barriers) against the
    7-launch chain: identical statuses and counts, including heavy ties at the threshold (quantised logits), k = 0 / all
    and sizes above the fused path's capacity (where both calls take the chain)."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, ops
    lib = _lib.load()
    rng = np.random.default_rng(m + quant)
    lc = (rng.standard_normal((m, 3)) * 4).astype(np.float32)
    lv = (rng.standard_normal((n, 3)) * 4).astype(np.float32)
    if quant:                                                # few distinct values -> many nodes share the threshold key
        lc, lv = np.round(lc * quant / 16) / quant, np.round(lv * quant / 16) / quant
    tc, tv = torch.from_numpy(lc.astype(np.float32)).to(cuda), torch.from_numpy(lv.astype(np.float32)).to(cuda)
    for k in (None, 0, m + n, max(1, (m + n) // 3)):
        prev = lib.lpgnn_set_select_fused(1)
        try:
            torch.cuda.synchronize()
            before = lib.lpgnn_launch_count()
            s1, c1 = ops.basis_select(tc, tv, k_basic=k, int64=int64, want_counts=True)
            fused_launches = lib.lpgnn_launch_count() - before
            lib.lpgnn_set_select_fused(0)
            s0, c0 = ops.basis_select(tc, tv, k_basic=k, int64=int64, want_counts=True)
        finally:
            lib.lpgnn_set_select_fused(prev)
        assert torch.equal(s1, s0) and torch.equal(c1, c0), (k, int((s1 != s0).sum()))
        if m + n <= 1_000_000:
            assert fused_launches == 1
        kk = m if k is None else k
        assert int((s1 == 1).sum()) == kk
