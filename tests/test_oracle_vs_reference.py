"""The oracle port against the reference's own Python imported verbatim (only where /root/reference
exists, i.e. the build container; skipped on the GPU box)."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import port
from oracle.ref_import import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def R():
    return load_reference()


def test_state_dict_keys_and_param_count(R):
    torch.manual_seed(0)
    ref = R.arch.GCN_FC(8, 8, hids=1024, depth=3)
    torch.manual_seed(0)
    mine = port.PortGCN_FC(8, 8, hids=1024, depth=3)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    assert sum(p.numel() for p in ref.parameters()) == 4_237_318
    for k, v in ref.state_dict().items():
        assert torch.equal(v, mine.state_dict()[k]), k               # same init draw order


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_generator_features_match_reference(R, seed):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import features, synth
    c, b_l, A, b_u, l, u = synth.raw_lp(200, 400, 2000, seed)
    with contextlib.redirect_stdout(io.StringIO()):
        rc, rbl, rA, rbu, rl, ru = R.dataset.scaling(c.copy(), b_l.copy(), A.copy(), b_u.copy(), l.copy(), u.copy())
    rv, rcf = R.dataset.cvt_to_features(rc, rbl, rA, rbu, rl, ru)
    for impl in (port.scaling, features.scale_lp):
        c2, bl2, A2, bu2, l2, u2 = impl(c, b_l, A, b_u, l, u)
        np.testing.assert_array_equal(c2, rc)
        np.testing.assert_array_equal(A2.toarray(), rA.toarray())
        np.testing.assert_array_equal(bl2, rbl)
        np.testing.assert_array_equal(u2, ru)
    for impl in (port.cvt_to_features, features.node_features):
        v, cf = impl(rc, rbl, rA, rbu, rl, ru)
        np.testing.assert_allclose(v, rv, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(cf, rcf, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("logit_scale", [1.0, 8.0])
def test_inference_gnn_matches_reference(R, logit_scale):
    rng = np.random.default_rng(3)
    m, n = 700, 1500
    logits = (rng.standard_normal((m + n, 3)) * logit_scale).astype(np.float32)
    ref = R.val.inference_gnn(torch.from_numpy(logits), m).numpy()
    np.testing.assert_array_equal(port.inference_gnn_np(logits, m), ref)


def test_add_knowledge_matches_reference(R):
    rng = np.random.default_rng(4)
    l, r = rng.standard_normal((50, 3)).astype(np.float32), rng.standard_normal((80, 3)).astype(np.float32)
    fl, fr = rng.integers(-1, 2, (50, 8)).astype(np.float32), rng.integers(-1, 2, (80, 8)).astype(np.float32)
    a, b = R.arch.add_knowledge(torch.from_numpy(l), torch.from_numpy(r), torch.from_numpy(fl), torch.from_numpy(fr))
    c, d = port.add_knowledge_np(l, r, fl, fr)
    np.testing.assert_allclose(c, a.numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(d, b.numpy(), rtol=0, atol=2e-6)
    e, f = port.add_knowledge_t(torch.from_numpy(l), torch.from_numpy(r), torch.from_numpy(fl), torch.from_numpy(fr))
    assert torch.equal(e, a) and torch.equal(f, b)


def test_basis_file_writers_match_reference_byte_for_byte(R, tmp_path):
    """write_bas_highs / write_bas / write_sort_vars (scripts/pred_basis.py:14-67) vs the product's."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import pred_basis as mine
    rng = np.random.default_rng(7)
    m, n = 9, 17
    logits = torch.from_numpy(rng.standard_normal((m + n, 3)).astype(np.float32) * 3)
    pred = R.val.inference_gnn(logits.clone(), m).numpy()
    pred_con, pred_var = pred[:m], pred[m:]
    # a consistent named basis for write_bas: as many basic variables as non-basic constraints
    pv = np.zeros(n, dtype=np.int64); pc = np.ones(m, dtype=np.int64)
    pc[[1, 4, 6]] = [0, 2, 0]; pv[[2, 3, 11]] = 1; pv[[5, 16]] = 2
    var_nms, con_nms = [f"x{i}" for i in range(n)], [f"c{i}" for i in range(m)]
    ref_dir, my_dir = tmp_path / "ref", tmp_path / "mine"
    ref_dir.mkdir(); my_dir.mkdir()
    for mod, d in ((R.pred_basis, ref_dir), (mine, my_dir)):
        mod.write_bas_highs(str(d / "a.bas"), var_nms, con_nms, pred_var, pred_con)
        mod.write_bas(str(d / "b.bas"), var_nms, con_nms, pv, pc)
        mod.write_sort_vars(str(d / "a.bas.sort"), logits, m)
    for name in ("a.bas", "b.bas", "a.bas.sort"):
        assert (ref_dir / name).read_bytes() == (my_dir / name).read_bytes(), name
    # the split-probability form used by run() writes the same file
    p1 = torch.softmax(logits, dim=-1)[:, 1].numpy()
    mine.write_sort_vars(str(my_dir / "c.sort"), p1[m:], p1[:m])
    assert (my_dir / "c.sort").read_bytes() == (ref_dir / "a.bas.sort").read_bytes()


def test_golden_fixtures_are_reproducible_outputs_of_the_reference(R, tmp_path, monkeypatch):
    """tests/golden/*.npz are what pins the oracle and the CUDA path (the reference ships no vectors of its own):
    re-running oracle/make_golden.py on the verbatim reference must reproduce every committed array -- integer
    arrays exactly, floating-point arrays to 1e-6 (thread-count dependent summation order in MKL)."""
    import glob
    import os

    from oracle import make_golden
    committed = os.path.join(os.path.dirname(__file__), "golden")
    monkeypatch.setattr(make_golden, "OUT", str(tmp_path))
    with contextlib.redirect_stdout(io.StringIO()):
        make_golden.main()
    names = sorted(os.path.basename(f) for f in glob.glob(os.path.join(committed, "*.npz")))
    assert names == sorted(os.listdir(tmp_path)) and len(names) >= 10
    for nm in names:
        a, b = np.load(os.path.join(committed, nm)), np.load(os.path.join(tmp_path, nm))
        assert sorted(a.files) == sorted(b.files), nm
        for k in a.files:
            x, y = a[k], b[k]
            assert x.shape == y.shape and x.dtype == y.dtype, (nm, k)
            if np.issubdtype(x.dtype, np.floating):
                # gradients / logits: relative to the array's scale
                scale = max(1.0, float(np.abs(x[np.isfinite(x)]).max())) if x.size and np.isfinite(x).any() else 1.0
                np.testing.assert_allclose(y, x, rtol=0, atol=1e-5 * scale, err_msg=f"{nm}:{k}")
            else:
                np.testing.assert_array_equal(y, x, err_msg=f"{nm}:{k}")


def test_graph_construction_port_vs_reference_random_patterns(R):
    """Property test (hypothesis): LPDataset.get's to_undirected + the verbatim MyToBipartite + SparseTensor.t()
    (dataset.py:250-252, 275-332; arch.py:71) vs oracle.port on random sparsity patterns -- empty rows / columns,
    single entries, dense blocks, COO given in shuffled order -- bit-exact indices and values in both orientations."""
    import scipy.sparse as sp
    from hypothesis import given, settings, strategies as st

    from oracle.make_golden import reference_batch_via_stubs

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(m=st.integers(1, 12), n=st.integers(1, 15), density=st.floats(0.02, 0.9), seed=st.integers(0, 10_000))
    def check(m, n, density, seed):
        rng = np.random.default_rng(seed)
        mask = rng.random((m, n)) < density
        if not mask.any():
            mask[rng.integers(m), rng.integers(n)] = True
        r, c = np.nonzero(mask)
        v = rng.uniform(-1, 1, r.shape[0]).astype(np.float32)
        v[v == 0] = 0.5
        perm = rng.permutation(r.shape[0])
        A = sp.coo_matrix((v[perm], (r[perm], c[perm])), shape=(m, n))
        cf, vf = rng.standard_normal((m, 8)).astype(np.float32), rng.standard_normal((n, 8)).astype(np.float32)
        y_s, y_t = rng.integers(0, 3, m), rng.integers(0, 3, n)
        uni, batch = reference_batch_via_stubs(R, A, cf, vf, y_s, y_t)
        ei, ea = port.unipartite_edges(A.row, A.col, A.data, m)
        np.testing.assert_array_equal(ei, uni["edge_index"])
        np.testing.assert_array_equal(ea, uni["edge_attr"])
        g = port.to_bipartite(uni["edge_index"], uni["edge_attr"], uni["is_vars"])
        adj = batch.edge_index
        st_, tt = adj.storage, adj.t().storage
        np.testing.assert_array_equal(g.rowptr, st_.rowptr().numpy())
        np.testing.assert_array_equal(g.col, st_.col().numpy())
        np.testing.assert_array_equal(g.val, st_.value().numpy())
        np.testing.assert_array_equal(g.colptr, tt.rowptr().numpy())
        np.testing.assert_array_equal(g.row_csc, tt.col().numpy())
        np.testing.assert_array_equal(g.val_csc, tt.value().numpy())
        # independent check of the stand-in itself: both orientations equal scipy's canonical CSR / CSC
        ref_csr = sp.csr_matrix((A.data, (A.row, A.col)), shape=(m, n)); ref_csr.sort_indices()
        ref_csc = ref_csr.tocsc(); ref_csc.sort_indices()
        np.testing.assert_array_equal(g.col, ref_csr.indices); np.testing.assert_array_equal(g.val, ref_csr.data)
        np.testing.assert_array_equal(g.row_csc, ref_csc.indices); np.testing.assert_array_equal(g.val_csc, ref_csc.data)
        assert (batch.s_bs, batch.t_bs) == (m, n)
        np.testing.assert_array_equal(batch.x_s.numpy(), cf)
        np.testing.assert_array_equal(batch.y_t.numpy(), y_t)

    check()


def test_inference_gnn_with_ties_nan_and_extreme_k(R):
    """val.py:106-124 on inputs where torch.topk's tie order is implementation-defined (quantised logits -> many equal
    P(basic)), rows with NaN (softmax -> NaN -> 0, val.py:110) and k at its extremes.  The reference's own output must
    satisfy the definition the port (and the CUDA kernel) implement, and the two may differ only inside the tie class at
    the threshold."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(total=st.integers(2, 60), frac=st.floats(0.0, 1.0), levels=st.integers(1, 4), nan_rows=st.integers(0, 3),
           seed=st.integers(0, 10_000))
    def check(total, frac, levels, nan_rows, seed):
        rng = np.random.default_rng(seed)
        m = min(total - 1, max(1, int(round(frac * total))))          # the reference asserts 0 < #basic = m <= total
        logits = rng.integers(-levels, levels + 1, (total, 3)).astype(np.float32)
        for r in rng.choice(total, size=min(nan_rows, total), replace=False):
            logits[r, rng.integers(3)] = np.nan
        ref = R.val.inference_gnn(torch.from_numpy(logits.copy()), m).numpy()
        got = port.inference_gnn_np(logits.copy(), m)
        x = torch.softmax(torch.from_numpy(logits), dim=-1).numpy()
        x[np.isnan(x)] = 0
        p1 = x[:, 1]
        for pred in (ref, got):
            basic = pred == 1
            assert int(basic.sum()) == m
            if 0 < m < total:
                assert p1[basic].min() >= p1[~basic].max()
            np.testing.assert_array_equal(pred[~basic], np.where(x[~basic, 0] >= x[~basic, 2], 0, 2))
        thr = np.sort(p1)[::-1][m - 1]
        off_tie = p1 != thr
        np.testing.assert_array_equal(got[off_tie] == 1, ref[off_tie] == 1)
        # the port's documented tie rule: lowest node index first
        tie_idx = np.nonzero(~off_tie)[0]
        k_tie = int((got[tie_idx] == 1).sum())
        assert (got[tie_idx[:k_tie]] == 1).all() and (got[tie_idx[k_tie:]] != 1).all()

    check()


@pytest.mark.parametrize("labels_s,labels_t", [([0, 1, 1, 2, 2, 2, 1], [1, 0, 0, 2, 0, 1, 1, 0, 2]),
                                               ([1, 1, 0, 1], [0, 0, 1, 0, 0]),          # two classes: no l/u merge
                                               ([1, 2, 2, 1, 1], [0, 2, 1, 0, 2, 1])])
def test_training_losses_match_reference_train_py(R, labels_s, labels_t):
    """train.py:18-53 (`balanced`, `unbalanced`, `focal`) and utils.py:286-299 (`labels_to_balanced_weights`) of the
    verbatim reference vs the product's losses.py and the oracle port: values and gradients w.r.t. the logits."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import losses
    rng = np.random.default_rng(5)
    y_s, y_t = torch.tensor(labels_s), torch.tensor(labels_t)
    base_c = torch.from_numpy(rng.standard_normal((len(labels_s), 3)).astype(np.float32) * 4)
    base_v = torch.from_numpy(rng.standard_normal((len(labels_t), 3)).astype(np.float32) * 4)
    for y in (y_s, y_t):
        assert torch.allclose(losses.labels_to_balanced_weights(y), R.utils.labels_to_balanced_weights(y).float())

    def run(fn):
        lc, lv = base_c.clone().requires_grad_(), base_v.clone().requires_grad_()
        loss = fn(lc, lv, y_s, y_t)
        loss.backward()
        return float(loss), lc.grad.clone(), lv.grad.clone()

    for name in ("balanced", "unbalanced", "focal"):
        ref = run(getattr(R.train, name))
        mine = run(losses.LOSSES[name])
        assert abs(ref[0] - mine[0]) <= 1e-6 * max(1.0, abs(ref[0])), name
        assert torch.allclose(ref[1], mine[1], rtol=1e-5, atol=1e-7) and torch.allclose(ref[2], mine[2], rtol=1e-5, atol=1e-7), name
    ref = run(R.train.balanced)
    got = run(port.balanced_loss)
    assert abs(ref[0] - got[0]) <= 1e-6 * max(1.0, abs(ref[0])) and torch.allclose(ref[1], got[1], rtol=1e-5, atol=1e-7)


def test_host_helpers_match_reference_utils(R):
    """utils.py:256-263 (split_idxs_train_val), 301-309 (extract_fn, including the upstream 'sol' 'txt' literal that
    fuses into one suffix) vs lpgnn_b200.io_utils."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import io_utils
    for n in (1, 2, 3, 10, 37, 128):
        for seed in (0, 3):
            a, b = R.utils.split_idxs_train_val(n, seed)
            c, d = io_utils.split_idxs_train_val(n, seed)
            np.testing.assert_array_equal(a, c)
            np.testing.assert_array_equal(b, d)
    for name in ("/a/b/lp7.mps.gz", "x.y.pk", "model.sol", "notes.txt", "a.soltxt", "q.bas.sort", "plain", "d/e.lp.json",
                 "run.log.tar.gz", "k.1.2.pk"):
        assert io_utils.extract_fn(name) == R.utils.extract_fn(name), name
