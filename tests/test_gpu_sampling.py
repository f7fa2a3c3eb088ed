"""(f-4) sampled-subgraph path (``lpgnn_sample_mark`` / ``lpgnn_induced_count`` / ``lpgnn_induced_fill``,
``lpgnn_b200.sampling``) against the oracle's restatement of NeighborLoader(directed=False) + MyToBipartite
(reference train.py:107-116, val.py:14-36, dataset.py:275-332)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _resident(m, n, seed, dev, structure="staircase"):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import synth
    from lpgnn_b200.graph import BipartiteCSR
    from lpgnn_b200.sampling import ResidentLP
    lp = synth.processed_lp(m, n, 5 * n, seed, structure)
    g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), m, n, dev, is_sorted=True)
    res = ResidentLP(g, torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev),
                     torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev))
    A = sp.csr_matrix((lp.a_data.astype(np.float32), (lp.row, lp.col)), shape=(m, n))
    return lp, res, A


def _check_batch_is_induced(batch, A):
    cn, vn = batch.n_id_s.cpu().numpy(), batch.n_id_t.cpu().numpy()
    assert len(np.unique(cn)) == len(cn) and len(np.unique(vn)) == len(vn)
    rowptr, col, val = port.induced_bipartite_subgraph(A, cn, vn)
    g = batch.edge_index
    np.testing.assert_array_equal(g.rowptr.cpu().numpy(), rowptr)              # bit-exact structure and values
    np.testing.assert_array_equal(g.col.cpu().numpy(), col)
    np.testing.assert_array_equal(g.val.cpu().numpy().view(np.uint32), val.view(np.uint32))
    assert g.m == len(cn) and g.n == len(vn)


def test_full_neighbourhood_batches_match_oracle_expansion(cuda):
    from lpgnn_b200.sampling import NeighborSubgraphLoader
    m, n = 3000, 6000
    lp, res, A = _resident(m, n, 3, cuda)
    loader = NeighborSubgraphLoader(res, [-1, -1], batch_size=2500, shuffle=False)
    assert len(loader) == 4
    seen = 0
    for b, batch in enumerate(loader):
        seeds = np.arange(b * 2500, min((b + 1) * 2500, m + n))
        cs, vs = seeds[seeds < m], seeds[seeds >= m] - m
        cn, vn = port.khop_full_neighbourhood(A, cs, vs, 2)
        np.testing.assert_array_equal(batch.n_id_s.cpu().numpy(), cn)
        np.testing.assert_array_equal(batch.n_id_t.cpu().numpy(), vn)
        assert (batch.s_bs, batch.t_bs, batch.bs) == (len(cs), len(vs), len(seeds))
        _check_batch_is_induced(batch, A)
        np.testing.assert_array_equal(batch.x_s.cpu().numpy(), lp.c_feas[cn])
        np.testing.assert_array_equal(batch.y_t.cpu().numpy(), lp.y_t[vn])
        seen += batch.bs
    assert seen == m + n


@pytest.mark.parametrize("structure", ["staircase", "uniform"])
def test_full_neighbourhood_logits_equal_full_graph_logits(cuda, structure):
    """The reference's own check (val.py:44-47): with num_neighbors=[-1]*depth the seed logits are the full-graph ones."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch
    from lpgnn_b200.data import Data
    from lpgnn_b200.sampling import NeighborSubgraphLoader
    m, n = 2000, 4000
    lp, res, A = _resident(m, n, 5, cuda, structure)
    torch.manual_seed(1)
    model = arch.GCN_FC(8, 8, hids=64, depth=3).to(cuda).eval()
    with torch.no_grad():
        full_c, full_v = model(Data(x_s=res.x_s, x_t=res.x_t, edge_index=res.graph))
        lc, lv = [], []
        for batch in NeighborSubgraphLoader(res, [-1, -1], batch_size=1700, shuffle=False):
            c, v = model(batch)
            lc.append(c[:batch.s_bs]); lv.append(v[:batch.t_bs])
        lc, lv = torch.cat(lc), torch.cat(lv)
    assert lc.shape == full_c.shape and lv.shape == full_v.shape
    # same sums in a different neighbour order after relabelling: fp32 rounding only (relative to the row norm 10)
    assert float((lc - full_c).abs().max()) < 1e-3 and float((lv - full_v).abs().max()) < 1e-3


def test_fanout_sampling_properties(cuda):
    from lpgnn_b200.sampling import NeighborSubgraphLoader
    m, n, fan = 2000, 4000, 6
    lp, res, A = _resident(m, n, 7, cuda, "uniform")
    At = A.T.tocsr()
    loader = NeighborSubgraphLoader(res, [fan, fan], batch_size=1000, shuffle=True, drop_last=True, seed=11)
    assert len(loader) == 6
    batches = list(loader)
    all_seeds = []
    for batch in batches:
        cn, vn = batch.n_id_s.cpu().numpy(), batch.n_id_t.cpu().numpy()
        _check_batch_is_induced(batch, A)
        assert batch.bs == 1000 and batch.s_bs + batch.t_bs == 1000
        all_seeds += list(cn[:batch.s_bs]) + list(vn[:batch.t_bs] + m)
        # growth bound: every hop adds at most `fan` neighbours per frontier node
        assert len(cn) + len(vn) <= 1000 * (1 + fan + fan * fan)
        # every sampled non-seed node is adjacent to a sampled node of the other side (it was reached through an edge)
        sub = A[cn][:, vn]
        assert (np.diff(sub.tocsr().indptr)[batch.s_bs:] > 0).all()
        assert (np.diff(sub.tocsc().indptr)[batch.t_bs:] > 0).all()
    assert len(set(all_seeds)) == 6000                        # seeds of an epoch are distinct (a permutation prefix)
    # reproducible: same seed -> same batches; the next epoch draws differently
    again = list(NeighborSubgraphLoader(res, [fan, fan], batch_size=1000, shuffle=True, drop_last=True, seed=11))
    for a, b in zip(batches, again):
        assert torch.equal(a.n_id_s, b.n_id_s) and torch.equal(a.n_id_t, b.n_id_t)
    second_epoch = list(loader)
    assert not torch.equal(second_epoch[0].n_id_s, batches[0].n_id_s)


def test_sample_mark_takes_exactly_fanout_distinct_neighbours(cuda):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import sampling
    m, n = 500, 800
    lp, res, A = _resident(m, n, 9, cuda, "uniform")
    csr, _ = res.graph.views()
    deg = np.diff(A.indptr)
    for fan in (1, 3, 6):
        for f in np.random.default_rng(fan).integers(0, m, 20):
            marks = torch.zeros(n, dtype=torch.uint8, device=cuda)
            sampling._sample_mark(csr, torch.tensor([f], dtype=torch.int32, device=cuda), fan, 1234, marks)
            got = np.nonzero(marks.cpu().numpy())[0]
            assert len(got) == min(deg[f], fan)
            assert set(got) <= set(A[f].indices)
    # the choice is spread over the row (not always the first `fan` entries)
    f = int(np.argmax(deg))
    picks = set()
    for s in range(40):
        marks = torch.zeros(n, dtype=torch.uint8, device=cuda)
        sampling._sample_mark(csr, torch.tensor([f], dtype=torch.int32, device=cuda), 2, s, marks)
        picks |= set(np.nonzero(marks.cpu().numpy())[0])
    assert len(picks) >= min(deg[f], 5)


def test_train_and_validate_on_lp_above_threshold(cuda, tmp_path):
    """train.py / val.model_inference_with_batch with an LP above edge_num_thresh go through the sampled path."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch, dataset, train, val
    from lpgnn_b200.data import Data
    root = str(tmp_path / "ds")
    dataset.write_synthetic_dataset(root, [(300, 600, 3000), (320, 640, 3200), (280, 560, 2800), (310, 620, 3100)], seed=3)
    args = train.parse_args(["--dataset_processed_prefix", root, "--arch", "GCN_FC(8,8,hids=32,depth=3)", "--epochs", "2",
                             "--edge_num_thresh", "100", "--batch_size", "256", "--log_dir", str(tmp_path / "run"),
                             "--log_every", "1"])
    model, history = train.run_exp(args)
    assert len(history) >= 2 * 2 * 3 and all(np.isfinite(h["loss"]) for h in history)    # >= 3 mini-batches per LP
    ds = dataset.LPDataset(root, dataset.MyToBipartite(thresh_num=100))
    g = ds[0]
    assert not hasattr(g, "x_s")                                                 # stayed unipartite (above threshold)
    args.batch_size = 400
    lc, lv = val.model_inference_with_batch(model, g, args)
    ncons = int((g.is_vars == 0).sum())
    assert lc.shape == (ncons, 3) and lv.shape == (g.num_nodes - ncons, 3)
    full = dataset.MyToBipartite(thresh_num=np.inf)(ds.get(0))
    fc, fv = val.model_inference_with_batch(model, full, args)
    assert float((lc - fc).abs().max()) < 1e-3 and float((lv - fv).abs().max()) < 1e-3


def test_induced_fill_writes_only_its_output(cuda):
    """Canary regions around the COO arrays written by lpgnn_induced_fill (no compute-sanitizer on the GPU pool)."""
    from lpgnn_b200 import _lib
    lib = _lib.load()
    m, n = 1500, 3000
    lp, res, A = _resident(m, n, 13, cuda, "uniform")
    (ptr_, idx, val, _), _ = res.graph.views()
    rng = np.random.default_rng(0)
    cons = torch.from_numpy(rng.permutation(m)[:700].astype(np.int32)).to(cuda)
    vars_ = rng.permutation(n)[:1100]
    map_v = torch.full((n,), -1, dtype=torch.int32, device=cuda)
    map_v[torch.from_numpy(vars_).to(cuda)] = torch.arange(len(vars_), dtype=torch.int32, device=cuda)
    counts = torch.zeros(700, dtype=torch.int32, device=cuda)
    _lib.check(lib.lpgnn_induced_count(ptr_.data_ptr(), idx.data_ptr(), cons.data_ptr(), 700, map_v.data_ptr(), counts.data_ptr(),
                                       _lib.stream_ptr()), "count")
    csum = torch.cumsum(counts.long(), 0)
    z = int(csum[-1])
    offsets = (csum - counts.long()).contiguous()
    pad = 256
    bufs = [torch.full((z + 2 * pad,), -7, dtype=torch.int32, device=cuda) for _ in range(2)] + \
           [torch.full((z + 2 * pad,), -7.0, dtype=torch.float32, device=cuda)]
    _lib.check(lib.lpgnn_induced_fill(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), cons.data_ptr(), 700, map_v.data_ptr(),
                                      offsets.data_ptr(), bufs[0][pad:].data_ptr(), bufs[1][pad:].data_ptr(),
                                      bufs[2][pad:].data_ptr(), _lib.stream_ptr()), "fill")
    torch.cuda.synchronize()
    for b in bufs:
        assert bool((b[:pad] == -7).all()) and bool((b[pad + z:] == -7).all())
    sub = A[cons.cpu().numpy()][:, vars_]
    assert z == sub.nnz
    assert bool((bufs[0][pad:pad + z] >= 0).all()) and int(bufs[1][pad:pad + z].max()) < len(vars_)
