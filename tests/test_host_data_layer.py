"""Host-side data layer (no GPU): processed-file format, LPDataset.get / MyToBipartite against the oracle's
restatement of reference dataset.py:229-332, loaders, split, writers, sharding."""
import os

import numpy as np
import pytest
import torch

from oracle import port

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pkg():
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import data, dataset, io_utils, pred_basis, synth, train
    return dict(data=data, dataset=dataset, io=io_utils, pred=pred_basis, synth=synth, train=train)


def test_msgpack_numpy_roundtrip(pkg, tmp_path):
    io = pkg["io"]
    obj = [np.arange(5, dtype=np.int64), np.linspace(0, 1, 7).astype(np.float32).reshape(7, 1), 12, "name",
           {"k": np.float64(3.5)}, [np.array([True, False])]]
    fn = tmp_path / "x.pk"
    io.msgpack_dump(obj, fn)
    back = io.msgpack_load(fn)
    np.testing.assert_array_equal(back[0], obj[0])
    assert back[0].dtype == np.int64 and back[1].dtype == np.float32 and back[1].shape == (7, 1)
    assert back[2] == 12 and back[3] == "name" and float(back[4]["k"]) == 3.5
    # the msgpack_numpy wire format: a map with nd/type/shape/data keys
    import msgpack
    raw = msgpack.unpackb(open(fn, "rb").read(), raw=True, strict_map_key=False)
    assert set(raw[0].keys()) == {b"nd", b"type", b"kind", b"shape", b"data"} and raw[0][b"nd"] is True


# The msgpack_numpy wire format (third-party, absent here; the reference pins no version -- 0.4.8 is the last release
# and its encode() is what utils.py:193-200 patches in) assembled BY HAND from its published encode() and the msgpack
# spec, so the reader/writer is checked against a vector this repo's encoder did not produce:
#   ndarray  -> {b'nd': True, b'type': dtype.str, b'kind': b'', b'shape': shape, b'data': raw bytes}
#   np scalar (not a Python float/int subclass) -> {b'nd': False, b'type': dtype.str, b'data': raw bytes}
MSGPACK_NUMPY_GOLDEN = bytes.fromhex(
    "93"                                                 # list of 3
    "85"                                                 # map of 5: np.arange(3, dtype='<i4')
    "c4026e64" "c3"                                      # b'nd': True
    "c40474797065" "a33c6934"                            # b'type': '<i4'
    "c4046b696e64" "c400"                                # b'kind': b''
    "c4057368617065" "9103"                              # b'shape': [3]
    "c40464617461" "c40c" "000000000100000002000000"     # b'data': 12 bytes, little endian
    "83"                                                 # map of 3: np.float32(3.5)
    "c4026e64" "c2"                                      # b'nd': False
    "c40474797065" "a33c6634"                            # b'type': '<f4'
    "c40464617461" "c404" "00006040"                     # b'data'
    "a3783a31")                                          # the str 'x:1' (names are plain msgpack strings)


def test_msgpack_numpy_wire_vector(pkg, tmp_path):
    io = pkg["io"]
    fn = tmp_path / "golden.pk"
    fn.write_bytes(MSGPACK_NUMPY_GOLDEN)
    arr, scalar, name = io.msgpack_load(fn)
    np.testing.assert_array_equal(arr, np.arange(3, dtype=np.int32))
    assert arr.dtype == np.dtype("<i4") and arr.flags.writeable            # copy=True of dataset.py:189
    assert scalar.dtype == np.float32 and float(scalar) == 3.5 and name == "x:1"
    io.msgpack_dump([np.arange(3, dtype="<i4"), np.float32(3.5), "x:1"], fn)
    assert fn.read_bytes() == MSGPACK_NUMPY_GOLDEN


def test_dataset_get_and_to_bipartite_match_oracle(pkg, tmp_path):
    ds_mod, synth = pkg["dataset"], pkg["synth"]
    root = str(tmp_path / "ds")
    ds_mod.write_synthetic_dataset(root, [(60, 130, 500), (200, 350, 1500), (5, 9, 20)], seed=3)
    ds = ds_mod.LPDataset(root, transform=ds_mod.MyToBipartite(thresh_num=np.inf), load_meta=True)
    assert len(ds) == 3 and ds.processed_file_names == ["lp0.pk", "lp1.pk", "lp2.pk"]
    for i in range(3):
        uni = ds.get(i)
        lp = synth.processed_lp(*[(60, 130, 500), (200, 350, 1500), (5, 9, 20)][i], seed=3 * 100_003 + i)
        ei, ea = port.unipartite_edges(lp.row, lp.col, lp.a_data, lp.m)            # dataset.py:250-252
        np.testing.assert_array_equal(uni.edge_index.numpy(), ei)
        np.testing.assert_array_equal(uni.edge_attr.numpy(), ea)
        assert uni.num_nodes == lp.m + lp.n and int(uni.is_vars.sum()) == lp.n
        ref = port.to_bipartite(ei, ea, uni.is_vars.numpy())                        # dataset.py:283-304
        batch = ds[i]
        g = batch.edge_index
        assert g.sparse_sizes() == (lp.m, lp.n) and g.nnz() == lp.nnz and not g.is_cuda and g._sorted_hint
        r, c, v = g._coo
        np.testing.assert_array_equal(c.numpy(), ref.col.astype(np.int32))
        np.testing.assert_array_equal(v.numpy(), ref.val)
        np.testing.assert_array_equal(r.numpy(), np.repeat(np.arange(lp.m), np.diff(ref.rowptr)).astype(np.int32))
        np.testing.assert_array_equal(batch.x_s.numpy(), lp.c_feas)
        np.testing.assert_array_equal(batch.x_t.numpy(), lp.v_feas)
        np.testing.assert_array_equal(batch.y_t.numpy(), lp.y_t)
        assert (batch.s_bs, batch.t_bs, batch.bs) == (lp.m, lp.n, lp.m + lp.n)
        assert not hasattr(batch, "x") and not hasattr(batch, "is_vars") and not hasattr(batch, "edge_attr")
        assert batch.con_nms[:2] == ["c0", "c1"]


def test_golden_unipartite_graph_through_product_transform(pkg):
    """The reference's own unipartite tensors (golden) through the product's MyToBipartite."""
    ds_mod, data = pkg["dataset"], pkg["data"]
    z = np.load(os.path.join(GOLD, "graph_small_300x600.npz"))
    m, n = int(z["m"]), int(z["n"])
    uni = data.Data(x=torch.zeros(m + n, 8), y=torch.zeros(m + n, dtype=torch.long), is_vars=torch.from_numpy(z["is_vars"]),
                    edge_index=torch.from_numpy(z["uni_edge_index"]), edge_attr=torch.from_numpy(z["uni_edge_attr"]),
                    num_nodes=m + n)
    batch = ds_mod.MyToBipartite()(uni)
    r, c, v = batch.edge_index._coo
    np.testing.assert_array_equal(c.numpy(), z["col"].astype(np.int32))
    np.testing.assert_array_equal(v.numpy().view(np.uint32), z["val"].view(np.uint32))
    assert (batch.s_bs, batch.t_bs) == (int(z["s_bs"]), int(z["t_bs"]))


def test_loader_split_and_process(pkg, tmp_path):
    ds_mod, data, io = pkg["dataset"], pkg["data"], pkg["io"]
    root = str(tmp_path / "ds")
    ds_mod.write_synthetic_dataset(root, [(20 + i, 50 + i, 120) for i in range(10)], seed=1)
    ds = ds_mod.LPDataset(root, transform=ds_mod.MyToBipartite())
    tr, va = io.split_train_val(ds, seed=0)
    assert len(tr) == 7 and len(va) == 3 and sorted(list(tr.indices()) + list(va.indices())) == list(range(10))
    np.random.seed(0)
    perm = np.random.permutation(10)                                   # utils.py:259-262 semantics
    assert list(tr.indices()) == sorted(perm[:7].tolist())
    # dataset.py:107-157 size table: whole dataset cached in size.json, rows restricted to the (sub-)dataset
    info = ds.cache_size_info()
    assert set(info.columns) == {"idx", "nedges", "nnodes", "fn", "ncons", "nvars", "density", "num_basis_vars"}
    assert len(info) == 10 and os.path.exists(os.path.join(root, "size.json"))
    b0 = ds[0]
    r0 = info.loc[0]
    assert (r0.ncons, r0.nvars, r0.nnodes) == (20, 50, 70) and r0.nedges == b0.edge_index.nnz()
    assert abs(r0.density - r0.nedges / (20 * 50)) < 1e-15 and r0.num_basis_vars == int((b0.y_t == 1).sum())
    assert not r0.fn.endswith(".pk")
    assert list(va.cache_size_info().index) == list(va.indices())
    dumped = ds.dump_size_info(os.path.join(root, "size_split.json"))
    assert sorted(dumped.index[dumped.split == "val"]) == sorted(va.indices())
    assert ds.dump_size_info(os.path.join(root, "size_split.json")) is None          # exists: left alone
    loader = data.DataLoader(va, batch_size=1, shuffle=False, num_workers=0)
    names = [b.processed_path[0] for b in loader]
    assert len(names) == 3 and all(isinstance(b, str) for b in names)
    # raw -> processed (LPDataset.process) on a raw LP with names/labels
    synth = pkg["synth"]
    raw_root = str(tmp_path / "raw_ds")
    os.makedirs(os.path.join(raw_root, "raw"))
    c, b_l, A, b_u, l, u = synth.raw_lp(30, 70, 200, seed=4)
    coo = A.tocoo()
    io.msgpack_dump([c, b_l, (coo.row, coo.col, coo.data), b_u, l, u, np.ones(30, dtype=np.int64), np.ones(70, dtype=np.int64),
                     [f"r{i}" for i in range(30)], [f"v{i}" for i in range(70)]], os.path.join(raw_root, "raw", "a.pk"))
    ds2 = ds_mod.LPDataset(raw_root, transform=ds_mod.MyToBipartite(), load_meta=True)
    ds2.process()
    b = ds2[0]
    assert b.x_s.shape == (30, 8) and b.x_t.shape == (70, 8) and b.var_nms[0] == "v0"
    assert float(b.x_s.abs().max()) <= 1.0


def test_bas_writers_and_extract_fn(pkg, tmp_path):
    pred, io = pkg["pred"], pkg["io"]
    fn = tmp_path / "out" / "lp7.bas"
    pred.write_bas_highs(str(fn), None, None, np.array([1, 0, 2, 1], dtype=np.uint8), np.array([0, 1], dtype=np.uint8))
    assert open(fn).read() == "HIGHS v1\nValid\n# Columns 4\n1 0 2 1\n# Rows 2\n0 1\n"     # pred_basis.py:19-23
    pred.write_sort_vars(str(fn) + ".sort", np.array([0.5, 0.25], dtype=np.float32), np.array([0.125], dtype=np.float32))
    assert open(str(fn) + ".sort").read().splitlines()[0] == "2 "
    assert io.extract_fn("/a/b/lp7.mps.gz") == "lp7" and io.extract_fn("x.y.pk") == "x.y"


def test_shard_indices_partition(pkg):
    io = pkg["io"]
    for world in (1, 2, 3, 8):
        parts = [io.shard_indices(37, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(37))
        w = np.random.default_rng(0).integers(1, 1000, 37)
        parts = [io.shard_indices(37, r, world, weights=w) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(37))
        loads = [w[p].sum() for p in parts]
        assert max(loads) - min(loads) <= w.max()                       # LPT balance
        parts = [io.shard_indices(37, r, world, weights=w, equal_counts=True) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(37))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1         # snake deal: equal counts ...
        loads = [w[p].sum() for p in parts]
        assert max(loads) - min(loads) <= w.max()                       # ... and near-equal weight


def test_losses_match_reference_semantics(pkg):
    """losses.py (train.py:32-53, utils.py:286-299) vs the oracle restatement, on CPU tensors."""
    from lpgnn_b200 import losses
    rng = np.random.default_rng(0)
    for labels in ([0, 1, 1, 2, 2, 2], [1, 1, 0], [1, 2, 2, 2], [1, 1, 1], [0, 2], [0, 1, 2]):
        y = torch.tensor(labels)
        assert torch.allclose(losses.labels_to_balanced_weights(y), port.labels_to_balanced_weights(y))
    m, n = 40, 90
    lc = torch.from_numpy(rng.standard_normal((m, 3)).astype(np.float32))
    lv = torch.from_numpy(rng.standard_normal((n, 3)).astype(np.float32))
    ys, yt = torch.from_numpy(rng.integers(0, 3, m)), torch.from_numpy(rng.integers(1, 3, n))
    assert torch.allclose(losses.balanced(lc, lv, ys, yt), port.balanced_loss(lc, lv, ys, yt), rtol=1e-6)
    ce = torch.nn.functional.cross_entropy(torch.cat((lc, lv)), torch.cat((ys, yt)))
    assert torch.allclose(losses.unbalanced(lc, lv, ys, yt), ce)
    assert torch.allclose(losses.focal(lc, lv, ys, yt), (1 - torch.exp(-ce)) ** 2 * ce)


def test_oracle_induced_subgraph_and_khop_expansion():
    """(f-4) the oracle's deterministic restatement of NeighborLoader(directed=False) + MyToBipartite on a hand case."""
    import scipy.sparse as sp
    from oracle import port
    A = sp.csr_matrix(np.array([[1., 0, 2, 0], [0, 3, 0, 0], [0, 0, 4, 5]], dtype=np.float32))
    cn, vn = port.khop_full_neighbourhood(A, [1], [], 2)          # c1 -> v1 -> (no new constraint)
    assert list(cn) == [1] and list(vn) == [1]
    cn, vn = port.khop_full_neighbourhood(A, [0], [], 2)          # c0 -> v0, v2 -> c2
    assert list(cn) == [0, 2] and list(vn) == [0, 2]
    rowptr, col, val = port.induced_bipartite_subgraph(A, [2, 0], [2, 0])   # relabelled: rows (c2, c0), cols (v2, v0)
    assert list(rowptr) == [0, 1, 3] and list(col) == [0, 0, 1] and list(val) == [4., 2., 1.]


def test_sorted_claim_is_only_made_for_verified_edge_order(pkg):
    """MyToBipartite hands `is_sorted=True` to the device build only after checking the (row, col) order on the host: a
    hand-built unipartite graph with constraint-first nodes but UNSORTED edges must take the sort path (a false claim
    would leave rowptr entries unwritten, graph_build.cu prep_sorted_kernel)."""
    from lpgnn_b200.dataset import MyToBipartite, UnipartiteData
    m, n = 3, 4
    row = np.array([0, 0, 1, 2, 2]); col = np.array([1, 3, 0, 2, 3]); val = np.arange(1, 6, dtype=np.float32)

    def uni(order):
        r, c, v = row[order], col[order] + m, val[order]
        ei = torch.from_numpy(np.stack([np.concatenate([r, c]), np.concatenate([c, r])]))
        ea = torch.from_numpy(np.concatenate([v, v]))
        is_vars = torch.zeros(m + n, dtype=torch.long); is_vars[m:] = 1
        return UnipartiteData(x=torch.zeros(m + n, 8), y=torch.zeros(m + n, dtype=torch.long), is_vars=is_vars,
                              edge_index=ei, edge_attr=ea, num_nodes=m + n)

    g_sorted = MyToBipartite()(uni(np.arange(5))).edge_index
    g_unsorted = MyToBipartite()(uni(np.array([3, 0, 4, 2, 1]))).edge_index
    assert g_sorted._sorted_hint is True and g_unsorted._sorted_hint is False


def test_basis_file_readers_round_trip_the_writers(pkg, tmp_path):
    """read_bas / read_bas_highs (reference scripts/cvt_to_pkl.py:166-209, the readers behind val.validation_wrt_converged)
    against the writers of pred_basis.py: named (MPS-style) and HiGHS files give back the statuses that were written."""
    pred = pkg["pred"]
    rng = np.random.default_rng(4)
    m, n = 7, 12
    st = np.zeros(m + n, dtype=np.int64)
    st[rng.permutation(m + n)[:m]] = 1                                # exactly m basic nodes
    free = np.flatnonzero(st == 0)
    st[free[rng.random(len(free)) < 0.5]] = 2
    pc, pv = st[:m], st[m:]
    con_nms, var_nms = [f"c{i}" for i in range(m)], [f"x{j}" for j in range(n)]
    pred.write_bas(str(tmp_path / "named.bas"), var_nms, con_nms, pv, pc)
    rc, rv = pred.read_bas(str(tmp_path / "named.bas"), con_nms, var_nms)
    np.testing.assert_array_equal(rc, pc)
    np.testing.assert_array_equal(rv, pv)
    pred.write_bas_highs(str(tmp_path / "h.bas"), var_nms, con_nms, pv, pc)
    for reader in (pred.read_bas, pred.read_bas_highs):
        rc, rv = reader(str(tmp_path / "h.bas"))[:2] if reader is pred.read_bas_highs else reader(str(tmp_path / "h.bas"), con_nms, var_nms)
        np.testing.assert_array_equal(rc, pc)
        np.testing.assert_array_equal(rv, pv)
    # a hand-written file in the solver's format: BS / LL lines and unnamed entries take the defaults
    (tmp_path / "hand.bas").write_text("NAME x\n XU x1 c0 \n LL x2 \n BS x3 \n UL x4 \nENDATA\n")
    rc, rv = pred.read_bas(str(tmp_path / "hand.bas"), ["c0", "c1"], ["x0", "x1", "x2", "x3", "x4"])
    assert list(rc) == [2, 1] and list(rv) == [0, 1, 0, 1, 2]
