"""The oracle port (oracle/port.py) against the golden vectors produced by the reference's own
Python (oracle/make_golden.py ran /root/reference verbatim).  Runs everywhere, no GPU."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import port

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["tiny_5x7", "small_300x600", "c1_1000x2000"]


def _load(kind, name):
    return np.load(os.path.join(GOLD, f"{kind}_{name}.npz"))


@pytest.mark.parametrize("name", NAMES + ["bounds_40x90"])
def test_scaling_and_features_match_reference(name):
    z = _load("lp_features", name)
    m, n = z["in_shape"]
    A = sp.csr_matrix((z["in_A_data"], z["in_A_indices"], z["in_A_indptr"]), shape=(m, n))
    c, b_l, A2, b_u, l, u = port.scaling(z["in_c"], z["in_b_l"], A, z["in_b_u"], z["in_l"], z["in_u"])
    A2.sort_indices()
    for got, key in ((c, "out_c"), (b_l, "out_b_l"), (b_u, "out_b_u"), (l, "out_l"), (u, "out_u"), (A2.data, "out_A_data")):
        np.testing.assert_array_equal(got, z[key])                 # float64, same operations -> identical bits
    np.testing.assert_array_equal(A2.indices, z["out_A_indices"])
    np.testing.assert_array_equal(A2.indptr, z["out_A_indptr"])
    assert np.abs(A2.data).max() <= 1 and np.abs(c).max() <= 1      # dataset.py:235-238
    v_feas, c_feas = port.cvt_to_features(c, b_l, A2, b_u, l, u)
    np.testing.assert_allclose(v_feas, z["out_v_feas"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(c_feas, z["out_c_feas"], rtol=1e-12, atol=1e-15)
    # tag columns are exact
    for col in (5, 7):
        np.testing.assert_array_equal(v_feas[:, col], z["out_v_feas"][:, col])
        np.testing.assert_array_equal(c_feas[:, col], z["out_c_feas"][:, col])


@pytest.mark.parametrize("name", NAMES)
def test_graph_construction_matches_reference_bit_exact(name):
    z = _load("graph", name)
    g = port.to_bipartite(z["uni_edge_index"], z["uni_edge_attr"], z["is_vars"])
    assert (g.m, g.n) == (int(z["m"]), int(z["n"]))
    for key in ("rowptr", "col", "colptr", "row_csc", "csr2csc"):
        np.testing.assert_array_equal(getattr(g, key), z[key])
    np.testing.assert_array_equal(g.val.view(np.uint32), z["val"].view(np.uint32))
    np.testing.assert_array_equal(g.val_csc.view(np.uint32), z["val_csc"].view(np.uint32))
    # the unipartite edge list itself (LPDataset.get + to_undirected) from the processed COO
    counts = np.diff(z["rowptr"])
    row = np.repeat(np.arange(g.m), counts)
    ei, ea = port.unipartite_edges(row, z["col"], z["val"], g.m)
    np.testing.assert_array_equal(ei, z["uni_edge_index"])
    np.testing.assert_array_equal(ea, z["uni_edge_attr"])
    # canonical order == scipy sort_indices order
    s = sp.csr_matrix((z["val"], z["col"], z["rowptr"]), shape=(g.m, g.n))
    assert s.has_sorted_indices or (s.sort_indices() or True)


def _port_model(z):
    hids, depth = int(z["hids"]), int(z["depth"])
    model = port.PortGCN_FC(8, 8, hids=hids, depth=depth)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    assert list(sd.keys()) == list(model.state_dict().keys())       # state_dict key contract
    model.load_state_dict(sd)
    return model, hids, depth


@pytest.mark.parametrize("name", NAMES)
def test_forward_logits_and_basis_match_reference(name):
    zg, zm = _load("graph", name), _load("model", name)
    model, hids, depth = _port_model(zm)
    g = port.to_bipartite(zg["uni_edge_index"], zg["uni_edge_attr"], zg["is_vars"])
    x_s, x_t = torch.from_numpy(zg["x_s"]), torch.from_numpy(zg["x_t"])
    model.eval()
    with torch.no_grad():
        lc, lv = model(x_s, x_t, port.TorchGraph(g))
    np.testing.assert_allclose(lc.numpy(), zm["logits_cons"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(lv.numpy(), zm["logits_vars"], rtol=0, atol=2e-5)
    # numpy sequential-order model and its float64 twin agree with the reference to fp32 rounding
    nc, nv = port.gcn_fc_forward_np(model.state_dict(), zg["x_s"], zg["x_t"], g, depth)
    np.testing.assert_allclose(nc, zm["logits_cons"], rtol=0, atol=1e-4)
    dc, dv = port.gcn_fc_forward_np(model.state_dict(), zg["x_s"], zg["x_t"], g, depth, acc_dtype=np.float64)
    assert np.abs(dv - zm["logits_vars"]).max() / 10 < 1e-5
    m = g.m
    pred = port.inference_gnn_np(np.concatenate([zm["logits_cons"], zm["logits_vars"]]), m)
    assert np.mean(pred == zm["pred"]) >= 0.999
    assert int((pred == 1).sum()) == m                               # val.py:119
    assert int((pred[m:] == 1).sum()) == int((pred[:m] != 1).sum())  # val.py:121-122
    pred_t = port.inference_gnn_t(torch.from_numpy(np.concatenate([zm["logits_cons"], zm["logits_vars"]])), m)
    np.testing.assert_array_equal(pred_t.numpy(), zm["pred"])


@pytest.mark.parametrize("name", NAMES)
def test_balanced_loss_and_gradients_match_reference(name):
    zg, zm = _load("graph", name), _load("model", name)
    model, hids, depth = _port_model(zm)
    g = port.to_bipartite(zg["uni_edge_index"], zg["uni_edge_attr"], zg["is_vars"])
    model.eval()
    lc, lv = model(torch.from_numpy(zg["x_s"]), torch.from_numpy(zg["x_t"]), port.TorchGraph(g))
    loss = port.balanced_loss(lc, lv, torch.from_numpy(zg["y_s"]), torch.from_numpy(zg["y_t"]))
    assert abs(loss.item() - float(zm["loss"])) < 1e-4 * max(1.0, abs(float(zm["loss"])))
    loss.backward()
    for k, p in model.named_parameters():
        ref = zm[f"g::{k}"]
        scale = max(np.abs(ref).max(), 1e-6)
        assert np.abs(p.grad.numpy() - ref).max() / scale < 1e-3, k


def test_knowledge_mask_is_bit_exact_and_labels_consistent():
    """add_knowledge mask columns and the dataset.py:203-207 label invariants on the golden inputs."""
    for name in NAMES:
        zg, zm = _load("graph", name), _load("model", name)
        for logits, feas, y in ((zm["logits_cons"], zg["x_s"], zg["y_s"]), (zm["logits_vars"], zg["x_t"], zg["y_t"])):
            lo, up = feas[:, 5] != 0, feas[:, 7] != 0
            assert (y[lo] != 0).all() and (y[up] != 2).all()
            assert (logits[lo, 0] <= 0).all() and (logits[up, 2] <= 0).all()
            un = logits.copy()
            un[lo, 0] += 10
            un[up, 2] += 10
            np.testing.assert_allclose(np.linalg.norm(un, axis=1), 10, rtol=1e-5)


def test_standins_against_independent_dense_float64_model():
    """The third-party stand-ins (SparseTensor / spmm / GraphConv) checked against dense float64 algebra,
    so an error in the stand-ins themselves cannot hide in the golden vectors (SURVEY 8c iii)."""
    from oracle import pyg_standins as S
    rng = np.random.default_rng(0)
    m, n, z = 23, 31, 150
    key = np.unique(rng.integers(0, m, z) * n + rng.integers(0, n, z))
    rng.shuffle(key)
    row, col = torch.from_numpy(key // n), torch.from_numpy(key % n)
    val = torch.from_numpy(rng.uniform(-1, 1, key.shape[0]).astype(np.float32))
    st = S.SparseTensor.from_edge_index(torch.stack([row, col]), val, (m, n))
    dense = torch.zeros(m, n, dtype=torch.float64)
    dense[row, col] = val.double()
    assert torch.equal(st.to_dense(), dense) and torch.equal(st.t().to_dense(), dense.T)
    x = torch.from_numpy(rng.standard_normal((n, 5)).astype(np.float32))
    assert float((S.spmm_sum(st, x).double() - dense @ x.double()).abs().max()) < 1e-5
    assert float((S.spmm_sum_sequential(st, x).double() - dense @ x.double()).abs().max()) < 1e-5
    conv = S.GraphConv((5, 4), 6)
    xd = torch.from_numpy(rng.standard_normal((m, 4)).astype(np.float32))
    out = conv((x, xd), st)
    exp = (dense @ x.double()) @ conv.lin_rel.weight.double().T + conv.lin_rel.bias.double() + xd.double() @ conv.lin_root.weight.double().T
    assert float((out.double() - exp).abs().max()) < 1e-5
    # backward wrt the dense operand is the transposed product
    x.requires_grad_(True)
    S.spmm_sum(st, x).sum().backward()
    assert float((x.grad.double() - dense.T @ torch.ones(m, 5, dtype=torch.float64)).abs().max()) < 1e-5
