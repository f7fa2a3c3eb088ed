"""(f-2) scaling + node features on the device (``lpgnn_lp_features``) against the golden vectors made by the
reference's own ``dataset.scaling`` / ``dataset.cvt_to_features`` (dataset.py:23-96) and against the oracle port
at larger sizes.  Bars: scaled LP (float64) bit-exact; tag columns and degree columns exact; cosine columns
within 1e-6 of the float32-cast reference (their vector norms are summed in a different order than numpy's
pairwise sum, a last-bits float64 difference before the cast)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import port

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run(c, b_l, A, b_u, l, u, dev):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import features
    g, x_s, x_t, scaled = features.prepare_lp_device(c, b_l, A, b_u, l, u, dev)
    torch.cuda.synchronize()
    g.check()
    return g, x_s.cpu().numpy(), x_t.cpu().numpy(), {k: v.cpu().numpy() for k, v in scaled.items()}


def _compare(g, x_s, x_t, scaled, ref):
    for k in ("c", "b_l", "b_u", "l", "u"):
        np.testing.assert_array_equal(scaled[k], ref[k], err_msg=k)        # float64, bit-exact (inf == inf)
    np.testing.assert_array_equal(scaled["A"], ref["A"].data)
    val32 = ref["A"].data.astype(np.float32)
    np.testing.assert_array_equal(g.val.cpu().numpy(), val32)
    csc = ref["A"].tocsc()
    csc.sort_indices()
    np.testing.assert_array_equal(g.val_csc.cpu().numpy(), csc.data.astype(np.float32))
    for got, want in ((x_t, ref["v_feas"]), (x_s, ref["c_feas"])):
        want = want.astype(np.float32)
        for col in (1, 4, 5, 6, 7):                                          # degree, bound values, tags: exact
            np.testing.assert_array_equal(got[:, col], want[:, col], err_msg=f"col {col}")
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    np.testing.assert_array_equal(x_t[:, 0], ref["v_feas"][:, 0].astype(np.float32))   # c_j


@pytest.mark.parametrize("name", ["tiny_5x7", "small_300x600", "bounds_40x90", "c1_1000x2000"])
def test_device_features_match_reference_golden(cuda, name):
    z = np.load(os.path.join(GOLD, f"lp_features_{name}.npz"))
    m, n = z["in_shape"]
    A = sp.csr_matrix((z["in_A_data"], z["in_A_indices"], z["in_A_indptr"]), shape=(m, n))
    g, x_s, x_t, scaled = _run(z["in_c"], z["in_b_l"], A, z["in_b_u"], z["in_l"], z["in_u"], cuda)
    A_ref = sp.csr_matrix((z["out_A_data"], z["out_A_indices"], z["out_A_indptr"]), shape=(m, n))
    ref = {"c": z["out_c"], "b_l": z["out_b_l"], "b_u": z["out_b_u"], "l": z["out_l"], "u": z["out_u"], "A": A_ref,
           "v_feas": z["out_v_feas"], "c_feas": z["out_c_feas"]}
    _compare(g, x_s, x_t, scaled, ref)


@pytest.mark.parametrize("m,n,seed", [(2000, 4000, 11), (20_000, 40_000, 12), (1, 3, 13)])
def test_device_features_match_oracle_port(cuda, m, n, seed):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import synth
    c, b_l, A, b_u, l, u = synth.raw_lp(m, n, 5 * n, seed)
    g, x_s, x_t, scaled = _run(c, b_l, A, b_u, l, u, cuda)
    c2, bl2, A2, bu2, l2, u2 = port.scaling(c, b_l, A, b_u, l, u)
    A2 = sp.csr_matrix(A2)
    A2.sort_indices()
    v_feas, c_feas = port.cvt_to_features(c2, bl2, A2, bu2, l2, u2)
    _compare(g, x_s, x_t, scaled, {"c": c2, "b_l": bl2, "b_u": bu2, "l": l2, "u": u2, "A": A2, "v_feas": v_feas,
                                   "c_feas": c_feas})


def test_device_features_feed_the_model(cuda):
    """raw LP -> device features -> GCN_FC -> statuses equals the host-feature route."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch, features, ops, synth
    from lpgnn_b200.data import Data
    from lpgnn_b200.graph import BipartiteCSR
    m, n = 3000, 6000
    c, b_l, A, b_u, l, u = synth.raw_lp(m, n, 5 * n, 21)
    g, x_s, x_t, _ = features.prepare_lp_device(c, b_l, A, b_u, l, u, cuda)
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=64, depth=3).to(cuda).eval()
    with torch.no_grad():
        lc, lv = model(Data(x_s=x_s, x_t=x_t, edge_index=g))
        st_dev = ops.basis_select(lc, lv, k_basic=m, int64=False).cpu().numpy()
        c2, bl2, A2, bu2, l2, u2 = features.scale_lp(c, b_l, A, b_u, l, u)
        v_feas, c_feas = features.node_features(c2, bl2, A2, bu2, l2, u2)
        A2 = sp.csr_matrix(A2)
        A2.sort_indices()
        row = np.repeat(np.arange(m), np.diff(A2.indptr))
        g2 = BipartiteCSR.from_coo_arrays(row, A2.indices, A2.data.astype(np.float32), m, n, cuda, is_sorted=True)
        hs = torch.from_numpy(c_feas.astype(np.float32)).to(cuda)
        ht = torch.from_numpy(v_feas.astype(np.float32)).to(cuda)
        lc2, lv2 = model(Data(x_s=hs, x_t=ht, edge_index=g2))
        st_host = ops.basis_select(lc2, lv2, k_basic=m, int64=False).cpu().numpy()
    assert (st_dev == st_host).mean() >= 0.999
    assert float((lc - lc2).abs().max()) < 1e-3 and float((lv - lv2).abs().max()) < 1e-3
