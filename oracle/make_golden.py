"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Generates tests/golden/*.npz by running the reference's OWN Python (imported verbatim from
/root/reference through oracle/ref_import.py) on seeded inputs.  Run in the build container:

    python -m oracle.make_golden

The reference ships no golden vectors for this path, so these files -- outputs of the
reference itself -- are what pins the oracle port and the CUDA path.  Each file stores the inputs,
the weights (so nothing depends on RNG reproducibility) and the reference outputs:

  lp_features_*.npz   dataset.scaling + dataset.cvt_to_features          (dataset.py:23-96)
  graph_*.npz         LPDataset.get tensor part + MyToBipartite + .t()    (dataset.py:250-304, arch.py:71)
  model_*.npz         GCN_FC.forward logits, val.inference_gnn statuses, train.balanced loss and
                      parameter gradients                                  (arch.py:167-193, val.py:106-124,
                                                                            train.py:39-46)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def raw_small_lp(m, n, z, seed, finite_bounds=True):
    """A raw LP with every kind of bound (finite / +-inf / zero / 1e308 'infinite' markers)."""
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, m, size=z)
    cols = rng.integers(0, n, size=z)
    key = np.unique(rows * n + cols)
    rows, cols = key // n, key % n
    vals = rng.uniform(-10, 10, size=key.shape[0])
    unit = rng.random(key.shape[0]) < 0.5
    vals[unit] = rng.choice([-1.0, 1.0], size=int(unit.sum()))
    A = sp.csr_matrix((vals, (rows, cols)), shape=(m, n))
    rhs = rng.normal(0, 5, size=m)
    kind = rng.random(m)
    b_l = np.where(kind < 0.4, -np.inf, rhs)
    b_u = np.where((kind >= 0.4) & (kind < 0.8), np.inf, rhs + (kind >= 0.9) * 2.0)
    b_l[0] = -np.inf
    if m > 3:
        b_u[1] = 1.5e308          # "infinite" marker above 1e308 -> inf (dataset.py:24-27)
        b_l[2] = 0.0
        b_u[2] = 0.0
    l = np.zeros(n)
    u = np.where(rng.random(n) < 0.6, np.inf, rng.uniform(1, 10, size=n))
    if finite_bounds:
        l[rng.random(n) < 0.2] = -np.inf
        l[rng.random(n) < 0.1] = -3.0
    c = rng.normal(0, 1, size=n)
    return c, b_l, A, b_u, l, u


def gen_features(R, name, m, n, z, seed, finite_bounds):
    c, b_l, A, b_u, l, u = raw_small_lp(m, n, z, seed, finite_bounds)
    inp = dict(c=c.copy(), b_l=b_l.copy(), b_u=b_u.copy(), l=l.copy(), u=u.copy(),
               A_data=A.data.copy(), A_indices=A.indices.copy(), A_indptr=A.indptr.copy(), shape=np.array([m, n]))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        c2, bl2, A2, bu2, l2, u2 = R.dataset.scaling(c.copy(), b_l.copy(), A.copy(), b_u.copy(), l.copy(), u.copy())
    v_feas, c_feas = R.dataset.cvt_to_features(c2, bl2, A2, bu2, l2, u2)
    A2 = A2.tocsr()
    A2.sort_indices()
    np.savez_compressed(os.path.join(OUT, f"lp_features_{name}.npz"), **{f"in_{k}": v for k, v in inp.items()},
                        out_c=c2, out_b_l=bl2, out_b_u=bu2, out_l=l2, out_u=u2, out_A_data=A2.data,
                        out_A_indices=A2.indices, out_A_indptr=A2.indptr, out_v_feas=v_feas, out_c_feas=c_feas)
    return c2, bl2, A2, bu2, l2, u2, v_feas, c_feas


def reference_batch_via_stubs(R, A, c_feas, v_feas, y_s, y_t):
    """LPDataset.get (tensor part, dataset.py:241-262) + MyToBipartite.__call__ (275-332, verbatim)."""
    from oracle.pyg_standins import to_undirected
    coo = A.tocoo()
    row, col, A_data = coo.row, coo.col, coo.data
    ncons = c_feas.shape[0]
    nnodes = ncons + v_feas.shape[0]
    cf, vf = torch.from_numpy(c_feas), torch.from_numpy(v_feas)
    edge_attr = torch.from_numpy(A_data).float()                                    # dataset.py:250
    edge_index = torch.from_numpy(np.asarray([row, col + ncons], dtype=np.int64))    # dataset.py:251
    edge_index, edge_attr = to_undirected(edge_index, edge_attr)                     # dataset.py:252
    is_vars = torch.zeros(nnodes, dtype=torch.long)
    is_vars[ncons:] = 1
    data = R.dataset.UnipartiteData(x=torch.cat((cf, vf), dim=0), y=torch.cat((torch.from_numpy(y_s), torch.from_numpy(y_t))),
                                    is_vars=is_vars, edge_index=edge_index, edge_attr=edge_attr, num_nodes=nnodes,
                                    processed_path="synthetic")
    uni = dict(edge_index=edge_index.numpy().copy(), edge_attr=edge_attr.numpy().copy(), is_vars=is_vars.numpy().copy())
    data.batch = torch.zeros(nnodes, dtype=torch.long)
    batch = R.dataset.MyToBipartite(thresh_num=np.inf)(data)                         # verbatim transform
    return uni, batch


def labels(c_feas, v_feas, seed):
    rng = np.random.default_rng(seed)

    def draw(f):
        y = rng.integers(0, 3, size=f.shape[0])
        y = np.where((y == 0) & (f[:, 5] != 0), 1, y)
        y = np.where((y == 2) & (f[:, 7] != 0), 1, y)
        return y.astype(np.int64)
    return draw(c_feas), draw(v_feas)


def gen_model(R, name, m, n, z, seed, hids, depth, finite_bounds=False):
    c2, bl2, A2, bu2, l2, u2, v_feas, c_feas = gen_features(R, name, m, n, z, seed, finite_bounds)
    # float32 cast as torch.FloatTensor does at dataset.py:196
    v32, c32 = v_feas.astype(np.float32), c_feas.astype(np.float32)
    y_s, y_t = labels(c32, v32, seed + 1)
    uni, batch = reference_batch_via_stubs(R, A2, c32, v32, y_s, y_t)
    st = batch.edge_index                      # SparseTensor stand-in (canonical order == torch_sparse)
    st_t = st.t()
    np.savez_compressed(
        os.path.join(OUT, f"graph_{name}.npz"),
        uni_edge_index=uni["edge_index"], uni_edge_attr=uni["edge_attr"], is_vars=uni["is_vars"],
        m=np.array(m), n=np.array(n),
        rowptr=st.storage.rowptr().numpy(), col=st.storage.col().numpy(), val=st.storage.value().numpy(),
        colptr=st_t.storage.rowptr().numpy(), row_csc=st_t.storage.col().numpy(), val_csc=st_t.storage.value().numpy(),
        csr2csc=st_t._csr2csc.numpy(),
        x_s=batch.x_s.numpy(), x_t=batch.x_t.numpy(), y_s=batch.y_s.numpy(), y_t=batch.y_t.numpy(),
        s_bs=np.array(batch.s_bs), t_bs=np.array(batch.t_bs))

    torch.manual_seed(0)
    model = R.arch.GCN_FC(8, 8, hids=hids, depth=depth)
    sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        lc, lv = model(batch)
        pred = R.val.inference_gnn(torch.cat((lc, lv), dim=0), m)
    # training step quantities (dp only matters in train mode; use eval-mode forward for determinism)
    model.zero_grad()
    lc_g, lv_g = model(batch)
    bal = R.utils.labels_to_balanced_weights
    crit_s = torch.nn.CrossEntropyLoss(weight=bal(batch.y_s))
    crit_t = torch.nn.CrossEntropyLoss(weight=bal(batch.y_t))
    loss = (m + n) / m * crit_s(lc_g, batch.y_s) + (m + n) / n * crit_t(lv_g, batch.y_t)   # train.py:39-46
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}
    np.savez_compressed(
        os.path.join(OUT, f"model_{name}.npz"), hids=np.array(hids), depth=np.array(depth),
        logits_cons=lc.numpy(), logits_vars=lv.numpy(), pred=pred.numpy(), loss=np.array(loss.item()),
        **{f"w::{k}": v for k, v in sd.items()}, **{f"g::{k}": v for k, v in grads.items()})
    print(f"{name}: m={m} n={n} nnz={A2.nnz} hids={hids} depth={depth} loss={loss.item():.6f} "
          f"basic={(pred == 1).sum().item()}")


def main():
    os.makedirs(OUT, exist_ok=True)
    R = load_reference()
    gen_model(R, "tiny_5x7", 5, 7, 14, seed=11, hids=64, depth=3, finite_bounds=True)
    gen_model(R, "small_300x600", 300, 600, 3000, seed=12, hids=128, depth=3, finite_bounds=True)
    gen_model(R, "c1_1000x2000", 1000, 2000, 10_000, seed=1235, hids=64, depth=2, finite_bounds=False)
    gen_features(R, "bounds_40x90", 40, 90, 500, seed=13, finite_bounds=True)


if __name__ == "__main__":
    main()
