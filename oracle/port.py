"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Standalone CPU restatement (numpy / scipy / torch-CPU) of the lp-gnn hot path.  Unlike
``ref_import.py`` it does not need /root/reference, so it travels to the GPU box, where it is
the checker for the ``-m gpu`` parity tests and the timed ``cpu_baseline`` ("port") of bench.py.

Every function cites the reference file:line it follows.  It is validated in the build
container against the verbatim reference (tests/test_oracle_vs_reference.py) and everywhere
against the committed golden vectors (tests/golden/*.npz, produced by oracle/make_golden.py
from the verbatim reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
import torch
from torch import nn
import torch.nn.functional as F


# ======================================================================================
# LP -> features   (dataset.py:23-96, utils.py:323-383)   [float64 on the host, once per LP]
# ======================================================================================
def _unit_where_degenerate(v):
    """|v| with inf and 0 replaced by 1 (dataset.py:30-33, 50-53, 57-58)."""
    s = np.abs(v)
    s[np.isinf(s) | (s == 0)] = 1.0
    return s


def scaling(c, b_l, A, b_u, l, u):
    """Row / column / objective scaling so that |A|<=1 and |c|<=1 (dataset.py:23-76).
    Returns new arrays; inputs are not modified (the reference mutates in place)."""
    c, b_l, b_u, l, u = (np.array(v, dtype=np.float64, copy=True) for v in (c, b_l, b_u, l, u))
    b_u[b_u > 1e308] = np.inf
    b_l[b_l < -1e308] = -np.inf
    u[u > 1e308] = np.inf
    l[l < -1e308] = -np.inf

    # rows: divide by max(|b_l|, |b_u|) (degenerate -> 1)      dataset.py:29-37
    s_row = np.maximum(_unit_where_degenerate(b_l), _unit_where_degenerate(b_u))
    A = sp.csr_matrix(A, dtype=np.float64, copy=True)
    A.data /= np.repeat(s_row, np.diff(A.indptr))             # utils.py:323-328
    b_l /= s_row
    b_u /= s_row

    # columns: divide by max(colmax|A|, 1/|l|, 1/|u|)           dataset.py:50-66
    s_col2 = np.maximum(1.0 / _unit_where_degenerate(l), 1.0 / _unit_where_degenerate(u))
    s_col = np.asarray(abs(A).max(axis=0).todense()).ravel().astype(np.float64)
    s_col[np.isinf(s_col) | (s_col == 0)] = 1.0
    s_col = np.maximum(s_col, s_col2)
    Ac = A.tocsc(copy=True)
    Ac.data /= np.repeat(s_col, np.diff(Ac.indptr))           # utils.py:329-332
    A = Ac.tocsr()
    l *= s_col
    u *= s_col
    c = c / s_col

    s_c = np.abs(c).max() if c.size else 0.0                  # dataset.py:68-74
    if s_c == 0.0:
        s_c = 1.0
    c /= s_c
    return c, b_l, A, b_u, l, u


def _cos_vec_cols(v, A, bound=1e8):
    """cos(v, A[:,j]) for every column j (utils.py:349-360)."""
    v = np.clip(v, -bound, bound)
    n_v = math.sqrt(float((v ** 2).sum()))
    n_cols = np.sqrt(np.asarray(A.multiply(A).sum(axis=0)).ravel())
    dot = v * A                                               # same scipy expression -> same bits
    n_cols = n_cols.copy()
    n_cols[n_cols == 0] = 1e-6
    if n_v == 0:
        n_v = 1e-6
    return dot / (n_v * n_cols)


def _value_and_inf_tag(v):
    """(value with +-inf -> 0, tag in {+1,-1,0}) (utils.py:368-374)."""
    tag = np.zeros_like(v)
    tag[v == np.inf] = 1
    tag[v == -np.inf] = -1
    val = v.copy()
    val[np.isinf(val)] = 0
    return np.stack([val, tag], axis=1)


def _nnz_count(A, axis):
    """count_nonzero_sparse_mat (utils.py:335-347): counts stored entries that are non-zero."""
    r, c = A.nonzero()
    if axis == "col":
        return np.bincount(c, minlength=A.shape[1]).astype(np.float64)
    return np.bincount(r, minlength=A.shape[0]).astype(np.float64)


def cvt_to_features(c, b_l, A, b_u, l, u):
    """8 features per variable and per constraint (dataset.py:79-96; layout SURVEY Appendix A)."""
    m, n = A.shape
    At = A.T
    v_feas = np.concatenate([
        c.reshape(-1, 1),
        (_nnz_count(A, "col") / m).reshape(-1, 1),
        _cos_vec_cols(b_l, A).reshape(-1, 1),
        _cos_vec_cols(b_u, A).reshape(-1, 1),
        _value_and_inf_tag(l), _value_and_inf_tag(u)], axis=1)
    c_feas = np.concatenate([
        _cos_vec_cols(c, At).reshape(-1, 1),
        (_nnz_count(A, "row") / n).reshape(-1, 1),
        _cos_vec_cols(l, At).reshape(-1, 1),
        _cos_vec_cols(u, At).reshape(-1, 1),
        _value_and_inf_tag(b_l), _value_and_inf_tag(b_u)], axis=1)
    return v_feas, c_feas


# ======================================================================================
# graph construction   (dataset.py:229-264 get, 275-332 MyToBipartite, arch.py:71 .t())
# ======================================================================================
@dataclass
class BipartiteGraph:
    m: int
    n: int
    rowptr: np.ndarray   # int64 [m+1]
    col: np.ndarray      # int64 [z]   ascending within a row
    val: np.ndarray      # float32 [z]
    colptr: np.ndarray   # int64 [n+1]
    row_csc: np.ndarray  # int64 [z]   ascending within a column
    val_csc: np.ndarray  # float32 [z]
    csr2csc: np.ndarray  # int64 [z]   val_csc = val[csr2csc]

    @property
    def nnz(self):
        return int(self.col.shape[0])

    def scipy_csr(self, dtype=np.float32):
        return sp.csr_matrix((self.val.astype(dtype), self.col, self.rowptr), shape=(self.m, self.n))

    def scipy_csc_as_csr_of_transpose(self, dtype=np.float32):
        return sp.csr_matrix((self.val_csc.astype(dtype), self.row_csc, self.colptr),
                             shape=(self.n, self.m))


def unipartite_edges(row, col, a_data, ncons):
    """LPDataset.get (dataset.py:250-252): edge_index=[row, col+ncons], to_undirected ->
    sorted by (src*N+dst), attrs duplicated.  Returns (edge_index[2,2z] int64, edge_attr[2z] f32)."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64) + ncons
    attr = np.asarray(a_data).astype(np.float32)
    src = np.concatenate([row, col])
    dst = np.concatenate([col, row])
    att = np.concatenate([attr, attr])
    n_nodes = int(max(src.max(), dst.max())) + 1 if src.size else 0
    order = np.argsort(src * n_nodes + dst, kind="stable")
    return np.stack([src[order], dst[order]]), att[order]


def graph_from_coo(row, col, val, m, n, normalize=None) -> BipartiteGraph:
    """SparseTensor.from_edge_index (dataset.py:301-304) + .t() (arch.py:71): canonical CSR =
    stable sort by row*n+col; CSC view = stable sort of the CSR entries by col*m+row.
    ``normalize='mean'``: the aggregation PyG calls aggr='mean' (left commented out at arch.py:57,60): each
    orientation's values divided by the degree of its destination node (float32 division)."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=np.float32)
    order = np.argsort(row * n + col, kind="stable")
    row, col, val = row[order], col[order], val[order]
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=m), out=rowptr[1:])
    csr2csc = np.argsort(col * m + row, kind="stable")
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(col, minlength=n), out=colptr[1:])
    val_csc = val[csr2csc]
    if normalize == "mean":
        deg_r = np.maximum(np.diff(rowptr), 1).astype(np.float32)
        deg_c = np.maximum(np.diff(colptr), 1).astype(np.float32)
        val = (val / deg_r[row]).astype(np.float32)
        val_csc = (val_csc / deg_c[col[csr2csc]]).astype(np.float32)
    elif normalize is not None:
        raise ValueError(normalize)
    return BipartiteGraph(m, n, rowptr, col, val, colptr, row[csr2csc], val_csc, csr2csc)


def to_bipartite(edge_index, edge_attr, is_vars):
    """MyToBipartite.__call__ (dataset.py:283-304): relabel so constraints are 0..m-1 and
    variables m..m+n-1, keep the cons->var half of the undirected edges, build the m x n matrix."""
    is_vars = np.asarray(is_vars).astype(bool)
    n = int(is_vars.sum())
    m = int(is_vars.shape[0]) - n
    mapping = np.empty(m + n, dtype=np.int64)
    mapping[~is_vars] = np.arange(m)
    mapping[is_vars] = np.arange(n) + m
    src = mapping[np.asarray(edge_index[0])]
    dst = mapping[np.asarray(edge_index[1])]
    keep = src < m
    assert int(keep.sum()) * 2 == keep.shape[0]               # dataset.py:297
    return graph_from_coo(src[keep], dst[keep] - m, np.asarray(edge_attr)[keep], m, n)


# ======================================================================================
# model   (arch.py:51-81 GraphConvTwoDirection, 129-141 add_knowledge, 167-193 GCN_FC)
# ======================================================================================
def spmm_sequential(ptr, idx, val, x):
    """Y[i] = sum_e val[e]*X[idx[e]], e ascending, accumulated in x's dtype -- the reference
    CPU kernel's order (torch_sparse spmm_sum).  scipy's csr_matvecs is exactly this loop."""
    rows = ptr.shape[0] - 1
    a = sp.csr_matrix((val.astype(x.dtype), idx, ptr), shape=(rows, x.shape[0]))
    return np.asarray(a @ x)


def add_knowledge_np(left, right, x_s, x_t, bound=10.0):
    """arch.py:129-141 in numpy (float32): row L2-normalise (eps 1e-12) x10, then -bound on
    class 0 where feature[-3] != 0 and on class 2 where feature[-1] != 0."""
    out = []
    for logit, feas in ((left, x_s), (right, x_t)):
        logit = logit.astype(np.float32)
        nrm = np.sqrt((logit * logit).sum(axis=1, keepdims=True, dtype=np.float32))
        y = logit / np.maximum(nrm, np.float32(1e-12)) * np.float32(10)
        lo = feas[:, -3] != 0
        up = feas[:, -1] != 0
        y[lo, 0] -= np.float32(bound)
        y[up, 2] -= np.float32(bound)
        out.append(y)
    return out[0], out[1]


def add_knowledge_t(left, right, x_s, x_t, bound=10):
    """arch.py:129-141 with torch ops (autograd-capable, out-of-place)."""
    def one(logit, feas):
        y = F.normalize(logit) * 10
        off = torch.zeros_like(y)
        off[:, 0] = feas[:, -3].abs().bool().to(y.dtype) * bound
        off[:, 2] = feas[:, -1].abs().bool().to(y.dtype) * bound
        return y - off
    return one(left, x_s), one(right, x_t)


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd_csr, bwd_csr):
        ctx.bwd = bwd_csr
        return torch.sparse.mm(fwd_csr, x)

    @staticmethod
    def backward(ctx, g):
        return torch.sparse.mm(ctx.bwd, g.contiguous()), None, None


class TorchGraph:
    """A (m x n) and A^T as torch CSR tensors (A^T built from the CSC view: same entry order
    as the reference's .t())."""

    def __init__(self, g: BipartiteGraph, dtype=torch.float32):
        self.g = g
        t = torch.from_numpy
        self.A = torch.sparse_csr_tensor(t(g.rowptr), t(g.col), t(g.val).to(dtype), size=(g.m, g.n))
        self.At = torch.sparse_csr_tensor(t(g.colptr), t(g.row_csc), t(g.val_csc).to(dtype),
                                          size=(g.n, g.m))


class _Lin(nn.Module):
    """torch_geometric Linear: weight [out,in], kaiming_uniform(a=sqrt 5), bias U(+-1/sqrt(in))."""

    def __init__(self, i, o, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        self.bias = nn.Parameter(torch.empty(o)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            nn.init.uniform_(self.bias, -1 / math.sqrt(i), 1 / math.sqrt(i))

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class _GraphConv(nn.Module):
    def __init__(self, in_src, in_dst, out):
        super().__init__()
        self.lin_rel = _Lin(in_src, out, True)
        self.lin_root = _Lin(in_dst, out, False)


class PortGraphConvTwoDirection(nn.Module):
    """arch.py:51-81: right' = lin_rel_l2r(A^T left) + lin_root_l2r(right);
    left' = lin_rel_r2l(A right) + lin_root_r2l(left); both read the OLD left/right."""

    def __init__(self, left_dim, right_dim, out_dim):
        super().__init__()
        self.left2right = _GraphConv(left_dim, right_dim, out_dim)
        self.right2left = _GraphConv(right_dim, left_dim, out_dim)

    def forward(self, left, right, tg: TorchGraph):
        agg_t = _SpmmFn.apply(left, tg.At, tg.A)
        agg_s = _SpmmFn.apply(right, tg.A, tg.At)
        right_new = self.left2right.lin_rel(agg_t) + self.left2right.lin_root(right)
        left_new = self.right2left.lin_rel(agg_s) + self.right2left.lin_root(left)
        return left_new, right_new


class PortGCN_FC(nn.Module):
    """arch.py:167-193.  Same constructor order (so ``torch.manual_seed`` gives the same
    initial weights as the reference) and the same state_dict keys."""

    def __init__(self, p, q, hids=128, depth=3, dp=.1):
        super().__init__()
        self.conv1 = PortGraphConvTwoDirection(p, q, hids)
        self.layers = nn.ModuleList(
            [PortGraphConvTwoDirection(hids, hids, hids) for _ in range(depth - 2)])
        self.lin_left = nn.Linear(hids, 3)
        self.lin_right = nn.Linear(hids, 3)
        self.dp = dp

    def forward(self, x_s, x_t, tg: TorchGraph):
        left, right = self.conv1(x_s, x_t, tg)
        left, right = left.relu(), right.relu()
        for conv in self.layers:
            left, right = conv(left, right, tg)
            left = F.dropout(left, p=self.dp, training=self.training)
            right = F.dropout(right, p=self.dp, training=self.training)
            left, right = left.relu(), right.relu()
        left, right = self.lin_left(left), self.lin_right(right)
        return add_knowledge_t(left, right, x_s, x_t)


def gcn_fc_forward_np(sd, x_s, x_t, g: BipartiteGraph, depth, acc_dtype=np.float32):
    """GCN_FC.forward in eval mode with numpy, strictly sequential SpMM accumulation
    (``acc_dtype=float64`` gives the high-precision model used to bound rounding error)."""
    f = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k]).astype(acc_dtype)
    left, right = x_s.astype(acc_dtype), x_t.astype(acc_dtype)

    def conv(prefix, left, right):
        agg_t = spmm_sequential(g.colptr, g.row_csc, g.val_csc, left)
        agg_s = spmm_sequential(g.rowptr, g.col, g.val, right)
        r = agg_t @ f(prefix + "left2right.lin_rel.weight").T + f(prefix + "left2right.lin_rel.bias") \
            + right @ f(prefix + "left2right.lin_root.weight").T
        l = agg_s @ f(prefix + "right2left.lin_rel.weight").T + f(prefix + "right2left.lin_rel.bias") \
            + left @ f(prefix + "right2left.lin_root.weight").T
        return l, r

    left, right = conv("conv1.", left, right)
    left, right = np.maximum(left, 0), np.maximum(right, 0)
    for i in range(depth - 2):
        left, right = conv(f"layers.{i}.", left, right)
        left, right = np.maximum(left, 0), np.maximum(right, 0)
    left = left @ f("lin_left.weight").T + f("lin_left.bias")
    right = right @ f("lin_right.weight").T + f("lin_right.bias")
    if acc_dtype == np.float64:
        out = []
        for logit, feas in ((left, x_s), (right, x_t)):
            nrm = np.sqrt((logit * logit).sum(axis=1, keepdims=True))
            y = logit / np.maximum(nrm, 1e-12) * 10
            y[feas[:, -3] != 0, 0] -= 10
            y[feas[:, -1] != 0, 2] -= 10
            out.append(y)
        return out[0], out[1]
    return add_knowledge_np(left, right, x_s, x_t)


# ======================================================================================
# basis decision   (val.py:106-124 inference_gnn)
# ======================================================================================
def inference_gnn_np(logits, m):
    """softmax over 3 classes (float32), NaN->0, the m largest P(basic) over all m+n nodes get
    status 1, every other node gets 0 if p0 >= p2 else 2.  Ties in the top-m are broken towards
    the LOWER node index (the reference's torch.topk leaves ties implementation-defined)."""
    x = np.asarray(logits, dtype=np.float32)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    pr = e / e.sum(axis=1, keepdims=True, dtype=np.float32)
    pr[np.isnan(pr)] = 0
    p1 = pr[:, 1]
    order = np.argsort(-p1, kind="stable")[:m]
    pred = np.where(pr[:, 0] >= pr[:, 2], 0, 2).astype(np.int64)
    pred[order] = 1
    return pred


def inference_gnn_t(logits, m):
    """val.py:106-124 with the same torch ops (ties as torch.topk on this build resolves them)."""
    pr = F.softmax(logits.float(), dim=-1).clone()
    pr[torch.isnan(pr)] = 0
    _, topk_idx = pr[:, 1].topk(m)
    pr[:, 1] = pr.min() - 1
    pr[topk_idx, 1] = pr.max() + 1
    return pr.argmax(-1)


# ======================================================================================
# loss   (train.py:39-46 balanced, utils.py:286-299 labels_to_balanced_weights)
# ======================================================================================
def labels_to_balanced_weights(labels):
    res = torch.zeros(3)
    lbl, cnt = torch.unique(labels, return_counts=True)
    res[lbl] = (cnt.sum() / cnt).float()
    if len(lbl) != 2:
        res[0] = res[2] = (res[0] + res[2]) / 2.
    return res


def balanced_loss(logit_cons, logit_vars, y_s, y_t):
    m, n = len(y_s), len(y_t)
    loss = (m + n) / m * F.cross_entropy(logit_cons, y_s, weight=labels_to_balanced_weights(y_s))
    loss = loss + (m + n) / n * F.cross_entropy(logit_vars, y_t, weight=labels_to_balanced_weights(y_t))
    return loss


# ---------------------------------------------------------------------------------------------
# (f-4) sampled-subgraph path: what NeighborLoader(directed=False) + MyToBipartite produce for a
# GIVEN node set (reference train.py:107-116, val.py:22-27, dataset.py:275-332).  The sampler's
# random choices are third-party (pyg-lib / torch_sparse neighbor_sample) and not reproducible, so
# the oracle restates the deterministic parts: full-neighbourhood expansion and the induced,
# relabelled bipartite subgraph.
# ---------------------------------------------------------------------------------------------
def khop_full_neighbourhood(A_csr, cons_seeds, var_seeds, hops):
    """Node lists (seeds first, then the nodes discovered at hop 1, 2, ... in ascending id order) reached from the
    seeds by `hops` full-neighbourhood expansions of the bipartite graph of A."""
    A = sp.csr_matrix(A_csr)
    At = A.T.tocsr()
    cons, vars_ = [np.asarray(cons_seeds, dtype=np.int64)], [np.asarray(var_seeds, dtype=np.int64)]
    in_c = np.zeros(A.shape[0], dtype=bool); in_c[cons[0]] = True
    in_v = np.zeros(A.shape[1], dtype=bool); in_v[vars_[0]] = True
    front_c, front_v = cons[0], vars_[0]
    for _ in range(hops):
        nv = np.unique(A[front_c].indices) if front_c.size else np.zeros(0, dtype=np.int64)
        nc = np.unique(At[front_v].indices) if front_v.size else np.zeros(0, dtype=np.int64)
        nv, nc = nv[~in_v[nv]], nc[~in_c[nc]]
        in_v[nv] = True; in_c[nc] = True
        cons.append(nc.astype(np.int64)); vars_.append(nv.astype(np.int64))
        front_c, front_v = nc, nv
    return np.concatenate(cons), np.concatenate(vars_)


def induced_bipartite_subgraph(A_csr, cons_nodes, var_nodes):
    """(rowptr, col, val) of A[cons_nodes][:, var_nodes] in canonical order (rows = local constraint ids, ascending local
    column ids inside a row): the graph MyToBipartite builds from the sampled, relabelled unipartite subgraph."""
    A = sp.csr_matrix(A_csr)
    sub = A[np.asarray(cons_nodes)][:, np.asarray(var_nodes)].tocsr()
    sub.sort_indices()
    return sub.indptr.astype(np.int64), sub.indices.astype(np.int64), sub.data.astype(np.float32)
