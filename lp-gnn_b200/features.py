"""Host-side LP preprocessing: scaling and the 8-per-node feature layout.

Mirrors ``dataset.scaling`` (reference dataset.py:23-76) and ``dataset.cvt_to_features``
(dataset.py:79-96, helpers utils.py:323-383).  This runs once per LP, offline, in float64 on
the host (out of scope for kernels); it exists so that the
synthetic-LP generator and ``LPDataset.process`` emit exactly the reference's feature/tag
layout that the mask kernel reads (columns 5 and 7 are the +-inf tags).

Variable features  x_t[j] = [c_j, nnz(A[:,j])/m, cos(b_l,A[:,j]), cos(b_u,A[:,j]), l_j|0, tag(l_j), u_j|0, tag(u_j)]
Constraint feats   x_s[i] = [cos(A[i,:],c), nnz(A[i,:])/n, cos(A[i,:],l), cos(A[i,:],u), b_l,i|0, tag, b_u,i|0, tag]
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

_BIG = 1e308
_CLIP = 1e8
_TINY = 1e-6


def _safe_abs(v: np.ndarray) -> np.ndarray:
    a = np.abs(v)
    return np.where(np.isinf(a) | (a == 0.0), 1.0, a)


def scale_lp(c, b_l, A, b_u, l, u):
    """Returns scaled copies ``(c, b_l, A_csr, b_u, l, u)`` with |A| <= 1 and |c| <= 1."""
    c = np.array(c, dtype=np.float64)
    b_l = np.array(b_l, dtype=np.float64)
    b_u = np.array(b_u, dtype=np.float64)
    l = np.array(l, dtype=np.float64)
    u = np.array(u, dtype=np.float64)
    b_u[b_u > _BIG] = np.inf
    b_l[b_l < -_BIG] = -np.inf
    u[u > _BIG] = np.inf
    l[l < -_BIG] = -np.inf

    A = sp.csr_matrix(A, dtype=np.float64, copy=True)
    A.sort_indices()
    # row scale: the larger finite, non-zero |bound| of the row
    row_scale = np.maximum(_safe_abs(b_l), _safe_abs(b_u))
    A.data /= row_scale[np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))]
    b_l /= row_scale
    b_u /= row_scale

    # column scale: max(|A| column max, 1/|l|, 1/|u|)
    col_scale = np.zeros(A.shape[1], dtype=np.float64)
    np.maximum.at(col_scale, A.indices, np.abs(A.data))
    col_scale[np.isinf(col_scale) | (col_scale == 0.0)] = 1.0
    col_scale = np.maximum(col_scale, np.maximum(1.0 / _safe_abs(l), 1.0 / _safe_abs(u)))
    A.data /= col_scale[A.indices]
    l *= col_scale
    u *= col_scale
    c = c / col_scale

    c_max = float(np.abs(c).max()) if c.size else 0.0
    c /= (c_max if c_max != 0.0 else 1.0)
    return c, b_l, A, b_u, l, u


def _cosine_with_columns(v: np.ndarray, A_csc: sp.csc_matrix) -> np.ndarray:
    """cos(v, A[:,j]) for all j.  The dot product runs over each column's entries in ascending
    row order (CSC), which is the order scipy uses for ``v * A`` in the reference."""
    v = np.clip(v, -_CLIP, _CLIP)
    norm_v = np.sqrt((v ** 2).sum())
    sq = np.asarray(A_csc.multiply(A_csc).sum(axis=0)).ravel()
    norm_cols = np.sqrt(sq)
    dot = np.asarray(A_csc.T @ v).ravel()
    norm_cols[norm_cols == 0] = _TINY
    if norm_v == 0:
        norm_v = _TINY
    return dot / (norm_v * norm_cols)


def _value_tag(v: np.ndarray) -> np.ndarray:
    tag = np.where(v == np.inf, 1.0, np.where(v == -np.inf, -1.0, 0.0))
    val = np.where(np.isinf(v), 0.0, v)
    return np.stack([val, tag], axis=1)


def node_features(c, b_l, A, b_u, l, u):
    """Returns ``(v_feas[n,8], c_feas[m,8])`` float64 (the caller casts to float32,
    as ``torch.FloatTensor`` does at dataset.py:196)."""
    A = sp.csr_matrix(A)
    m, n = A.shape
    A_csc = A.tocsc()
    At_csc = A.T.tocsc()           # columns of A^T = rows of A
    nz_r, nz_c = A.nonzero()
    deg_col = np.bincount(nz_c, minlength=n) / m
    deg_row = np.bincount(nz_r, minlength=m) / n
    v_feas = np.column_stack([
        c, deg_col, _cosine_with_columns(b_l, A_csc), _cosine_with_columns(b_u, A_csc),
        _value_tag(l), _value_tag(u)])
    c_feas = np.column_stack([
        _cosine_with_columns(c, At_csc), deg_row,
        _cosine_with_columns(l, At_csc), _cosine_with_columns(u, At_csc),
        _value_tag(b_l), _value_tag(b_u)])
    return v_feas, c_feas


def prepare_lp_device(c, b_l, A, b_u, l, u, device):
    """Raw LP on the host -> model inputs on ``device`` with scaling and features computed THERE
    (``lpgnn_lp_features``; same arithmetic as :func:`scale_lp` + :func:`node_features`, i.e. reference
    dataset.py:23-96).  Returns ``(graph, x_s, x_t, scaled)``: a built ``BipartiteCSR`` with the scaled
    coefficients, the feature rows, and the float64 scaled LP.  This is the "predict on a fresh LP" entry:
    no offline ``process()`` pass."""
    import torch
    from . import ops
    from .graph import BipartiteCSR
    A = sp.csr_matrix(A, dtype=np.float64, copy=True)
    A.sum_duplicates()
    A.sort_indices()
    m, n = A.shape
    row = np.repeat(np.arange(m, dtype=np.int32), np.diff(A.indptr))
    dev = torch.device(device)
    g = BipartiteCSR.from_coo(torch.from_numpy(row), torch.from_numpy(A.indices.astype(np.int32)),
                              torch.zeros(A.nnz, dtype=torch.float32), m, n, is_sorted=True).to(dev)
    up = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
    x_s, x_t, scaled = ops.lp_features(g, up(A.data), up(c), up(b_l), up(b_u), up(l), up(u))
    return g, x_s, x_t, scaled
