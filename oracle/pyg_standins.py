"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the third-party arithmetic the reference's hot path delegates to.
None of these packages is vendored under /root/reference and none is installed in this
image; the reference pins them only in prose (readme.md:51-52: ``pytorch-sparse==0.6.12``,
``torch_geometric==2.1``).  Their published algorithms are restated here and anchored on
the reference's call sites:

* ``SparseTensor``      <- torch_sparse 0.6.12 ``SparseTensor`` / ``SparseStorage``
                           (call sites dataset.py:301-304, arch.py:71, dataset.py:133-144,
                           utils.py:909-915)
* ``spmm_sum``          <- torch_sparse ``spmm_sum`` CPU kernel: row-parallel, sequential fp32
                           accumulation over the row's nnz in CSR order, backward wrt the
                           dense operand = SpMM with the CSC view (arch.py:75-80 via GraphConv)
* ``GraphConv``         <- torch_geometric 2.1 ``nn.GraphConv`` with ``aggr='add'``
                           (call sites arch.py:57-62, 75-80)
* ``to_undirected``     <- torch_geometric 2.1 ``utils.to_undirected`` (dataset.py:252)
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn


# --------------------------------------------------------------------------------------
# torch_sparse.SparseTensor stand-in
# --------------------------------------------------------------------------------------
class _Storage:
    def __init__(self, owner):
        self._o = owner

    def value(self):
        return self._o._val

    def row(self):
        return self._o._row

    def col(self):
        return self._o._col

    def rowptr(self):
        return self._o._rowptr


class SparseTensor:
    """COO(sorted by row*ncols+col) + rowptr, value optional.  Mirrors the subset of the
    torch_sparse API the reference touches."""

    def __init__(self, row, col, value, sparse_sizes, is_sorted=False):
        m, n = int(sparse_sizes[0]), int(sparse_sizes[1])
        row = row.to(torch.long)
        col = col.to(torch.long)
        if not is_sorted and row.numel() > 1:
            # SparseStorage.__init__: sort only when the linear index is not already
            # non-decreasing (torch_sparse/storage.py, "if not is_sorted").
            idx = row * n + col
            if bool((idx[1:] < idx[:-1]).any()):
                perm = torch.argsort(idx, stable=True)
                row, col = row[perm], col[perm]
                if value is not None:
                    value = value[perm]
        self._row, self._col, self._val = row, col, value
        self._m, self._n = m, n
        counts = torch.bincount(row, minlength=m) if row.numel() else torch.zeros(m, dtype=torch.long)
        self._rowptr = torch.zeros(m + 1, dtype=torch.long)
        self._rowptr[1:] = torch.cumsum(counts, 0)
        self._t_cache = None
        self.storage = _Storage(self)

    # -- constructors ------------------------------------------------------------------
    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None, is_sorted=False):
        return cls(edge_index[0], edge_index[1], edge_attr, sparse_sizes, is_sorted=is_sorted)

    # -- structure ---------------------------------------------------------------------
    def sparse_sizes(self):
        return (self._m, self._n)

    def size(self, dim):
        return (self._m, self._n)[dim]

    def nnz(self):
        return int(self._row.numel())

    def density(self):
        return self.nnz() / (self._m * self._n)

    def t(self):
        """CSC view: permutation ``csr2csc = argsort(col*nrows + row)`` (SparseStorage.csr2csc)."""
        if self._t_cache is None:
            key = self._col * self._m + self._row
            perm = torch.argsort(key, stable=True)
            val = None if self._val is None else self._val[perm]
            self._t_cache = SparseTensor(self._col[perm], self._row[perm], val,
                                         (self._n, self._m), is_sorted=True)
            self._t_cache._csr2csc = perm
        return self._t_cache

    def clone(self):
        return SparseTensor(self._row.clone(), self._col.clone(),
                            None if self._val is None else self._val.clone(),
                            (self._m, self._n), is_sorted=True)

    def set_value(self, value, layout=None):
        return SparseTensor(self._row, self._col, value, (self._m, self._n), is_sorted=True)

    def sum(self, dim):
        val = self._val if self._val is not None else torch.ones(self.nnz())
        if dim == 0:
            return torch.zeros(self._n, dtype=val.dtype).index_add_(0, self._col, val)
        return torch.zeros(self._m, dtype=val.dtype).index_add_(0, self._row, val)

    def half(self):
        return self.set_value(None if self._val is None else self._val.half())

    def float(self):
        return self.set_value(None if self._val is None else self._val.float())

    def to(self, *args, **kwargs):
        return self  # CPU-only oracle

    def cpu(self):
        return self

    def to_dense(self, dtype=torch.float64):
        d = torch.zeros(self._m, self._n, dtype=dtype)
        d.index_put_((self._row, self._col), self._val.to(dtype), accumulate=True)
        return d

    def to_scipy_csr(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self._val.numpy(), self._col.numpy(), self._rowptr.numpy()),
                             shape=(self._m, self._n))

    def to_torch_csr(self, dtype):
        return torch.sparse_csr_tensor(self._rowptr, self._col, self._val.to(dtype),
                                       size=(self._m, self._n))


class _SpmmSum(torch.autograd.Function):
    """torch_sparse ``spmm_sum``: Y[i] = sum_e val[e] * X[col[e]] (row-parallel, sequential
    within a row); backward wrt X = transposed SpMM (CSC view).  No gradient for value (A is data)."""

    @staticmethod
    def forward(ctx, adj: SparseTensor, x: torch.Tensor):
        ctx.adj = adj
        return torch.sparse.mm(adj.to_torch_csr(x.dtype), x)

    @staticmethod
    def backward(ctx, gy):
        adj_t = ctx.adj.t()
        return None, torch.sparse.mm(adj_t.to_torch_csr(gy.dtype), gy.contiguous())


def spmm_sum(adj: SparseTensor, x: torch.Tensor) -> torch.Tensor:
    return _SpmmSum.apply(adj, x)


def spmm_sum_sequential(adj: SparseTensor, x: torch.Tensor) -> torch.Tensor:
    """Same op through scipy's csr_matvecs (strictly sequential fp32 axpy per nnz, CSR order):
    the bit-level model of the reference CPU kernel.  No autograd."""
    a = adj.to_scipy_csr().astype(x.numpy().dtype)
    return torch.from_numpy(np.asarray(a @ x.numpy()))


# --------------------------------------------------------------------------------------
# torch_geometric.nn.GraphConv stand-in (aggr='add')
# --------------------------------------------------------------------------------------
class _PygLinear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear: weight [out,in]; default init
    kaiming_uniform(a=sqrt(5)) (weight_initializer=None), bias uniform(+-1/sqrt(in))."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1.0 / math.sqrt(in_channels) if in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight, self.bias)


class GraphConv(nn.Module):
    """out = lin_rel(sum_j e_ji * x_src[j]) + lin_root(x_dst); ``adj_t`` has shape (N_dst, N_src)."""

    def __init__(self, in_channels, out_channels, aggr='add', bias=True, **kwargs):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        assert aggr == 'add'
        self.lin_rel = _PygLinear(in_channels[0], out_channels, bias=bias)
        self.lin_root = _PygLinear(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index, edge_weight=None, size=None):
        if isinstance(x, torch.Tensor):
            x = (x, x)
        out = spmm_sum(edge_index, x[0])
        out = self.lin_rel(out)
        if x[1] is not None:
            out = out + self.lin_root(x[1])
        return out


class _Unused(nn.Module):  # LayerNorm / GENConv are only touched by out-of-scope archs
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError


# --------------------------------------------------------------------------------------
# torch_geometric.utils.to_undirected
# --------------------------------------------------------------------------------------
def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce='add'):
    """cat((row,col),(col,row)), duplicate attr, coalesce = sort by row*N+col and sum
    duplicates (none occur for a bipartite graph)."""
    row, col = edge_index[0], edge_index[1]
    if num_nodes is None:
        num_nodes = int(max(row.max(), col.max())) + 1 if row.numel() else 0
    r = torch.cat([row, col])
    c = torch.cat([col, row])
    a = None if edge_attr is None else torch.cat([edge_attr, edge_attr])
    key = r * num_nodes + c
    perm = torch.argsort(key, stable=True)
    key, r, c = key[perm], r[perm], c[perm]
    if a is not None:
        a = a[perm]
    keep = torch.ones_like(key, dtype=torch.bool)
    keep[1:] = key[1:] != key[:-1]
    if not bool(keep.all()):
        seg = torch.cumsum(keep.long(), 0) - 1
        if a is not None:
            a = torch.zeros(int(seg[-1]) + 1, dtype=a.dtype).index_add_(0, seg, a)
        r, c = r[keep], c[keep]
    ei = torch.stack([r, c])
    return ei if edge_attr is None else (ei, a)
