"""``BipartiteCSR``: the LP coefficient matrix A (m constraints x n variables) in HBM.

Drop-in for the ``torch_sparse.SparseTensor`` that the reference stores in ``batch.edge_index``
(built at reference dataset.py:301-304, transposed at arch.py:71, queried at dataset.py:133-144,
moved at train.py:118 / utils.py:909-915).  It keeps BOTH orientations resident --

    CSR  rowptr[m+1], col[z], val[z]          (row-major, ascending column inside a row)
    CSC  colptr[n+1], row_csc[z], val_csc[z]  (column-major, ascending row inside a column)
    csr2csc[z]                                (val_csc = val[csr2csc])

-- as int32 / float32, because the forward pass of one direction and the backward pass of the other
use the same view.  Construction happens on the device (``lpgnn_graph_build``); a graph created in
a DataLoader worker stays a host COO until ``.to(cuda)`` is called in the main process, so CUDA is
never initialised in forked workers ("CUDA in DataLoader workers").
"""
from __future__ import annotations

import torch

from . import _lib


_STATUS_POOL = {}


def _status_slot(dev, pool_words=4096):
    """Hands out zero-initialised int32 status words from a per-device pool so that the hot loop does not pay a
    memset launch per graph; the pool is replaced (one ``torch.zeros``) every ``pool_words`` graphs.  One pool per
    (device, stream), so a word is always zeroed on the stream that uses it."""
    key = (dev.type, dev.index, torch.cuda.current_stream(dev).cuda_stream)   # per stream: the zero fill is stream-ordered
    pool = _STATUS_POOL.get(key)
    if pool is None or pool[1] >= pool_words:
        pool = [torch.zeros(pool_words, dtype=torch.int32, device=dev), 0]
        _STATUS_POOL[key] = pool
    word = pool[0][pool[1]:pool[1] + 1]
    pool[1] += 1
    return word


class _StorageView:
    """``edge_index.storage.value()`` etc. (reference arch.py:21)."""

    def __init__(self, g):
        self._g = g

    def value(self):
        return self._g.values()

    def row(self):
        return self._g.coo()[0]

    def col(self):
        return self._g.coo()[1]

    def rowptr(self):
        g = self._g
        g._require_built()
        return (g.colptr if g._transposed else g.rowptr).long()


class BipartiteCSR:
    def __init__(self):
        self.m = self.n = 0
        self._coo = None            # (row, col, val) host/device tensors before the build
        self.rowptr = self.col = self.val = None
        self.colptr = self.row_csc = self.val_csc = self.csr2csc = None
        self._transposed = False    # True: this object presents A^T (shares buffers with its parent)
        self._sorted_hint = False   # caller's is_sorted=True: skip the COO sort (verified on the device)
        self.normalize = None       # 'mean': values divided by the destination degree at build time (inference only)
        self._status = None         # device int32: bit0 = sorted claim false, bit1 = index out of range

    @property
    def storage(self):
        """``edge_index.storage`` (created on demand: a stored back-reference would make every graph a reference
        cycle, which keeps its HBM buffers alive until the cyclic GC runs and forces fresh cudaMallocs)."""
        return _StorageView(self)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None, is_sorted=False, normalize=None):
        """Same signature as ``SparseTensor.from_edge_index`` (reference dataset.py:301-304).
        ``edge_index`` [2,z] integer (any order), ``edge_attr`` [z] float (default: ones)."""
        if sparse_sizes is None:
            raise ValueError("sparse_sizes=(m, n) is required")
        return cls.from_coo(edge_index[0], edge_index[1], edge_attr, int(sparse_sizes[0]), int(sparse_sizes[1]),
                            is_sorted=is_sorted, normalize=normalize)

    @classmethod
    def from_coo(cls, row, col, val, m, n, is_sorted=False, normalize=None):
        """COO given as separate tensors (host or device, any integer dtype).  Device inputs are built
        immediately on the current stream; host inputs are kept until ``.to(cuda)``.  ``normalize='mean'`` bakes the
        degree normalisation into the values (mean aggregation; the reference runs with sum aggregation, so this is
        OFF by default and supported for the forward pass only: SURVEY Appendix D)."""
        if normalize not in (None, "mean"):
            raise ValueError("normalize must be None or 'mean'")
        g = cls()
        g.normalize = normalize
        g.m, g.n = int(m), int(n)
        if val is None:
            val = torch.ones(row.shape[0], dtype=torch.float32, device=row.device)
        # range check only for host inputs (a device-side check would force a stream sync in the hot loop;
        # the build kernel reports out-of-range indices in its status word, see check())
        if not row.is_cuda and row.numel() and (
                int(row.max()) >= g.m or int(col.max()) >= g.n or int(row.min()) < 0 or int(col.min()) < 0):
            raise ValueError("edge_index out of range for sparse_sizes")
        g._coo = (row.to(torch.int32).contiguous(), col.to(torch.int32).contiguous(),
                  val.to(torch.float32).contiguous())
        g._sorted_hint = bool(is_sorted)
        if row.is_cuda:
            g._build()
        return g

    @classmethod
    def from_coo_arrays(cls, row, col, val, m, n, device, is_sorted=False, normalize=None):
        """numpy / tensor COO -> built graph on ``device`` (one H2D copy per array)."""
        g = cls.from_coo(torch.as_tensor(row), torch.as_tensor(col), torch.as_tensor(val), m, n, is_sorted=is_sorted,
                         normalize=normalize)
        return g.to(device)

    def _build(self):
        row, col, val = self._coo
        dev = row.device
        z = int(row.shape[0])
        lib = _lib.load()
        # one allocation for all seven arrays (16-byte aligned segments); views are carved out of it
        r4 = lambda k: (k + 3) // 4 * 4
        sizes = [r4(self.m + 1), r4(self.n + 1), r4(z), r4(z), r4(z), r4(z), r4(z)]
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + sz)
        buf = torch.empty(max(offs[-1], 4), dtype=torch.int32, device=dev)
        seg = lambda i, k: buf[offs[i]:offs[i] + k]
        self._buf = buf
        self.rowptr, self.colptr = seg(0, self.m + 1), seg(1, self.n + 1)
        self.col, self.row_csc, self.csr2csc = seg(2, z), seg(3, z), seg(4, z)
        self.val, self.val_csc = seg(5, z).view(torch.float32), seg(6, z).view(torch.float32)
        self._status = _status_slot(dev)             # pre-zeroed word (the kernels only OR bits in)
        ws_bytes = lib.lpgnn_graph_build_workspace_bytes(z, self.m, self.n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.lpgnn_graph_build(row.data_ptr(), col.data_ptr(), 0, val.data_ptr(), z, self.m, self.n,
                                       (_lib.COO_SORTED if self._sorted_hint else 0) | (_lib.GRAPH_MEAN if self.normalize == "mean" else 0),
                                       self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr(),
                                       self.colptr.data_ptr(), self.row_csc.data_ptr(), self.val_csc.data_ptr(),
                                       self.csr2csc.data_ptr(), self._status.data_ptr(), ws.data_ptr(), ws_bytes,
                                       _lib.stream_ptr())
        _lib.check(rc, "lpgnn_graph_build")
        # `ws` and the COO are allocated and consumed on the same (current) stream, so the caching allocator's
        # stream-ordered reuse is already safe; no record_stream (it defers reuse and forces fresh cudaMallocs).
        self._coo = None

    def check(self):
        """Synchronising validation of the device-side build status: raises if an ``is_sorted=True`` claim
        was false or an index was out of range.  (The hot loop never calls this; tests and loaders do.)"""
        self._require_built()
        st = int(self._status.item())
        if st & 1:
            raise ValueError("BipartiteCSR: is_sorted=True was passed but the COO is not in (row, col) order")
        if st & 2:
            raise ValueError("BipartiteCSR: edge index out of range for sparse_sizes")
        return self

    def _require_built(self):
        if self.rowptr is None:
            raise RuntimeError("BipartiteCSR is still a host COO: call .to('cuda') first "
                               "(graphs are built on the device; there is no CPU path)")

    # ------------------------------------------------------------------ movement
    @property
    def is_cuda(self):
        return self.rowptr is not None

    @property
    def device(self):
        return self.rowptr.device if self.rowptr is not None else self._coo[0].device

    def pin_memory(self):
        if self._coo is not None and not self._coo[0].is_cuda:
            self._coo = tuple(t.pin_memory() for t in self._coo)
        return self

    def to(self, device=None, *args, non_blocking=False, **kwargs):
        if device is None or isinstance(device, torch.dtype):
            return self
        device = torch.device(device)
        if device.type != "cuda":
            if self.rowptr is not None:
                raise RuntimeError("a built BipartiteCSR lives on the GPU; use .coo() to read it back")
            return self
        if self.rowptr is not None:
            if self.rowptr.device != device and device.index is not None:
                raise RuntimeError("moving a built BipartiteCSR between GPUs is not supported")
            return self
        self._coo = tuple(t.to(device, non_blocking=non_blocking) for t in self._coo)
        self._build()
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def half(self):       # reference utils.py:913 (`batch[nm].half()`): values stay fp32, features carry the dtype
        return self

    def float(self):
        return self

    # ------------------------------------------------------------------ SparseTensor API subset
    def t(self):
        """A^T as a view sharing the same buffers (reference arch.py:71)."""
        v = BipartiteCSR.__new__(BipartiteCSR)
        v.__dict__.update(self.__dict__)
        v._transposed = not self._transposed
        return v

    def sparse_sizes(self):
        return (self.n, self.m) if self._transposed else (self.m, self.n)

    def size(self, dim):
        return self.sparse_sizes()[dim]

    def nnz(self):
        return int(self.col.shape[0]) if self.col is not None else int(self._coo[0].shape[0])

    def density(self):
        return self.nnz() / float(self.m * self.n)

    def views(self):
        """(ptr, idx, val, n_dst) of the presented matrix and of its transpose."""
        self._require_built()
        csr = (self.rowptr, self.col, self.val, self.m)
        csc = (self.colptr, self.row_csc, self.val_csc, self.n)
        return (csc, csr) if self._transposed else (csr, csc)

    def values(self):
        self._require_built()
        return self.val_csc if self._transposed else self.val

    def coo(self):
        """(row, col, value) int64/int64/float32 in the presented orientation's canonical order."""
        (ptr, idx, val, rows), _ = self.views()
        counts = (ptr[1:] - ptr[:-1]).long()
        row = torch.repeat_interleave(torch.arange(rows, device=ptr.device), counts)
        return row, idx.long(), val

    def clone(self):
        g = BipartiteCSR.__new__(BipartiteCSR)
        g.__dict__.update(self.__dict__)
        for k in ("rowptr", "col", "val", "colptr", "row_csc", "val_csc", "csr2csc"):
            if getattr(self, k) is not None:
                setattr(g, k, getattr(self, k).clone())
        if self._coo is not None:
            g._coo = tuple(t.clone() for t in self._coo)
        return g

    def set_value(self, value, layout="coo"):
        """New graph with the same structure and ``value`` given in the presented orientation's
        canonical (coo/csr) order (reference dataset.py:136)."""
        self._require_built()
        g = self.clone()
        value = value.to(device=self.val.device, dtype=torch.float32)
        perm = self.csr2csc.long()
        if self._transposed:
            g.val_csc = value.contiguous()
            g.val = torch.empty_like(value)
            g.val[perm] = value
        else:
            g.val = value.contiguous()
            g.val_csc = value[perm].contiguous()
        return g

    def sum(self, dim):
        """Column sums (dim=0) or row sums (dim=1) of the presented matrix (dataset.py:137-138)."""
        (ptr, idx, val, rows), (ptr_t, _, val_t, rows_t) = self.views()
        src_ptr, src_val, n_out = (ptr, val, rows) if dim == 1 else (ptr_t, val_t, rows_t)
        csum = torch.zeros(src_val.shape[0] + 1, dtype=torch.float64, device=src_val.device)
        csum[1:] = torch.cumsum(src_val.double(), 0)
        return (csum[src_ptr[1:].long()] - csum[src_ptr[:-1].long()]).float()

    def __repr__(self):
        r, c = self.sparse_sizes()
        where = self.device if (self.rowptr is not None or self._coo is not None) else "?"
        return f"BipartiteCSR({r}x{c}, nnz={self.nnz()}, device={where}, built={self.rowptr is not None})"
