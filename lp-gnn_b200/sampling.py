"""(f-4) Sampled-subgraph path for LPs above ``edge_num_thresh``, on the device.

The reference feeds such LPs to ``torch_geometric.loader.NeighborLoader`` on the unipartite graph and converts every
sampled subgraph with ``MyToBipartite`` (train.py:103-116: ``num_neighbors=[6]*depth, shuffle=True, drop_last=True,
directed=False``; val.py:14-36: ``num_neighbors=[-1]*depth, shuffle=False``).  Here the whole LP stays resident in HBM
once (``ResidentLP``: the same CSR + CSC the full-graph path uses) and every mini-batch is cut out of it on the device
(``lpgnn_sample_nodes``, ``lpgnn_induced_offsets``, ``lpgnn_induced_fill_sorted``, ``csrc/sample.cu``) with one host
read of the resulting sizes:

* seeds   = ``batch_size`` consecutive entries of the node order (a permutation when shuffling) over the unipartite
            ids -- constraints ``0..m-1`` first, variables ``m..m+n-1`` (dataset.py:258-260);
* hops    = for every frontier node ``min(deg, fanout)`` neighbours without replacement (all when ``fanout < 0``);
            a hop from constraints walks CSR rows, a hop from variables CSC rows;
* edges   = the subgraph INDUCED by the sampled nodes (``directed=False``), relabelled so that the seeds come first
            on each side (``logits[:s_bs]`` / ``[:t_bs]`` are the seeds, train.py:122-123), then the nodes found at
            hop 1, hop 2, ... in ascending id order.

The batches carry the same fields as ``MyToBipartite`` output (``x_s, x_t, y_s, y_t, edge_index, bs, s_bs, t_bs``) plus
``n_id_s`` / ``n_id_t`` (global ids of the local nodes).  With full neighbourhoods (``fanout = -1``) and as many hops
as the model has conv layers, the seed logits equal the full-graph logits -- the equivalence the reference notes at
val.py:44-47 and ``tests/test_gpu_sampling.py`` asserts.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import _lib
from .data import Data
from .graph import BipartiteCSR


class ResidentLP:
    """One LP resident on the device: built graph (both orientations) + features + labels."""

    def __init__(self, graph, x_s, x_t, y_s=None, y_t=None):
        graph._require_built()
        self.graph, self.x_s, self.x_t, self.y_s, self.y_t = graph, x_s, x_t, y_s, y_t
        self.m, self.n = graph.m, graph.n
        self.device = x_s.device

    @classmethod
    def from_unipartite(cls, data, device):
        """``data``: the unipartite graph ``LPDataset.get`` returns (left untouched)."""
        from .dataset import MyToBipartite
        d = MyToBipartite(thresh_num=np.inf)(copy.copy(data))
        dev = torch.device(device)
        g = d.edge_index.to(dev)
        f = lambda t: None if t is None else t.to(dev)
        return cls(g, f(d.x_s).float(), f(d.x_t).float(), f(getattr(d, "y_s", None)), f(getattr(d, "y_t", None)))

    @property
    def num_nodes(self):
        return self.m + self.n


def _sample_mark(view, frontier, fanout, seed, marks):
    ptr, idx, _, _ = view
    if frontier.numel() == 0:
        return
    rc = _lib.load().lpgnn_sample_mark(ptr.data_ptr(), idx.data_ptr(), frontier.data_ptr(), frontier.numel(), int(fanout),
                                       int(seed) & (2 ** 64 - 1), marks.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "lpgnn_sample_mark")


def induced_subgraph(lp: ResidentLP, cons_nodes, var_nodes) -> BipartiteCSR:
    """Bipartite graph induced by the given constraint / variable ids (int32 device tensors, local order)."""
    dev = lp.device
    lib = _lib.load()
    (ptr, idx, val, _), _ = lp.graph.views() if not lp.graph._transposed else lp.graph.t().views()
    mc, nv = int(cons_nodes.numel()), int(var_nodes.numel())
    map_v = torch.full((lp.n,), -1, dtype=torch.int32, device=dev)
    map_v[var_nodes.long()] = torch.arange(nv, dtype=torch.int32, device=dev)
    counts = torch.zeros(max(mc, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.lpgnn_induced_count(ptr.data_ptr(), idx.data_ptr(), cons_nodes.data_ptr(), mc, map_v.data_ptr(),
                                           counts.data_ptr(), _lib.stream_ptr()), "lpgnn_induced_count")
        csum = torch.cumsum(counts.long(), 0)
        offsets = (csum - counts.long()).contiguous()
        z = int(csum[mc - 1].item()) if mc else 0            # one host sync per mini-batch: sizes the edge arrays
        row = torch.empty(z, dtype=torch.int32, device=dev)
        col = torch.empty(z, dtype=torch.int32, device=dev)
        v = torch.empty(z, dtype=torch.float32, device=dev)
        if z:
            _lib.check(lib.lpgnn_induced_fill(ptr.data_ptr(), idx.data_ptr(), val.data_ptr(), cons_nodes.data_ptr(), mc,
                                              map_v.data_ptr(), offsets.data_ptr(), row.data_ptr(), col.data_ptr(),
                                              v.data_ptr(), _lib.stream_ptr()), "lpgnn_induced_fill")
        # local column ids are not monotone in the global ones (seeds first) -> canonical order from the device sort
        return BipartiteCSR.from_coo(row, col, v, mc, nv, is_sorted=False)


class _SamplerBuffers:
    """Device buffers of the one-sync sampler, sized by the resident LP and reused by every mini-batch."""

    def __init__(self, lp: ResidentLP, max_seeds: int):
        lib = _lib.load()
        dev = lp.device
        i32 = lambda k: torch.empty(max(int(k), 1), dtype=torch.int32, device=dev)
        self.cons_nodes, self.var_nodes, self.map_c, self.map_v = i32(lp.m), i32(lp.n), i32(lp.m), i32(lp.n)
        self.offsets = i32(lp.m + 1)
        self.sizes_len = int(lib.lpgnn_sample_sizes_len())
        self.sizes = i32(self.sizes_len)
        self.sizes_host = torch.empty(self.sizes_len, dtype=torch.int32).pin_memory()
        self.max_seeds = int(max_seeds)
        self.ws_bytes = int(lib.lpgnn_sample_nodes_workspace_bytes(lp.m, lp.n, max_seeds))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)


class NeighborSubgraphLoader:
    """Iterates seed batches of a ``ResidentLP`` and yields bipartite mini-batches (see the module docstring).

    Every mini-batch is cut out by ``lpgnn_sample_nodes`` (seed split + all hops + ordered node lists, on the device),
    ``lpgnn_induced_offsets``, ONE host read of the sizes (nodes per side, seeds per side, nnz), ``lpgnn_induced_fill_sorted``
    and ``lpgnn_graph_build`` on its sorted path -- no ``nonzero`` / ``.item()`` per hop."""

    def __init__(self, lp: ResidentLP, num_neighbors, batch_size, shuffle=False, drop_last=False, seed=0):
        self.lp, self.num_neighbors = lp, [int(f) for f in num_neighbors]
        self.batch_size = int(min(batch_size, lp.num_nodes))
        self.shuffle, self.drop_last, self.seed, self.epoch = shuffle, drop_last, int(seed), 0
        self._buf = None

    def __len__(self):
        full, rem = divmod(self.lp.num_nodes, self.batch_size)
        return full if (self.drop_last or rem == 0) else full + 1

    def __iter__(self):
        """Mini-batch k+1's node sampling is ENQUEUED BEFORE mini-batch k is handed out: on the (in-order) stream it runs
        ahead of the consumer's training step k, so when the consumer comes back for batch k+1 its sizes are already on
        the host and the one host read never waits for an idle GPU."""
        lp = self.lp
        if self.shuffle:
            gen = torch.Generator(device="cpu").manual_seed(self.seed + 7919 * self.epoch)
            order = torch.randperm(lp.num_nodes, generator=gen).to(lp.device)
        else:
            order = torch.arange(lp.num_nodes, device=lp.device)
        self.epoch += 1
        nb = len(self)
        seeds_of = lambda b: order[b * self.batch_size:(b + 1) * self.batch_size]
        salt_of = lambda b: self.epoch * 1_000_003 + b
        pending = self._stage_a(seeds_of(0), salt_of(0)) if nb else None
        for b in range(nb):
            batch = self._stage_b(pending)
            pending = self._stage_a(seeds_of(b + 1), salt_of(b + 1)) if b + 1 < nb else None
            yield batch

    def sample(self, seeds, salt=0) -> Data:
        return self._stage_b(self._stage_a(seeds, salt))

    def _stage_a(self, seeds, salt):
        """Node sets + row offsets of the induced subgraph on the device, sizes on their way to the host."""
        import ctypes as C
        lp, dev = self.lp, self.lp.device
        lib = _lib.load()
        g = lp.graph if not lp.graph._transposed else lp.graph.t()
        (rowptr, col, _, _), (colptr, row_csc, _, _) = g.views()
        seeds = seeds.to(dev, torch.int64).contiguous()
        ns = int(seeds.numel())
        if self._buf is None or ns > self._buf.max_seeds:
            self._buf = _SamplerBuffers(lp, max(ns, self.batch_size))
        B = self._buf
        hops = len(self.num_neighbors)
        fan = (C.c_int32 * max(hops, 1))(*self.num_neighbors) if hops else None
        s = (self.seed * 0x9E3779B1 + salt * 0x85EBCA77) & (2 ** 64 - 1)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            _lib.check(lib.lpgnn_sample_nodes(rowptr.data_ptr(), col.data_ptr(), colptr.data_ptr(), row_csc.data_ptr(), lp.m, lp.n,
                                              seeds.data_ptr(), ns, fan, hops, s, B.cons_nodes.data_ptr(), B.var_nodes.data_ptr(),
                                              B.map_c.data_ptr(), B.map_v.data_ptr(), B.sizes.data_ptr(), B.ws.data_ptr(),
                                              B.ws_bytes, st), "lpgnn_sample_nodes")
            _lib.check(lib.lpgnn_induced_offsets(rowptr.data_ptr(), col.data_ptr(), B.cons_nodes.data_ptr(), lp.m,
                                                 B.map_v.data_ptr(), B.offsets.data_ptr(), B.sizes.data_ptr(), B.ws.data_ptr(),
                                                 B.ws_bytes, st), "lpgnn_induced_offsets")
            B.sizes_host.copy_(B.sizes, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        return ns, ev, seeds                  # (seeds kept alive until the kernels that read them have been enqueued)

    def _stage_b(self, pending) -> Data:
        """The one host read (sizes), then the sorted induced COO, the graph build and the feature gathers."""
        ns, ev, _ = pending
        lp, dev, B = self.lp, self.lp.device, self._buf
        lib = _lib.load()
        g = lp.graph if not lp.graph._transposed else lp.graph.t()
        (rowptr, col, val, _), _ = g.views()
        ev.synchronize()
        sz = B.sizes_host.tolist()
        mc, nv, s_bs, t_bs, z = sz[0], sz[1], sz[2], sz[3], sz[B.sizes_len - 2]
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            row = torch.empty(z, dtype=torch.int32, device=dev)
            colo = torch.empty(z, dtype=torch.int32, device=dev)
            v = torch.empty(z, dtype=torch.float32, device=dev)
            if z:
                _lib.check(lib.lpgnn_induced_fill_sorted(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), B.cons_nodes.data_ptr(),
                                                         mc, B.map_v.data_ptr(), B.offsets.data_ptr(), row.data_ptr(),
                                                         colo.data_ptr(), v.data_ptr(), st), "lpgnn_induced_fill_sorted")
            sub = BipartiteCSR.from_coo(row, colo, v, mc, nv, is_sorted=True)
            # features, labels and node ids of the sampled nodes: one launch (the node lists live in reused buffers)
            native = (lp.x_s.dtype == torch.float32 and lp.x_t.dtype == torch.float32 and lp.x_s.is_contiguous()
                      and lp.x_t.is_contiguous() and (lp.y_s is None or (lp.y_s.dtype == torch.int64 and lp.y_t.dtype == torch.int64
                                                                          and lp.y_s.is_contiguous() and lp.y_t.is_contiguous())))
            if native:
                p_, q_ = lp.x_s.shape[1], lp.x_t.shape[1]
                x_s = torch.empty((mc, p_), dtype=torch.float32, device=dev)
                x_t = torch.empty((nv, q_), dtype=torch.float32, device=dev)
                cons_nodes = torch.empty(mc, dtype=torch.int32, device=dev)
                var_nodes = torch.empty(nv, dtype=torch.int32, device=dev)
                y_s = torch.empty(mc, dtype=torch.int64, device=dev) if lp.y_s is not None else None
                y_t = torch.empty(nv, dtype=torch.int64, device=dev) if lp.y_s is not None else None
                _lib.check(lib.lpgnn_sample_gather(lp.x_s.data_ptr(), lp.x_t.data_ptr(), _lib.ptr(lp.y_s), _lib.ptr(lp.y_t),
                                                   B.cons_nodes.data_ptr(), mc, B.var_nodes.data_ptr(), nv, p_, q_, x_s.data_ptr(),
                                                   x_t.data_ptr(), _lib.ptr(y_s), _lib.ptr(y_t), cons_nodes.data_ptr(),
                                                   var_nodes.data_ptr(), st), "lpgnn_sample_gather")
            else:
                cons_nodes, var_nodes = B.cons_nodes[:mc].clone(), B.var_nodes[:nv].clone()
                ci, vi = cons_nodes.long(), var_nodes.long()
                x_s, x_t = lp.x_s[ci], lp.x_t[vi]
                y_s, y_t = (lp.y_s[ci], lp.y_t[vi]) if lp.y_s is not None else (None, None)
        batch = Data(x_s=x_s, x_t=x_t, edge_index=sub, n_id_s=cons_nodes, n_id_t=var_nodes)
        if lp.y_s is not None:
            batch.y_s, batch.y_t = y_s, y_t
        batch.bs = batch.batch_size = ns
        batch.s_bs, batch.t_bs = s_bs, t_bs
        return batch


def conv_depth(arch_str: str) -> int:
    """Hops to sample = conv layers of the arch string (train.py:108-110: ``depth - 1`` because of the FC head)."""
    import re
    d = re.findall(r"depth=(\d+)", arch_str or "")
    return 2 if not d else int(d[0]) - 1
