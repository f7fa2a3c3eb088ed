"""(f-4) Sampled-subgraph path for LPs above ``edge_num_thresh``, on the device.

The reference feeds such LPs to ``torch_geometric.loader.NeighborLoader`` on the unipartite graph and converts every
sampled subgraph with ``MyToBipartite`` (train.py:103-116: ``num_neighbors=[6]*depth, shuffle=True, drop_last=True,
directed=False``; val.py:14-36: ``num_neighbors=[-1]*depth, shuffle=False``).  Here the whole LP stays resident in HBM
once (``ResidentLP``: the same CSR + CSC the full-graph path uses) and every mini-batch is cut out of it by three
kernels (``lpgnn_sample_mark``, ``lpgnn_induced_count``, ``lpgnn_induced_fill``, ``csrc/sample.cu``):

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


class NeighborSubgraphLoader:
    """Iterates seed batches of a ``ResidentLP`` and yields bipartite mini-batches (see the module docstring)."""

    def __init__(self, lp: ResidentLP, num_neighbors, batch_size, shuffle=False, drop_last=False, seed=0):
        self.lp, self.num_neighbors = lp, [int(f) for f in num_neighbors]
        self.batch_size = int(min(batch_size, lp.num_nodes))
        self.shuffle, self.drop_last, self.seed, self.epoch = shuffle, drop_last, int(seed), 0

    def __len__(self):
        full, rem = divmod(self.lp.num_nodes, self.batch_size)
        return full if (self.drop_last or rem == 0) else full + 1

    def __iter__(self):
        lp = self.lp
        if self.shuffle:
            gen = torch.Generator(device="cpu").manual_seed(self.seed + 7919 * self.epoch)
            order = torch.randperm(lp.num_nodes, generator=gen).to(lp.device)
        else:
            order = torch.arange(lp.num_nodes, device=lp.device)
        self.epoch += 1
        for b in range(len(self)):
            yield self.sample(order[b * self.batch_size:(b + 1) * self.batch_size], salt=self.epoch * 1_000_003 + b)

    def sample(self, seeds, salt=0) -> Data:
        lp, dev = self.lp, self.lp.device
        g = lp.graph if not lp.graph._transposed else lp.graph.t()
        csr, csc = g.views()
        seeds = seeds.to(dev)
        cons_seeds = seeds[seeds < lp.m].to(torch.int32)
        var_seeds = (seeds[seeds >= lp.m] - lp.m).to(torch.int32)
        in_c = torch.zeros(lp.m, dtype=torch.uint8, device=dev)
        in_v = torch.zeros(lp.n, dtype=torch.uint8, device=dev)
        in_c[cons_seeds.long()] = 1
        in_v[var_seeds.long()] = 1
        cons_parts, var_parts = [cons_seeds], [var_seeds]
        front_c, front_v = cons_seeds, var_seeds
        with torch.cuda.device(dev):
            for hop, fan in enumerate(self.num_neighbors):
                mark_v, mark_c = torch.zeros_like(in_v), torch.zeros_like(in_c)
                s = (self.seed * 0x9E3779B1 + salt * 0x85EBCA77 + hop * 0xC2B2AE3D) & (2 ** 63 - 1)
                _sample_mark(csr, front_c.contiguous(), fan, s, mark_v)
                _sample_mark(csc, front_v.contiguous(), fan, s ^ 0x5555555555555555, mark_c)
                front_v = torch.nonzero(mark_v & (1 - in_v)).flatten().to(torch.int32)     # new variables, ascending
                front_c = torch.nonzero(mark_c & (1 - in_c)).flatten().to(torch.int32)
                in_v |= mark_v
                in_c |= mark_c
                cons_parts.append(front_c)
                var_parts.append(front_v)
        cons_nodes = torch.cat(cons_parts).contiguous()
        var_nodes = torch.cat(var_parts).contiguous()
        sub = induced_subgraph(lp, cons_nodes, var_nodes)
        ci, vi = cons_nodes.long(), var_nodes.long()
        batch = Data(x_s=lp.x_s[ci], x_t=lp.x_t[vi], edge_index=sub, n_id_s=cons_nodes, n_id_t=var_nodes)
        if lp.y_s is not None:
            batch.y_s, batch.y_t = lp.y_s[ci], lp.y_t[vi]
        batch.bs = batch.batch_size = int(seeds.numel())
        batch.s_bs, batch.t_bs = int(cons_seeds.numel()), int(var_seeds.numel())
        return batch


def conv_depth(arch_str: str) -> int:
    """Hops to sample = conv layers of the arch string (train.py:108-110: ``depth - 1`` because of the FC head)."""
    import re
    d = re.findall(r"depth=(\d+)", arch_str or "")
    return 2 if not d else int(d[0]) - 1
