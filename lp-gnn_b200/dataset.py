"""Drop-in for the reference ``dataset.py``: same processed-file layout, same ``LPDataset.get`` ->
``UnipartiteData`` -> ``MyToBipartite`` flow and the same batch fields (``x_s, x_t, y_s, y_t, edge_index, bs,
s_bs, t_bs, processed_path[, con_nms, var_nms]``; reference dataset.py:229-332), but ``batch.edge_index`` is a
``BipartiteCSR`` (host COO until ``.to(cuda)``; built on the device) instead of a torch_sparse ``SparseTensor``.

Host code: runs in DataLoader worker processes and never touches CUDA.
"""
from __future__ import annotations

import logging
import os
import os.path as osp

import numpy as np
import torch
from scipy.sparse import coo_matrix

from .data import Data, Dataset
from .features import node_features, scale_lp
from .graph import BipartiteCSR
from .io_utils import msgpack_dump, msgpack_load


class UnipartiteData(Data):
    def __init__(self, x, y, edge_index, edge_weight=None, **kwargs):
        super().__init__(x=x, y=y, edge_index=edge_index, edge_weight=edge_weight, **kwargs)


scaling = scale_lp                 # reference names (dataset.py:23, 79)
cvt_to_features = node_features


def to_undirected_sorted(row, col_shifted, attr, num_nodes):
    """``torch_geometric.utils.to_undirected`` for a bipartite edge list (dataset.py:252): both directions,
    attributes duplicated, sorted by ``src * N + dst`` (no duplicates can occur)."""
    src = np.concatenate([row, col_shifted])
    dst = np.concatenate([col_shifted, row])
    att = np.concatenate([attr, attr])
    order = np.argsort(src * np.int64(num_nodes) + dst, kind="stable")
    return np.stack([src[order], dst[order]]), att[order]


class LPDataset(Dataset):
    def __init__(self, root, transform=None, pre_transform=None, pre_filter=None, load_meta=False, chunk=None):
        self.load_meta = load_meta
        self.chunk = chunk
        self.p = self.q = 8
        super().__init__(root, transform, pre_transform, pre_filter)

    def cache_size_info(self, recache=False):
        """dataset.py:119-157: one row per LP -- ``idx, nedges, nnodes, fn, ncons, nvars, density, num_basis_vars`` --
        cached in ``<root>/size.json``; returned as a DataFrame restricted to this (sub-)dataset's indices.  Needs the
        bipartite transform (it reads ``x_s / x_t / y_t`` and ``edge_index.nnz() / .density()``); host only."""
        import json

        import pandas as pd
        dump_fn = osp.join(self.root, "size.json")
        res = None
        if not recache and osp.exists(dump_fn):
            try:
                with open(dump_fn) as f:
                    res = json.load(f)
            except (OSError, ValueError) as e:
                logging.info(f"err {e}, recache")
        if res is None:
            res = []
            for idx in range(self.len()):                 # the whole dataset, also when called on a sub-dataset
                data = self.get(idx)
                data = data if self.transform is None else self.transform(data)
                res.append(dict(idx=idx, nedges=int(data.edge_index.nnz()), nnodes=int(data.num_nodes),
                                fn=osp.basename(data.processed_path), ncons=int(data.x_s.shape[0]),
                                nvars=int(data.x_t.shape[0]), density=float(data.edge_index.density()),
                                num_basis_vars=int((data.y_t == 1).sum())))
            with open(dump_fn, "w") as f:
                json.dump(res, f, sort_keys=True, indent=4)
        df = pd.DataFrame(res).loc[list(self.indices()), :]
        if len(self.indices()) != len(res):
            logging.warning("sub-dataset get info, please use full dataset ")
        df["fn"] = df.fn.str.replace(".pk", "", regex=False)
        return df

    def dump_size_info(self, dst):
        """dataset.py:107-117: the size table with a ``split`` column (seed-0 train / val split).  The reference
        writes a pandas HDF file (``tables``, not on this image); here ``dst`` receives the same table as JSON."""
        from .io_utils import split_train_val
        df = self.cache_size_info()
        if osp.exists(dst):
            return None
        train_ds, val_ds = split_train_val(self, seed=0)
        df.loc[list(train_ds.indices()), "split"] = "train"
        df.loc[list(val_ds.indices()), "split"] = "val"
        df.to_json(dst, orient="records", indent=1)
        return df

    @property
    def raw_file_names(self):
        fns = []
        for folder in [self.raw_dir, self.processed_dir]:
            if not osp.exists(folder):
                continue
            now = sorted((f for f in os.listdir(folder) if f.endswith(".pk")), key=lambda nm: (len(nm), nm))
            if len(now) > len(fns):
                fns = now
        if len(fns) == 0:
            raise ValueError("not found pk")
        return fns

    @property
    def processed_file_names(self):
        return self.raw_file_names

    def len(self):
        return len(self.processed_file_names)

    def process(self):
        """raw ``.pk`` -> processed ``.pk`` + ``.meta`` (dataset.py:178-224)."""
        os.makedirs(self.processed_dir, exist_ok=True)
        for raw_path in self.raw_paths:
            processed_path = osp.join(self.processed_dir, osp.basename(raw_path))
            if osp.exists(processed_path) and osp.exists(processed_path + ".meta"):
                continue
            [c, b_l, (row, col, data), b_u, l, u, con_lbls, var_lbls, con_nms, var_nms] = msgpack_load(raw_path)
            ncons, nvars = len(con_nms), len(var_nms)
            A = coo_matrix((np.asarray(data), (np.asarray(row), np.asarray(col))), shape=(ncons, nvars)).tocsr()
            c, b_l, A, b_u, l, u = scale_lp(np.asarray(c), np.asarray(b_l), A, np.asarray(b_u), np.asarray(l), np.asarray(u))
            v_feas, c_feas = node_features(c, b_l, A, b_u, l, u)
            v_feas, c_feas = v_feas.astype(np.float32), c_feas.astype(np.float32)
            y_s, y_t = np.asarray(con_lbls, dtype=np.int64), np.asarray(var_lbls, dtype=np.int64)
            # labels never contradict the +-inf tags (dataset.py:201-207)
            assert (y_s[c_feas[:, -3] != 0] != 0).all() and (y_s[c_feas[:, -1] != 0] != 2).all()
            violates = int((y_t[v_feas[:, -3] != 0] == 0).sum())
            if violates:
                logging.warning(f"violate {violates}")
            assert (y_t[v_feas[:, -1] != 0] != 2).all()
            A = A.tocoo()
            msgpack_dump([A.row, A.col, A.data, c_feas, v_feas, y_s, y_t, ncons + nvars], processed_path)
            msgpack_dump(dict(num_cons=ncons, num_vars=nvars, raw_path=raw_path, processed_path=processed_path,
                              con_nms=list(con_nms), var_nms=list(var_nms)), processed_path + ".meta")

    def get(self, idx):
        """dataset.py:229-264: unipartite graph with cons first, undirected sorted edges."""
        fn = osp.join(self.processed_dir, self.processed_file_names[idx])
        [row, col, A_data, c_feas, v_feas, y_s, y_t, nnodes] = msgpack_load(fn)
        row, col, A_data = np.asarray(row), np.asarray(col), np.asarray(A_data)
        c_feas, v_feas = np.asarray(c_feas), np.asarray(v_feas)
        ncons = c_feas.shape[0]
        assert A_data.max() <= 1 and A_data.min() >= -1
        assert c_feas.max() <= 1 and c_feas.min() >= -1
        aux = dict(processed_path=fn)
        if self.load_meta:
            meta = msgpack_load(fn + ".meta")
            aux.update(con_nms=meta["con_nms"], var_nms=meta["var_nms"])
        nnodes = int(nnodes)
        ei, ea = to_undirected_sorted(row.astype(np.int64), col.astype(np.int64) + ncons, A_data.astype(np.float32), nnodes)
        is_vars = torch.zeros(nnodes, dtype=torch.long)
        is_vars[ncons:] = 1
        return UnipartiteData(
            x=torch.cat((torch.from_numpy(c_feas), torch.from_numpy(v_feas)), dim=0),
            y=torch.cat((torch.from_numpy(np.asarray(y_s)), torch.from_numpy(np.asarray(y_t))), dim=0),
            is_vars=is_vars, edge_index=torch.from_numpy(ei), edge_attr=torch.from_numpy(ea), num_nodes=nnodes, **aux)


class MyToBipartite:
    """dataset.py:268-332.  Output graph: ``BipartiteCSR`` host COO (cons -> var half of the edges, already in
    canonical (row, col) order because the undirected list is sorted by ``src*N+dst``), built on the device when
    the batch is moved with ``batch.to(dev)`` / ``batch_to``."""

    def __init__(self, dev="cpu", phase="train", thresh_num=np.inf):
        self.dev, self.phase, self.thresh_num = dev, phase, thresh_num

    def __call__(self, batch):
        if hasattr(batch, "x_s"):
            return batch
        if batch.edge_index.shape[-1] // 2 > self.thresh_num:
            return batch
        is_vars = batch.is_vars.bool()
        is_cons = ~is_vars
        nnodes = batch.num_nodes
        nvars = int(is_vars.sum())
        ncons = nnodes - nvars
        mapping = torch.empty(nnodes, dtype=torch.long)
        mapping[is_cons] = torch.arange(ncons)
        mapping[is_vars] = torch.arange(nvars) + ncons
        src, dst = mapping[batch.edge_index[0]], mapping[batch.edge_index[1]]
        mask = src < ncons
        assert int(mask.sum()) * 2 == mask.shape[0]                               # dataset.py:297
        r, c = src[mask], dst[mask] - ncons
        # canonical (row, col) order is CLAIMED to the device build only when it has been verified here: one pass over the
        # composite keys (LPDataset.get always produces it; a hand-built graph with unsorted edges takes the sort path)
        key = r * max(nvars, 1) + c
        in_order = bool(is_cons[:ncons].all()) and (key.numel() < 2 or bool((key[1:] >= key[:-1]).all()))
        batch.edge_index = BipartiteCSR.from_coo(r, c, batch.edge_attr[mask], ncons, nvars, is_sorted=in_order)
        del batch.edge_attr
        batch.x_s, batch.x_t = batch.x[is_cons, :], batch.x[is_vars, :]
        batch.y_s, batch.y_t = batch.y[is_cons], batch.y[is_vars]
        del batch.x, batch.y
        batch.bs = batch.batch_size if hasattr(batch, "batch_size") else batch.num_nodes
        t_bs = int(batch.is_vars[:batch.bs].sum())
        batch.s_bs, batch.t_bs = batch.bs - t_bs, t_bs
        del batch.is_vars
        if self.phase == "train" and hasattr(batch, "batch"):
            del batch.batch
        return batch


def write_synthetic_dataset(root, sizes, seed=0, structure="staircase"):
    """Materialises processed ``.pk`` files of synthetic LPs (``synth.processed_lp``) so that the loaders,
    train.py and pred_basis.py can run end to end without the MIRP data.  ``sizes`` = [(m, n, nnz), ...]."""
    from . import synth
    pdir = osp.join(root, "processed")
    os.makedirs(pdir, exist_ok=True)
    for i, (m, n, z) in enumerate(sizes):
        lp = synth.processed_lp(m, n, z, seed=seed * 100_003 + i, structure=structure)
        fn = osp.join(pdir, f"lp{i}.pk")
        msgpack_dump([lp.row, lp.col, lp.a_data, lp.c_feas, lp.v_feas, lp.y_s, lp.y_t, lp.m + lp.n], fn)
        msgpack_dump(dict(num_cons=lp.m, num_vars=lp.n, raw_path="", processed_path=fn,
                          con_nms=[f"c{j}" for j in range(lp.m)], var_nms=[f"x{j}" for j in range(lp.n)]), fn + ".meta")
    return pdir


def pack_bipartite(batches):
    """Block-diagonal pack of bipartite LP batches (``MyToBipartite`` output, host side): ONE graph (the direct sum of the LPs'
    matrices, what a PyG ``DataLoader(batch_size > 1)`` would collate -- the reference runs ``batch_size=1``, train.py:70),
    concatenated features / labels and the segment pointers ``cons_ptr`` / ``vars_ptr`` (int32 [B+1]) that
    ``losses.balanced_packed`` and the segmented basis decision need.  Entries stay in canonical order when every LP's are."""
    from .data import Data
    ms = [int(b.x_s.shape[0]) for b in batches]
    ns = [int(b.x_t.shape[0]) for b in batches]
    c_ptr = np.concatenate([[0], np.cumsum(ms)]).astype(np.int64)
    v_ptr = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    rows, cols, vals, in_order = [], [], [], True
    for k, b in enumerate(batches):
        g = b.edge_index
        if g._coo is None:
            raise ValueError("pack_bipartite packs host-side graphs (before .to(device))")
        r, c, v = g._coo
        rows.append(r.to(torch.int64) + int(c_ptr[k]))
        cols.append(c.to(torch.int64) + int(v_ptr[k]))
        vals.append(v)
        in_order = in_order and bool(g._sorted_hint)
    M, N = int(c_ptr[-1]), int(v_ptr[-1])
    graph = BipartiteCSR.from_coo(torch.cat(rows), torch.cat(cols), torch.cat(vals), M, N, is_sorted=in_order)
    out = Data(x_s=torch.cat([b.x_s for b in batches]), x_t=torch.cat([b.x_t for b in batches]), edge_index=graph,
               cons_ptr=torch.from_numpy(c_ptr.astype(np.int32)), vars_ptr=torch.from_numpy(v_ptr.astype(np.int32)))
    if all(hasattr(b, "y_s") for b in batches):
        out.y_s, out.y_t = torch.cat([b.y_s for b in batches]), torch.cat([b.y_t for b in batches])
    out.n_lps = len(batches)
    out.lp_sizes = (ms, ns)                        # host-side sizes of the packed LPs (constraints, variables)
    out.bs = out.batch_size = M + N
    out.s_bs, out.t_bs = M, N
    return out
