"""File formats and small helpers of the reference's ``utils.py`` that the data layer needs:
msgpack with the msgpack_numpy ndarray extension (utils.py:193-224), the 70/30 split (utils.py:256-272),
``extract_fn`` (utils.py:301-309), ``batch_to`` (utils.py:909-915).  Host code only.

msgpack_numpy (not installed here) encodes an ndarray as the map
``{b'nd': True, b'type': dtype.str | descr, b'kind': b'' | b'V', b'shape': shape, b'data': raw bytes}`` and a
numpy scalar as ``{b'nd': False, b'type': dtype.str, b'data': raw bytes}``; this module restates that published
format so processed ``.pk`` files written by the reference can be read and vice versa.
"""
from __future__ import annotations

import logging
import os
import os.path as osp

import msgpack
import numpy as np


def _key(d, name):
    return d[name] if name in d else d[name.encode()]


def _np_encode(obj):
    if isinstance(obj, np.ndarray):
        if obj.dtype.kind == "O":
            return obj.tolist()
        return {b"nd": True, b"type": obj.dtype.str, b"kind": b"", b"shape": list(obj.shape),
                b"data": np.ascontiguousarray(obj).tobytes()}
    if isinstance(obj, (np.bool_, np.number)):
        return {b"nd": False, b"type": obj.dtype.str, b"data": obj.tobytes()}
    if isinstance(obj, complex):
        return {b"complex": True, b"data": repr(obj)}
    return obj


def _np_decode(obj):
    try:
        if "nd" in obj or b"nd" in obj:
            dtype = _key(obj, "type")
            dtype = dtype.decode() if isinstance(dtype, bytes) else dtype
            data = _key(obj, "data")
            if _key(obj, "nd"):
                return np.frombuffer(data, dtype=np.dtype(dtype)).reshape(_key(obj, "shape")).copy()
            return np.frombuffer(data, dtype=np.dtype(dtype))[0]
    except (KeyError, TypeError):
        pass
    return obj


def msgpack_dump(obj, file, **kwargs):
    with open(str(file), "wb") as fp:
        msgpack.pack(obj, fp, default=_np_encode, use_bin_type=True, **kwargs)


def msgpack_load(file, **kwargs):
    kwargs.pop("copy", None)
    assert osp.exists(file), file
    with open(str(file), "rb") as f:
        return msgpack.unpack(f, object_hook=_np_decode, use_list=True, raw=False, strict_map_key=False)


def mkdir_p(path):
    if path and not osp.exists(path):
        os.makedirs(path, exist_ok=True)


def split_idxs_train_val(ngraphs, seed=0):
    """utils.py:256-263: 70/30 permutation split with the legacy numpy global RNG."""
    ntrain = int(max(ngraphs * 7 / 10, 1))
    rs = np.random.RandomState(seed)            # == np.random.seed(seed); np.random.permutation(n)
    idxs = rs.permutation(ngraphs)
    return np.sort(idxs[:ntrain]), np.sort(idxs[ntrain:])


def split_train_val(ds, seed=0):
    if seed != 0:
        logging.warning("seed for train val not 0, will force set to 0")
        seed = 0
    tr, va = split_idxs_train_val(len(ds), seed)
    return ds[tr], ds[va]


def extract_fn(inp, suf=None):
    known = ["mps", "gz", "bas", "tar", "pk", "log", "lp", "sol" "txt", "json", "sort"]   # sic: 'sol' 'txt' are fused upstream
    res = ""
    for r in osp.basename(inp).split("."):
        if r not in known:
            res += r + "."
    return res[:-1]


def batch_to(batch, dev, half=False):
    """utils.py:909-915.  ``half`` selects the reduced-precision model path; the node features stay fp32 here
    (they are 8 wide and feed the fp32 gather of the input layer)."""
    for nm in ["edge_index", "x_s", "x_t", "y_s", "y_t"]:
        batch[nm] = batch[nm].to(dev)
    return batch


def shard_indices(n_items, rank, world, weights=None, equal_counts=False):
    """Deterministic partition of ``range(n_items)`` over ``world`` ranks: greedy longest-processing-time on
    ``weights`` (e.g. nnz) when given, round-robin otherwise.  ``equal_counts`` (with weights): deal the items in
    descending weight forwards then backwards over the ranks ("snake"), so shard sizes differ by at most one item and
    the shards carry a near-equal share of the weight -- for sweeps that process a fixed number of LPs per rank.
    Every rank computes the same answer locally."""
    if weights is None:
        return list(range(rank, n_items, world))
    order = np.argsort(-np.asarray(weights, dtype=np.float64), kind="stable")
    if equal_counts:
        pos = np.arange(n_items)
        lap, slot = pos // world, pos % world
        owner = np.where(lap % 2 == 0, slot, world - 1 - slot)
        return sorted(int(i) for i in order[owner == rank])
    load = np.zeros(world)
    mine = []
    for i in order:
        r = int(np.argmin(load))
        load[r] += float(weights[i])
        if r == rank:
            mine.append(int(i))
    return sorted(mine)
