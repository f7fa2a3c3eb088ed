"""Minimal stand-ins for the torch_geometric containers the reference's data layer uses
(``Data``, ``Dataset``, ``DataLoader(batch_size=1)``; reference dataset.py:3-11,267, train.py:10,70-77),
so the same ``dataset.py`` flow runs where PyG is not installed.  Pure host code."""
from __future__ import annotations

import copy
import os

import numpy as np
import torch
from torch.utils.data import DataLoader as _TorchLoader


class Data:
    """Attribute bag with PyG ``Data`` conveniences used by the reference: kwargs -> attributes,
    ``hasattr`` / ``del`` / ``[]``, ``.to(dev)`` moving every tensor-like field, ``num_nodes``."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return hasattr(self, key)

    def keys(self):
        return list(self.__dict__.keys())

    def to(self, device, non_blocking=False):
        for k, v in list(self.__dict__.items()):
            if hasattr(v, "to") and not isinstance(v, (str, bytes)):
                try:
                    setattr(self, k, v.to(device, non_blocking=non_blocking))
                except TypeError:
                    setattr(self, k, v.to(device))
        return self

    def pin_memory(self):
        for k, v in list(self.__dict__.items()):
            if hasattr(v, "pin_memory"):
                setattr(self, k, v.pin_memory())
        return self

    def __repr__(self):
        parts = []
        for k, v in self.__dict__.items():
            parts.append(f"{k}={list(v.shape)}" if isinstance(v, torch.Tensor) else f"{k}={v!r}"[:60])
        return f"{type(self).__name__}({', '.join(parts)})"


class Dataset(torch.utils.data.Dataset):
    """PyG ``Dataset`` subset: ``root``, ``raw_dir`` / ``processed_dir``, ``indices()``, integer and
    index-array ``__getitem__`` (the latter returns a shallow-copied sub-dataset, utils.py:272)."""

    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root = root
        self.transform = transform
        self._indices = None

    @property
    def raw_dir(self):
        return os.path.join(self.root, "raw")

    @property
    def processed_dir(self):
        return os.path.join(self.root, "processed")

    @property
    def raw_paths(self):
        return [os.path.join(self.raw_dir, f) for f in self.raw_file_names]

    def indices(self):
        return range(self.len()) if self._indices is None else self._indices

    def __len__(self):
        return len(self.indices())

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            data = self.get(self.indices()[int(idx)])
            return data if self.transform is None else self.transform(data)
        sub = copy.copy(self)
        idx = np.asarray(idx)
        base = np.asarray(list(self.indices()))
        sub._indices = [int(i) for i in (base[idx] if idx.dtype != bool else base[np.nonzero(idx)[0]])]
        return sub


def _collate_one(items):
    """batch_size=1 collate: the single graph, with string / list fields wrapped in a list exactly like PyG's
    collater does (callers index ``batch.processed_path[0]``, ``batch.con_nms[0]``)."""
    if len(items) != 1:
        raise ValueError("the lp-gnn loaders run with batch_size=1 (one LP graph per step, reference train.py:70)")
    data = items[0]
    for k, v in list(data.__dict__.items()):
        if isinstance(v, (str, list)):
            setattr(data, k, [v])
    return data


class DataLoader(_TorchLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        if kwargs.get("num_workers", 0) == 0:
            kwargs.pop("prefetch_factor", None)
            kwargs.pop("persistent_workers", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, collate_fn=_collate_one, **kwargs)
