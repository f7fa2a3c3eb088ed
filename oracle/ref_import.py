"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Imports the reference's OWN Python (arch.py, dataset.py, utils.py, val.py) verbatim from
/root/reference, with stand-ins for the third-party packages that are not installed in this
image (recipe: SURVEY.md Appendix C).  Only usable where /root/reference exists, i.e. in the
build container -- it is how ``tests/golden`` is generated and how ``oracle/port.py`` is
validated.  Nothing on the GPU box may depend on it.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import pyg_standins as S

REFERENCE_ROOT = os.environ.get("LPGNN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "arch.py"))


class _Data:
    """torch_geometric.data.Data: kwargs -> attributes; hasattr/del/[] supported."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __contains__(self, k):
        return hasattr(self, k)

    def to(self, *a, **k):
        return self


class _Dataset:
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root, self.transform = root, transform
        self._indices = None

    def indices(self):
        return range(self.len()) if self._indices is None else self._indices


class _BaseTransform:
    def __call__(self, data):
        raise NotImplementedError


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_CACHE = None


def load_reference():
    """Returns a namespace with the reference modules ``arch``, ``dataset``, ``utils``, ``val``."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")

    if not hasattr(np, "set_string_function"):  # removed in NumPy 2 (utils.py:22)
        np.set_string_function = lambda *a, **k: None

    stubs = {
        "seaborn": _module("seaborn"),
        "colorlog": _module("colorlog"),
        "easydict": _module("easydict", EasyDict=dict),
        "torch_sparse": _module("torch_sparse", SparseTensor=S.SparseTensor),
        "torch_geometric.utils": _module("torch_geometric.utils", to_undirected=S.to_undirected),
        "torch_geometric.nn": _module("torch_geometric.nn", GraphConv=S.GraphConv,
                                      LayerNorm=S._Unused, GENConv=S._Unused),
        "torch_geometric.typing": _module("torch_geometric.typing", Adj=object,
                                          OptPairTensor=object, OptTensor=object, Size=object),
        "torch_geometric.data": _module("torch_geometric.data", Data=_Data, Dataset=_Dataset),
        "torch_geometric.transforms": _module("torch_geometric.transforms",
                                              BaseTransform=_BaseTransform),
        "torch_geometric.loader": _module("torch_geometric.loader", DataLoader=object,
                                          DynamicBatchSampler=object, NeighborLoader=object),
    }
    tg = _module("torch_geometric")
    tg.utils = stubs["torch_geometric.utils"]
    tg.nn = stubs["torch_geometric.nn"]
    tg.data = stubs["torch_geometric.data"]
    tg.loader = stubs["torch_geometric.loader"]
    tg.transforms = stubs["torch_geometric.transforms"]
    stubs["torch_geometric"] = tg

    saved_modules = {k: sys.modules.get(k) for k in
                     list(stubs) + ["arch", "dataset", "utils", "val", "scripts", "scripts.cvt_to_pkl"]}
    saved_path = list(sys.path)
    try:
        sys.modules.update(stubs)
        # val.py:9 imports read_bas from scripts/cvt_to_pkl.py, which needs `mip` (absent).
        sys.modules["scripts"] = _module("scripts")
        sys.modules["scripts.cvt_to_pkl"] = _module("scripts.cvt_to_pkl", read_bas=None)
        for k in ("arch", "dataset", "utils", "val"):
            sys.modules.pop(k, None)
        sys.path.insert(0, REFERENCE_ROOT)
        import importlib
        ref_utils = importlib.import_module("utils")
        ref_arch = importlib.import_module("arch")
        ref_dataset = importlib.import_module("dataset")
        ref_val = importlib.import_module("val")
        # scripts/pred_basis.py (the writers of SURVEY 8 f-1): star-imports utils / arch / dataset / val, its
        # entry point sits under `if __name__ == '__main__'`
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_pred_basis",
                                                      os.path.join(REFERENCE_ROOT, "scripts", "pred_basis.py"))
        ref_pred = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_pred)
        # train.py (losses of SURVEY 8 a8; run_exp sits behind `if __name__ == '__main__'`)
        spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REFERENCE_ROOT, "train.py"))
        ref_train = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_train)
    finally:
        sys.path[:] = saved_path
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE = types.SimpleNamespace(arch=ref_arch, dataset=ref_dataset, utils=ref_utils,
                                   val=ref_val, pred_basis=ref_pred, train=ref_train, Data=_Data, SparseTensor=S.SparseTensor)
    return _CACHE
