"""Drop-in for the reference ``scripts/pred_basis.py``: for every LP run the model, decide the basis
(``inference_gnn``) and write a HiGHS ``.bas`` file plus the ``.bas.sort`` probability file
(pred_basis.py:14-23, 57-67, 70-111); then the per-LP inference timing loop (157-178).

Differences in mechanism, not in outputs: the basis decision runs on the device (no logits round trip), the
statuses come back as one uint8 D2H copy per LP, files are written by a small thread pool instead of a thread
per file, and under ``torchrun`` the LPs are sharded over the ranks (independent units, no collective).
"""
from __future__ import annotations

import json
import logging
import os
import os.path as osp
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops
from .arch import *  # noqa: F401,F403
from .data import DataLoader
from .dataset import LPDataset, MyToBipartite
from .io_utils import batch_to, extract_fn, mkdir_p, shard_indices, split_train_val
from .train import init_distributed, parse_args
from .val import InferenceManager  # noqa: F401  (eval(args.inference_manager))


def enc_vec(v):
    return " ".join(str(x) for x in v)


def write_bas_highs(fn, vnms=None, cnms=None, vbas=None, cbas=None):
    """HiGHS basis file (pred_basis.py:14-23): ``HIGHS v1 / Valid / # Columns n / <n ints> / # Rows m / <m ints>``."""
    assert vbas is not None
    mkdir_p(osp.dirname(fn))
    with open(fn, "w") as f:
        f.write("HIGHS v1\nValid\n")
        f.write(f"# Columns {len(vbas)}\n")
        f.write(enc_vec(np.asarray(vbas).tolist()) + "\n")
        f.write(f"# Rows {len(cbas)}\n")
        f.write(enc_vec(np.asarray(cbas).tolist()) + "\n")


def write_bas(fn, var_nms, con_nms, pred_var, pred_con):
    """Named (MPS-style) basis file (pred_basis.py:25-55, the alternative ``write_func``): basic variables are paired
    with the constraints at their lower (``XL``) then upper (``XU``) bound, in file order; then ``UL`` for the
    variables at their upper bound.  Variables at the lower bound are the format's default and are not written."""
    var_nms, con_nms = np.asarray(var_nms), np.asarray(con_nms)
    pred_var, pred_con = np.asarray(pred_var), np.asarray(pred_con)
    basic_vars = var_nms[pred_var == 1]
    cons_lo, cons_up = con_nms[pred_con == 0], con_nms[pred_con == 2]
    assert len(basic_vars) == len(cons_lo) + len(cons_up)
    lines = [f"NAME          0.mps  Iterations 0  Rows {len(con_nms)}  Cols {len(var_nms)} \n"]
    lines += [f" XL {v} {c} \n" for v, c in zip(basic_vars[:len(cons_lo)], cons_lo)]
    lines += [f" XU {v} {c} \n" for v, c in zip(basic_vars[len(cons_lo):], cons_up)]
    lines += [f" UL {v} \n" for v in var_nms[pred_var == 2]]
    lines.append("ENDATA")
    with open(fn, "w") as f:
        f.writelines(lines)


def read_bas_highs(fn):
    """HiGHS basis file -> ``(con_status, var_status)`` int arrays (reference scripts/cvt_to_pkl.py:166-178)."""
    assert os.path.exists(fn), fn
    with open(fn, "r") as f:
        lines = f.readlines()
    var_stas = con_stas = None
    for i, line in enumerate(lines):
        if "Columns" in line:
            var_stas = np.array(lines[i + 1].split(), "int")
        if "Rows" in line:
            con_stas = np.array(lines[i + 1].split(), "int")
    return con_stas, var_stas


def read_bas(fn, con_nms=None, var_nms=None):
    """Basis file -> ``(con_labels, var_labels)`` in {0: at lower, 1: basic, 2: at upper} (reference
    scripts/cvt_to_pkl.py:180-209; the reader behind ``val.validation_wrt_converged``).  Named (MPS-style) files: ``XL`` /
    ``XU`` pair a basic variable with a constraint at its lower / upper bound, ``LL`` / ``UL`` / ``BS`` set a variable;
    unnamed constraints default to basic, unnamed variables to the lower bound.  HiGHS files are recognised by their first
    line -- the reference tests for ``'HiGHS'`` while its own writer emits ``'HIGHS v1'``; both spellings are accepted."""
    status_str_to_int = {"LL": 0, "BS": 1, "UL": 2, "XU": (1, 2), "XL": (1, 0)}
    with open(fn, "r") as f:
        lines = f.readlines()
    if lines and ("HiGHS" in lines[0] or "HIGHS" in lines[0]):
        return read_bas_highs(fn)
    assert con_nms is not None
    con_label, var_label = {}, {}
    for line in lines:
        parts = line.split()
        if not parts or parts[0] not in status_str_to_int:
            continue
        st = status_str_to_int[parts[0]]
        if parts[0] in ("XU", "XL"):
            var_label[parts[1]], con_label[parts[2]] = st
        else:
            var_label[parts[1]] = st
    return (np.array([con_label.get(nm, 1) for nm in con_nms], dtype=np.int64),
            np.array([var_label.get(nm, 0) for nm in var_nms], dtype=np.int64))


def write_sort_vars(fn, logits, m):
    """``.bas.sort`` file (pred_basis.py:57-67): P(basic) of the variables, then of the constraints.  Two call
    forms: the reference's ``(fn, logits[m+n,3], m)`` (constraints first; softmax taken here), or
    ``(fn, p_basic_vars, p_basic_cons)`` with the probabilities already split (what ``run`` has in hand after the
    device-side decision)."""
    if isinstance(m, (int, np.integer)):
        pr = torch.softmax(torch.as_tensor(logits), dim=-1)[:, 1].cpu().numpy()
        p_basic_vars, p_basic_cons = pr[int(m):], pr[:int(m)]
    else:
        p_basic_vars, p_basic_cons = logits, m
    with open(fn, "w") as f:
        f.write(f"{len(p_basic_vars)} \n")
        f.write(enc_vec(p_basic_vars) + "\n")
        f.write(f"{len(p_basic_cons)} \n")
        f.write(enc_vec(p_basic_cons) + "\n")


@torch.no_grad()
def predict_one(model, batch, dev, fp16=False, args=None):
    """model -> logits -> (status uint8 [m+n], P(basic) [m+n]) on the host (pred_basis.py:70-85).  LPs above
    ``edge_num_thresh`` arrive unipartite (no ``x_s``) and go through ``val.model_inference_with_batch`` -- the sampled
    full-neighbourhood path -- exactly as the reference routes EVERY batch (pred_basis.py:79)."""
    if hasattr(batch, "x_s"):
        batch = batch_to(batch, dev, fp16)
        lc, lv = model(batch)
        lc, lv = lc[:batch.s_bs], lv[:batch.t_bs]
    else:
        from .val import model_inference_with_batch
        lc, lv = (t.to(dev).float() for t in model_inference_with_batch(model, batch, args))
    status = ops.basis_select(lc, lv, k_basic=lc.shape[0], int64=False)
    p1 = torch.softmax(torch.cat((lc, lv), 0), dim=-1)[:, 1]      # [m+n] values for the .sort file (tiny)
    return status.cpu().numpy(), p1.cpu().numpy(), lc.shape[0]


@torch.no_grad()
def inference_only(model, batch):
    """pred_basis.py:113-118: the timed region of the per-LP benchmark."""
    lc, lv = model(batch)
    return ops.basis_select(lc, lv, k_basic=lc.shape[0], int64=False)


def run(args):
    rank, world, local = init_distributed()
    if not torch.cuda.is_available():
        raise RuntimeError("pred_basis needs a CUDA device: the lp-gnn hot path has no CPU fallback")
    dev = torch.device("cuda", local if world > 1 else args.dev)
    torch.cuda.set_device(dev)
    inf_mng = eval(args.inference_manager)           # noqa: S307
    folder = inf_mng.get_basis_folder()
    model = eval(args.arch).to(dev)                  # noqa: S307
    if args.load_from.lower() != "none":
        model.load(args.load_from)
    if args.fp16:
        model.half()
    model.eval()
    out_dir = f"{args.log_dir}/{folder}/"
    mkdir_p(out_dir)
    ds = LPDataset(args.dataset_processed_prefix, MyToBipartite(thresh_num=args.edge_num_thresh), load_meta=True)
    train_ds, val_ds = split_train_val(ds, args.seed)
    idxs = list(val_ds.indices()) if args.split == "val" else list(val_ds.indices()) + list(train_ds.indices())
    mine = [idxs[i] for i in shard_indices(len(idxs), rank, world)]
    sub = ds[np.asarray(mine, dtype=np.int64)] if mine else None
    times = {}
    pool = ThreadPoolExecutor(max_workers=4)
    futures = []
    if sub is not None and getattr(args, "packed", 0):
        # sweep mode: LPs packed block-diagonally, one native forward per pack, segmented basis decision, no
        # per-LP Python round trips (the .sort probability files are not produced in this mode)
        from .pipeline import PackedBasisPipeline, pack_lp
        loader = DataLoader(sub, batch_size=1, shuffle=False, num_workers=args.num_workers)
        names, hosts = [], []
        for batch in loader:
            if not hasattr(batch, "x_s"):            # above edge_num_thresh: not packable, the per-LP sampled path takes it
                fn = extract_fn(batch.processed_path[0])
                status, _, m = predict_one(model, batch, dev, bool(args.fp16), args)
                futures.append(pool.submit(write_bas_highs, f"{out_dir}/{fn}.bas", None, None, status[m:], status[:m]))
                continue
            r, c, v = batch.edge_index._coo
            names.append(extract_fn(batch.processed_path[0]))
            hosts.append(pack_lp(r.numpy(), c.numpy(), v.numpy(), batch.x_s.numpy(), batch.x_t.numpy(),
                                 is_sorted=batch.edge_index._sorted_hint))
        for i, status in PackedBasisPipeline(model, dev).run(hosts):
            m = hosts[i].m
            futures.append(pool.submit(write_bas_highs, f"{out_dir}/{names[i]}.bas", None, None, status[m:], status[:m]))
    elif sub is not None:
        loader = DataLoader(sub, batch_size=1, shuffle=False, num_workers=args.num_workers)
        for batch in loader:
            fn = extract_fn(batch.processed_path[0])
            status, p1, m = predict_one(model, batch, dev, bool(args.fp16), args)
            futures.append(pool.submit(write_bas_highs, f"{out_dir}/{fn}.bas", None, None, status[m:], status[:m]))
            futures.append(pool.submit(write_sort_vars, f"{out_dir}/{fn}.bas.sort", p1[m:], p1[:m]))
        # timing pass (pred_basis.py:157-178): model + basis decision per LP, inputs already on the device
        loader = DataLoader(sub, batch_size=1, shuffle=False, num_workers=args.num_workers)
        for batch in loader:
            fn = extract_fn(batch.processed_path[0])
            if hasattr(batch, "x_s"):
                batch = batch_to(batch, dev, bool(args.fp16))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if hasattr(batch, "x_s"):
                inference_only(model, batch)
            else:                                    # sampled path: sampling + forward per seed batch + basis decision
                predict_one(model, batch, dev, bool(args.fp16), args)
            torch.cuda.synchronize()
            times[fn] = time.perf_counter() - t0
    for f in futures:
        f.result()
    pool.shutdown()
    with open(f"{args.log_dir}/inf_time.rank{rank}.json", "w") as f:
        json.dump(times, f)
    if world > 1:
        torch.distributed.barrier()
    return times


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    run(parse_args())
