"""Drop-in for the hot-path parts of the reference ``val.py``: ``inference_gnn`` (106-124),
``model_inference_with_batch`` (12-36), ``validation`` (43-69), ``accuracy`` (199-237), ``InferenceManager`` naming
(167-197) and the entry point (238-281: ``python -m lpgnn_b200.val --arch ... --load_from .../mdl.pth``)."""
from __future__ import annotations

import json
import logging
import os

import numpy as np
import torch

from . import ops
from .io_utils import batch_to


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("inference needs a CUDA (sm_100) device: the lp-gnn hot path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


@torch.no_grad()
def inference_gnn(logits, m, **kwargs):
    """Reference val.py:106-124 on the device: softmax, NaN -> 0, global top-m on P(basic) -> status 1, the rest
    argmax over {0, 2}.  ``logits`` [m+n,3] (constraints first); CPU inputs are moved to the GPU and the result
    is returned on the input's device, int64 like the reference."""
    src_dev = logits.device
    lg = logits.float()
    if not lg.is_cuda:
        lg = lg.to(_device())
    pred, counts = ops.basis_select(lg[:m], lg[m:], k_basic=m, int64=True, want_counts=True)
    c = counts.tolist()                                   # one host sync, as the reference's asserts (val.py:118-122)
    n = logits.shape[0] - m
    assert c[1] == m, (c, m)                              # #basic == number of rows
    assert c[3] == m - (c[1] - c[3]), c                   # basic structurals == non-basic rows
    return pred.to(src_dev)


@torch.no_grad()
def model_inference_with_batch(model, batched_graphs, args=None):
    """Reference val.py:12-36.  Bipartite batches (every LP below ``edge_num_thresh``) run the full-graph path; a
    unipartite graph (above the threshold) is kept resident on the device and walked in seed batches of
    ``args.batch_size`` nodes with FULL neighbourhoods over ``depth - 1`` hops (``num_neighbors=[-1]*depth``,
    val.py:22-27) by ``sampling.NeighborSubgraphLoader``.  Returns CPU logits like the reference."""
    dev = _device()
    model.eval()
    fp16 = bool(getattr(args, "fp16", 0))
    if hasattr(batched_graphs, "x_s"):
        batch = batch_to(batched_graphs, dev, fp16)
        logit_cons, logit_vars = model(batch)
        return logit_cons[:batch.s_bs].cpu(), logit_vars[:batch.t_bs].cpu()
    from .sampling import NeighborSubgraphLoader, ResidentLP, conv_depth
    lp = ResidentLP.from_unipartite(batched_graphs, dev)
    depth = conv_depth(getattr(args, "arch", ""))
    bs = min(int(getattr(args, "batch_size", 327_680)), lp.num_nodes)
    lc, lv = [], []
    for batch in NeighborSubgraphLoader(lp, [-1] * depth, bs, shuffle=False, drop_last=False):
        if fp16:
            batch.x_s, batch.x_t = batch.x_s.half(), batch.x_t.half()
        logit_cons, logit_vars = model(batch)
        lc.append(logit_cons[:batch.s_bs].cpu())
        lv.append(logit_vars[:batch.t_bs].cpu())
    return torch.cat(lc, dim=0), torch.cat(lv, dim=0)


@torch.no_grad()
def validation(model, loader, dev, dump_info=None, args=None):
    """Reference val.py:43-69: per LP, logits through ``model_inference_with_batch`` and ``accuracy(...,
    return_pr=True)`` against the labels of the bipartite view; returns ``(avg_loss, avg_acc)`` with ``avg_loss`` = 0
    like the reference (it never accumulates a loss here).  The model's training flag is restored.  ``dump_info``:
    the reference updates the acc / prec / recl columns of a pandas HDF file (``tables`` is not on this image); here
    it names a JSON file that receives ``{file name: {acc, prec, recl}}``."""
    from .dataset import MyToBipartite
    from .io_utils import extract_fn
    model.to(dev)
    was_training = model.training
    model.eval()
    avg_loss = avg_acc = 0.
    rows = {}
    n_lps = len(loader)
    for idx, batch in enumerate(loader):
        fn = extract_fn(batch.processed_path[0])
        logit_cons, logit_vars = model_inference_with_batch(model, batch, args)
        if not hasattr(batch, "x_s"):                     # above edge_num_thresh: labels from the bipartite view
            batch = MyToBipartite(thresh_num=np.inf)(batch)
        acc, prec, recl = accuracy(torch.cat((logit_cons, logit_vars), dim=0),
                                   torch.cat((batch.y_s, batch.y_t), dim=0).cpu(), logit_cons.shape[0], return_pr=True,
                                   dataset_name=getattr(args, "dataset", "") or "")
        avg_acc += acc / n_lps
        rows[fn] = dict(acc=float(acc), prec=float(prec), recl=float(recl))
        if idx % 9 == 1:
            logging.info(f"{idx} {fn} {n_lps} {acc} {prec} {recl}")
    if was_training:
        model.train()
    if dump_info:
        with open(dump_info, "w") as f:
            json.dump(rows, f)
    return avg_loss, avg_acc


@torch.no_grad()
def validation_wrt_converged(model, loader, dev, dump_info=None, args=None):
    """Reference val.py:71-104: accuracy of the predicted basis against the basis the solver CONVERGED to when started
    from it (``{log_dir}/opt-from-pred-basis/{fn}.bas``, written by the solver harness, which is out of scope here); LPs
    without such a file are skipped.  Returns ``(0, avg_acc)`` like the reference; ``dump_info`` names a JSON file that
    receives ``{file name: {"cvg/acc", "cvg/prec", "cvg/recl"}}`` (the reference updates a pandas HDF file)."""
    from .io_utils import extract_fn
    from .pred_basis import read_bas
    model.to(dev)
    was_training = model.training
    model.eval()
    avg_acc, rows, n_lps = 0., {}, len(loader)
    for idx, batch in enumerate(loader):
        con_nms, var_nms = batch.con_nms[0], batch.var_nms[0]
        fn = extract_fn(batch.processed_path[0])
        logit_cons, logit_vars = model_inference_with_batch(model, batch, args)
        tgt = f"{args.log_dir}/opt-from-pred-basis/{fn}.bas"
        if not os.path.exists(tgt):
            continue
        con_lbls, var_lbls = read_bas(tgt, con_nms, var_nms)
        acc, prec, recl = accuracy(torch.cat((logit_cons, logit_vars), dim=0),
                                   torch.cat((torch.from_numpy(np.asarray(con_lbls)), torch.from_numpy(np.asarray(var_lbls))), dim=0),
                                   logit_cons.shape[0], return_pr=True)
        avg_acc += acc / n_lps
        rows[fn] = {"cvg/acc": float(acc), "cvg/prec": float(prec), "cvg/recl": float(recl)}
        if idx % 9 == 1:
            logging.info(f"{idx} {fn} {n_lps} {acc} {prec} {recl}")
    if was_training:
        model.train()
    if dump_info:
        with open(dump_info, "w") as f:
            json.dump(rows, f)
    return 0, avg_acc


def run(args):
    """Reference val.py:238-281: load the dataset, split, build ``eval(args.arch)``, load the checkpoint, optional
    ``model.half()`` and report the mean validation accuracy."""
    from .arch import GCN_FC  # noqa: F401  (eval(args.arch))
    from .data import DataLoader
    from .dataset import LPDataset, MyToBipartite
    from .io_utils import split_train_val
    _device()                                             # raises without a CUDA device: no CPU fallback
    dev = torch.device("cuda", int(args.dev))
    torch.cuda.set_device(dev)
    os.makedirs(args.log_dir, exist_ok=True)
    ds = LPDataset(args.dataset_processed_prefix, transform=MyToBipartite(thresh_num=args.edge_num_thresh))
    assert len(ds) != 0, "should not empty"
    _, val_ds = split_train_val(ds, args.seed)
    loader = DataLoader(val_ds, batch_size=1, shuffle=False, drop_last=False, num_workers=args.num_workers)
    model = eval(args.arch).to(dev)                       # noqa: S307  (the reference's plugin mechanism, val.py:266)
    if args.load_from.lower() != "none":
        model.load(args.load_from)
    if args.fp16:
        model.half()
    model.eval()
    avg_loss, avg_acc = validation(model, loader, dev, dump_info=f"{args.log_dir}/val_metrics.json", args=args)
    print("avg val acc", avg_acc)
    return avg_loss, avg_acc


@torch.no_grad()
def accuracy_counts(logits, gt, num_cons):
    """The integers behind ``accuracy`` without leaving the device: int32[12] = the 4 counts of ``lpgnn_basis_select``
    (nodes with status 0 / 1 / 2, basic variables) followed by, per side, {#pred == gt, #pred == 1 and gt == 1, #pred == 1,
    #gt == 1} (``lpgnn_basis_metrics``).  No host sync; a training loop reads it on the steps it logs."""
    from . import _lib
    lg = logits.float()
    if not lg.is_cuda:
        lg = lg.to(_device())
    gt = gt.to(lg.device, torch.int64).contiguous()
    m = int(num_cons)
    n = lg.shape[0] - m
    pred, sel = ops.basis_select(lg[:m], lg[m:], k_basic=m, int64=False, want_counts=True)
    out = torch.empty(12, dtype=torch.int32, device=lg.device)
    out[:4] = sel
    with torch.cuda.device(lg.device):
        rc = _lib.load().lpgnn_basis_metrics(pred.data_ptr(), 0, gt[:m].data_ptr(), m, gt[m:].data_ptr(), n, out[4:].data_ptr(),
                                             _lib.stream_ptr())
    _lib.check(rc, "lpgnn_basis_metrics")
    return out


def metrics_from_counts(c, m, n, return_pr=False, dataset_name=""):
    """acc / precision / recall of reference val.py:199-237 from the 12 integers of ``accuracy_counts`` (sklearn's
    ``precision_score`` / ``recall_score`` with ``labels=[1], average='macro'``: TP / #pred == 1 and TP / #gt == 1, 0 when
    the denominator is empty)."""
    c = [int(x) for x in c]
    assert c[1] == m, (c, m)                              # #basic == number of rows (val.py:118-122)
    if c[4 + 2] == m and m > 0:
        logging.warning("warning: may collapse, basis==all slacks")
    div = lambda a, b: a / b if b else 0.0
    acc1, acc2 = div(c[4], m), div(c[8], n)
    p1, p2 = div(c[5], c[6]), div(c[9], c[10])
    r1, r2 = div(c[5], c[7]), div(c[9], c[11])
    if bool(dataset_name) and "stoch" in dataset_name:    # stoch constraints are always labelled non-basic
        acc1, p1, r1 = acc2, p2, r2
    acc = (acc1 + acc2) / 2.
    return (acc, (p1 + p2) / 2., (r1 + r2) / 2.) if return_pr else acc


@torch.no_grad()
def accuracy(logits, gt, num_cons, return_pr=False, dataset_name=""):
    """Reference val.py:199-237: mean of constraint / variable accuracy, macro precision / recall of class 1.  The
    basis decision and the confusion counts stay on the device; ONE read of twelve integers replaces the reference's
    two vector copies + sklearn."""
    m = int(num_cons)
    c = accuracy_counts(logits, gt, m).tolist()
    return metrics_from_counts(c, m, logits.shape[0] - m, return_pr, dataset_name)


class InferenceManager:
    """Reference val.py:167-197 (only ``inference_gnn`` is on the hot path; the other two are marked deprecated
    upstream, val.py:126)."""

    def __init__(self, which_func=0, mode=None, gnn_wei=None, run=0):
        self.which_func = ["inference_gnn", "inference_all_slacks", "inference_gnn_sparsity"][which_func]
        if self.which_func != "inference_gnn":
            raise NotImplementedError("only InferenceManager(0, ...) (inference_gnn) is supported")
        self.mode, self.gnn_wei, self.run = mode, gnn_wei, run

    def get_log_folder(self):
        return f"gnn-bas-{self.run}"

    def get_basis_folder(self):
        return "pred-basis" + (f"-{self.run}" if self.run != 0 else "")


if __name__ == "__main__":
    from .train import parse_args
    logging.basicConfig(level=logging.INFO)
    run(parse_args())
