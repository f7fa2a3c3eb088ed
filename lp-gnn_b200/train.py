"""Drop-in for the reference ``train.py`` (run_exp, 55-157) with data-parallel training added:

* same loop: one LP graph per step (``DataLoader(batch_size=1)``, train.py:70), ``model(batch)``, loss from
  ``--loss`` (balanced / unbalanced / focal), Adam/SGD with weight decay 5e-4 and StepLR(epochs//4, 0.1),
  weights saved to ``{log_dir}/mdl.pth`` every epoch;
* launched under ``torchrun`` (one process per GPU) every rank trains on its own shard of the LP graphs and the
  gradients are averaged with ONE NCCL all-reduce of the flattened 16.9 MB gradient bucket per step (the whole
  model is one bucket) -- effective batch = WORLD_SIZE graphs per optimiser step.

The per-step host syncs of the reference (``isnan(loss).item()``, ``loss.item()``, sklearn accuracy every step,
train.py:126-137) are reduced to one ``.item()`` every ``--log_every`` steps.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import time

import numpy as np
import torch
from torch.optim.lr_scheduler import StepLR

from .arch import *  # noqa: F401,F403  (``eval(args.arch)`` resolves GCN_FC(...) like the reference, train.py:79)
from .data import DataLoader
from .dataset import LPDataset, MyToBipartite, pack_bipartite
from .io_utils import shard_indices, split_train_val
from .losses import LOSSES, FocalLoss, balanced, balanced_packed, focal, unbalanced  # noqa: F401  (names of train.py:18-53)
from .val import accuracy


def dist_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend=None):
    rank, world, local = dist_info()
    if world > 1 and not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        torch.distributed.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_parameters(model, world):
    if world > 1:
        for p in model.parameters():
            torch.distributed.broadcast(p.data, src=0)


def _shared_flat_view(grads):
    """The gradients as ONE flat tensor when they already are consecutive views of a single buffer (the native training
    step hands them out that way, 16-byte aligned), else None."""
    g0 = grads[0]
    st = g0.untyped_storage()
    end = g0.storage_offset()
    for g in grads:
        if (g.dtype != g0.dtype or not g.is_contiguous() or g.untyped_storage().data_ptr() != st.data_ptr()
                or g.storage_offset() < end or g.storage_offset() - end > 3):
            return None
        end = g.storage_offset() + g.numel()
    return torch.empty(0, dtype=g0.dtype, device=g0.device).set_(st, g0.storage_offset(), (end - g0.storage_offset(),))


def _aligned_offsets(params):
    """Offsets of the parameters in the flat gradient buffer of the native training step (16-byte aligned views)."""
    offs = [0]
    for p in params:
        offs.append(offs[-1] + (p.numel() + 3) // 4 * 4)
    return offs


def allreduce_gradients(params, world):
    """Mean of the gradients over ranks with one collective.  NCCL over NVLink/NVSwitch on the GPU box; gloo in the
    CPU tests.  Deterministic (fixed bucket layout).  Every rank reduces a buffer of the SAME length whichever path
    produced its gradients: the native step's flat buffer in place, anything else staged into that layout."""
    if world == 1:
        return
    from .training import consume_synced_backward
    if consume_synced_backward():              # reduced inside the native backward, overlapped with its second half
        return
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    grads = [p.grad for p in params]
    flat = _shared_flat_view(grads)
    offs = _aligned_offsets(params)
    count = offs[-2] + params[-1].numel()
    if flat is not None and flat.numel() == count:   # (alignment gaps of <= 3 floats between the views are reduced too: harmless)
        torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
        flat.mul_(1.0 / world)
        return
    flat = torch.zeros(count, dtype=grads[0].dtype, device=grads[0].device)
    for g, o in zip(grads, offs):
        flat[o:o + g.numel()].copy_(g.reshape(-1))
    torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
    flat.mul_(1.0 / world)
    for g, o in zip(grads, offs):
        g.copy_(flat[o:o + g.numel()].view_as(g))


_host_group = [None]


def host_group(world):
    """gloo side group for host-side agreements between ranks (plain integers): going through the CPU keeps them off
    the GPU stream, so an agreement never drains the launch queue the way a NCCL all-reduce + ``.item()`` would."""
    if world == 1:
        return None
    if _host_group[0] is None:
        dist = torch.distributed
        _host_group[0] = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
    return _host_group[0]


def agreed_count(n_local, world):
    """MAX over ranks of a per-rank step count: all ranks then take part in the same number of gradient collectives."""
    if world == 1:
        return int(n_local)
    t = torch.tensor([int(n_local)], dtype=torch.int64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=host_group(world))
    return int(t[0])


def idle_step(model, params, opt, world):
    """A rank that has no k-th mini-batch where another rank has one (an empty graph, fewer sampled mini-batches for its
    LP) joins the collective with ZERO gradients and applies the same averaged update: collective counts and replicas
    stay identical, nothing hangs.  (The mean still divides by the world size.)"""
    from . import training
    offs = _aligned_offsets(params)
    overlapped = training._gradient_sync[0] is not None and world > 1
    flat = torch.zeros(offs[-1] if overlapped else offs[-2] + params[-1].numel(), dtype=params[0].dtype, device=params[0].device)
    for p, o in zip(params, offs):
        p.grad = flat[o:o + p.numel()].view_as(p)
    if overlapped:     # the two collectives of an active rank's native backward (tail, then head), same cut
        nh = len(getattr(model, "layers", []))
        cut = offs[6 + 6 * (nh - 1)] if nh > 0 else offs[6]
        handles = [training._gradient_sync[0](flat[cut:]), training._gradient_sync[0](flat[:cut])]
        for h in handles:
            if h is not None:
                h.wait()
        flat.mul_(training._gradient_sync[1])
    else:
        allreduce_gradients(params, world)
    opt.step()


def parse_args(argv=None, **defaults):
    """The flags of ``utils.Environment`` that the training / prediction entry points read (utils.py:743-770)."""
    ap = argparse.ArgumentParser(conflict_handler="resolve")
    ap.add_argument("--dev", type=int, default=0)
    ap.add_argument("--overlap_allreduce", type=int, default=0)
    ap.add_argument("--exp_nm", type=str, default="tmp")
    ap.add_argument("--opt", type=str, default="adam")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--arch", type=str, default="GCN_FC(8,8,hids=128)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--num_workers", type=int, default=0)
    ap.add_argument("--load_from", type=str, default="None")
    ap.add_argument("--dataset", type=str, default="None")
    ap.add_argument("--solver_prefix", type=str, default="highs-")
    ap.add_argument("--data_prefix", type=str, default="./lp-dataset/")
    ap.add_argument("--log_prefix", type=str, default="./runs/")
    ap.add_argument("--edge_num_thresh", type=float, default=4e6 * 3)
    ap.add_argument("--batch_size", type=int, default=int(4096 * 40 * 2),
                    help="seed nodes per sampled mini-batch for LPs above edge_num_thresh (utils.py:807)")
    ap.add_argument("--loss", type=str, default="balanced")
    ap.add_argument("--inference_manager", type=str, default="InferenceManager(0,)")
    ap.add_argument("--split", type=str, default="val")
    ap.add_argument("--fp16", type=int, default=0)
    ap.add_argument("--log_every", type=int, default=9)
    ap.add_argument("--packed", type=int, default=0, help="pred_basis: pack LPs block-diagonally (sweep mode)")
    ap.add_argument("--pack", type=int, default=1,
                    help="train: LP graphs per optimiser step, packed block-diagonally (balanced loss; 1 = the reference's loop)")
    ap.add_argument("--dataset_processed_prefix", type=str, default=None)
    ap.add_argument("--log_dir", type=str, default=None)
    ap.set_defaults(**defaults)
    args, _ = ap.parse_known_args(argv)
    if args.dataset_processed_prefix is None:       # utils.py:835-836
        args.dataset_processed_prefix = f"{args.data_prefix}/{args.dataset}/{args.solver_prefix}inp_tgt/"
    if args.log_dir is None:
        args.log_dir = f"{args.log_prefix}/{args.exp_nm}/"
    return args


def run_exp(args):
    rank, world, local = init_distributed()
    if not torch.cuda.is_available():
        raise RuntimeError("training needs CUDA devices: the lp-gnn hot path has no CPU fallback")
    dev = torch.device("cuda", local if world > 1 else args.dev)
    torch.cuda.set_device(dev)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    os.makedirs(args.log_dir, exist_ok=True)
    ds = LPDataset(args.dataset_processed_prefix, transform=MyToBipartite(thresh_num=args.edge_num_thresh))
    train_ds, _ = split_train_val(ds, args.seed)
    if world > 1:                                    # one shard of LP graphs per GPU
        train_ds = train_ds[np.asarray(shard_indices(len(train_ds), rank, world))]
    loader = DataLoader(train_ds, batch_size=1, shuffle=True, drop_last=True, num_workers=args.num_workers,
                        persistent_workers=args.num_workers > 0, pin_memory=False)
    model = eval(args.arch)                          # noqa: S307  (the reference's plugin mechanism, train.py:79)
    if args.load_from.lower() != "none":
        model.load(args.load_from)
    model = model.to(dev)
    if args.fp16:
        model.bfloat16()     # 16-bit TRAINING uses bf16 storage (half would need loss scaling); `.half()` is inference-only
    broadcast_parameters(model, world)
    from .training import enable_overlapped_allreduce
    # opt-in (--overlap_allreduce 1): tail gradients all-reduced under the rest of the backward pass.  Measured on 2 x B200:
    # no gain (2.57 vs 2.56 ms per C2 step) -- NCCL's kernel does not co-reside with the persistent tcgen05 GEMM CTAs (one
    # per SM, ~200 KB of shared memory each), so the collective still runs between kernels.
    enable_overlapped_allreduce(world if getattr(args, "overlap_allreduce", 0) else 1)
    params = list(model.parameters())
    if args.opt == "adam":
        opt = torch.optim.Adam(params, lr=args.lr, weight_decay=5e-4, fused=True)   # same update rule, one kernel
    else:
        opt = torch.optim.SGD(params, lr=args.lr, weight_decay=5e-4)
    scheduler = StepLR(opt, step_size=max(args.epochs // 4, 1), gamma=0.1)
    loss_fn = LOSSES[args.loss]
    glstep, history = 0, []
    steps_per_epoch = len(loader)
    if world > 1:                                    # every rank must take the same number of optimiser steps
        t = torch.tensor([steps_per_epoch], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        steps_per_epoch = int(t.item())
    t0 = time.time()
    pack_k = max(int(getattr(args, "pack", 1)), 1)
    if pack_k > 1 and args.loss != "balanced":
        raise ValueError("--pack > 1 trains with the per-LP balanced loss (losses.balanced_packed)")
    pending = []
    for epoch in range(args.epochs):
        model.train()
        for it, batch in enumerate(loader):
            if it >= steps_per_epoch:
                break
            if pack_k > 1 and hasattr(batch, "x_s"):
                # mini-batches of LP graphs: `pack` consecutive LPs form one block-diagonal graph and ONE optimiser step
                pending.append(batch)
                if len(pending) < pack_k and it + 1 < steps_per_epoch:
                    continue
                batch, pending = pack_bipartite(pending), []
            if hasattr(batch, "x_s"):                # already bipartite: one whole LP per step (train.py:103-104)
                sub_loader = [batch]
            elif batch.edge_index.shape[-1] == 0:    # train.py:106 skips empty graphs
                sub_loader = []
            else:                                    # above edge_num_thresh: sampled mini-batches (train.py:105-116)
                from .sampling import NeighborSubgraphLoader, ResidentLP, conv_depth
                lp_res = ResidentLP.from_unipartite(batch, dev)
                sub_loader = NeighborSubgraphLoader(lp_res, [6] * conv_depth(args.arch), min(args.batch_size, lp_res.num_nodes),
                                                    shuffle=True, drop_last=True, seed=args.seed + glstep)
            # data parallel: ranks whose LP yields fewer mini-batches (or none) than a peer's still join every gradient
            # collective of this outer step (idle_step), so the collective count is rank-invariant
            n_steps = agreed_count(len(sub_loader), world)
            sub_iter = iter(sub_loader)
            for _k in range(n_steps):
                batch = next(sub_iter, None)
                if batch is None:
                    idle_step(model, params, opt, world)
                    continue
                batch.to(dev, non_blocking=True)
                glstep += 1
                logit_cons, logit_vars = model(batch)
                logit_cons, logit_vars = logit_cons[:batch.s_bs], logit_vars[:batch.t_bs]
                y_s, y_t = batch.y_s[:batch.s_bs], batch.y_t[:batch.t_bs]
                if hasattr(batch, "cons_ptr"):       # a pack: per-LP class weights, mean over the pack's LPs
                    loss = balanced_packed(logit_cons, logit_vars, y_s, y_t, batch.cons_ptr, batch.vars_ptr)
                else:
                    loss = loss_fn(logit_cons, logit_vars, y_s, y_t)
                opt.zero_grad()
                loss.backward()
                allreduce_gradients(params, world)
                opt.step()
                if (glstep - 1) % max(args.log_every, 1) == 0:
                    lv = float(loss.item())
                    assert not np.isnan(lv)              # train.py:126
                    # train.py:131-137 (acc_meter): the reference scores every step; here on the logged steps only,
                    # since accuracy() ends in a host read
                    m0, n0 = (batch.lp_sizes[0][0], batch.lp_sizes[1][0]) if hasattr(batch, "lp_sizes") else \
                        (logit_cons.shape[0], logit_vars.shape[0])         # a pack is scored on its first LP
                    acc = float(accuracy(torch.cat((logit_cons[:m0], logit_vars[:n0]), dim=0).detach(),
                                         torch.cat((y_s[:m0], y_t[:n0]), dim=0), m0))
                    history.append(dict(epoch=epoch, step=glstep, loss=lv, acc=acc, lr=scheduler.get_last_lr()[0]))
                    if rank == 0:
                        logging.info(f"{epoch} {it}/{steps_per_epoch} step {glstep} loss {lv:.4f}")
        scheduler.step()
        if rank == 0:
            model.save(f"{args.log_dir}/mdl.pth")
    torch.cuda.synchronize()
    if rank == 0:
        model.save(f"{args.log_dir}/mdl.pth")
        with open(f"{args.log_dir}/train_log.json", "w") as f:
            json.dump(dict(history=history, seconds=time.time() - t0, steps=glstep, world=world), f)
    if world > 1:
        torch.distributed.barrier()
    enable_overlapped_allreduce(1)
    return model, history


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    run_exp(parse_args())
