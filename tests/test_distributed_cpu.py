"""N > 1 host logic on CPU: world_size-2 gloo processes run the gradient all-reduce and the LP sharding that the
multi-GPU training / prediction entry points use (the NCCL path is the same code with another backend)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch, io_utils, train
    r, w, _ = train.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                                   # different initial weights per rank
    model = arch.GCN_FC(8, 8, hids=64, depth=3)
    train.broadcast_parameters(model, world)                        # -> rank 0's weights everywhere
    torch.manual_seed(7 + rank)
    for p in model.parameters():
        p.grad = torch.randn_like(p)
    local = [p.grad.clone() for p in model.parameters()]
    train.allreduce_gradients(list(model.parameters()), world)
    g_cat = [p.grad.clone() for p in model.parameters()]
    # the native training step hands the gradients out as consecutive 16-byte-aligned views of ONE buffer: reduced in place
    params = list(model.parameters())
    assert train._shared_flat_view([p.grad for p in params]) is None
    offs = [0]
    for p in params:
        offs.append(offs[-1] + (p.numel() + 3) // 4 * 4)
    flat = torch.zeros(offs[-1])
    for p, o, l in zip(params, offs, local):
        p.grad = flat[o:o + p.numel()].view(p.shape)
        p.grad.copy_(l)
    fv = train._shared_flat_view([p.grad for p in params])
    assert fv is not None and fv.data_ptr() == flat.data_ptr() and fv.numel() == offs[-2] + params[-1].numel()
    train.allreduce_gradients(params, world)
    torch.save(dict(w=[p.detach().clone() for p in model.parameters()], g=g_cat, g_flat=[p.grad.clone() for p in params],
                    local=local, shard=io_utils.shard_indices(11, rank, world)), os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_broadcast_allreduce_and_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{i}.pt") for i in range(world))
    for a, b in zip(r0["w"], r1["w"]):
        assert torch.equal(a, b)                                    # broadcast made the replicas identical
    for g0, g1, l0, l1 in zip(r0["g"], r1["g"], r0["local"], r1["local"]):
        assert torch.equal(g0, g1)                                  # every rank holds the same averaged gradient
        assert torch.allclose(g0, (l0 + l1) / 2, atol=1e-6)
    for g0, gf0, gf1 in zip(r0["g"], r0["g_flat"], r1["g_flat"]):
        assert torch.equal(gf0, gf1) and torch.equal(gf0, g0)       # in-place flat path == flatten / unflatten path
    assert sorted(r0["shard"] + r1["shard"]) == list(range(11)) and not set(r0["shard"]) & set(r1["shard"])


def _uneven_worker(rank, world, port, out_dir):
    """Rank 0's LP yields 3 mini-batches, rank 1's only 1 (outer step 0); in outer step 1 rank 0 has an EMPTY graph
    (0 mini-batches) and rank 1 has 2: the collective count must still be the same on both ranks (train.agreed_count /
    train.idle_step), and the replicas must stay identical."""
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch, train
    train.init_distributed(backend="gloo")
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=64, depth=3)
    train.broadcast_parameters(model, world)
    params = list(model.parameters())
    opt = torch.optim.SGD(params, lr=0.1)
    schedule = [[3, 1], [0, 2]]                                      # [outer step][rank] -> local mini-batches
    gen = torch.Generator().manual_seed(50 + rank)
    active = idle = 0
    for counts in schedule:
        n_steps = train.agreed_count(counts[rank], world)
        assert n_steps == max(counts)
        for k in range(n_steps):
            if k < counts[rank]:
                for p in params:                                     # a "backward pass": separate gradient tensors
                    p.grad = torch.randn(p.shape, generator=gen)
                train.allreduce_gradients(params, world)
                opt.step()
                active += 1
            else:
                train.idle_step(model, params, opt, world)
                idle += 1
    torch.save(dict(w=[p.detach().clone() for p in params], active=active, idle=idle), os.path.join(out_dir, f"u{rank}.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_uneven_minibatch_counts_do_not_hang_or_diverge(tmp_path):
    world = 2
    mp.spawn(_uneven_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"u{i}.pt") for i in range(world))
    assert (r0["active"], r0["idle"]) == (3, 2) and (r1["active"], r1["idle"]) == (3, 2)
    for a, b in zip(r0["w"], r1["w"]):
        assert torch.equal(a, b)
