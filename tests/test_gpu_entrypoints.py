"""The reference's entry points end to end on the GPU: dataset files -> train.run_exp -> mdl.pth ->
pred_basis.run -> HiGHS .bas files, and val.inference_gnn / accuracy on CPU-resident logits like the callers use."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dataset_root(tmp_path_factory):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import dataset
    root = str(tmp_path_factory.mktemp("lpds"))
    dataset.write_synthetic_dataset(root, [(150 + 10 * i, 320 + 20 * i, 1500 + 100 * i) for i in range(8)], seed=2)
    return root


def test_train_then_predict_basis_files(cuda, dataset_root, tmp_path):
    from lpgnn_b200 import arch, dataset, pred_basis, train
    from lpgnn_b200.io_utils import split_train_val
    log_dir = str(tmp_path / "run") + "/"
    args = train.parse_args([], arch="GCN_FC(8,8,hids=64,depth=3)", epochs=3, lr=1e-3, loss="balanced",
                            dataset_processed_prefix=dataset_root, log_dir=log_dir, num_workers=0, log_every=1)
    model, history = train.run_exp(args)
    assert os.path.exists(log_dir + "mdl.pth") and len(history) >= 10
    assert all(np.isfinite(h["loss"]) for h in history)
    assert np.mean([h["loss"] for h in history[-5:]]) < np.mean([h["loss"] for h in history[:5]])
    # prediction sweep with the saved weights
    args2 = train.parse_args([], arch="GCN_FC(8,8,hids=64,depth=3)", load_from=log_dir + "mdl.pth",
                             dataset_processed_prefix=dataset_root, log_dir=log_dir, split="val", num_workers=0)
    times = pred_basis.run(args2)
    ds = dataset.LPDataset(dataset_root, dataset.MyToBipartite(), load_meta=True)
    _, val_ds = split_train_val(ds, 0)
    assert len(times) == len(val_ds) == 3 and all(t > 0 for t in times.values())
    # check every .bas file against the oracle run with the same weights
    ref = port.PortGCN_FC(8, 8, hids=64, depth=3).eval()
    ref.load_state_dict(torch.load(log_dir + "mdl.pth"))
    for i in val_ds.indices():
        uni = ds.get(i)
        fn = os.path.basename(uni.processed_path).replace(".pk", "")
        lines = open(f"{log_dir}/pred-basis/{fn}.bas").read().splitlines()
        g = port.to_bipartite(uni.edge_index.numpy(), uni.edge_attr.numpy(), uni.is_vars.numpy())
        m, n = g.m, g.n
        assert lines[:3] == ["HIGHS v1", "Valid", f"# Columns {n}"] and lines[4] == f"# Rows {m}"
        vbas = np.array(lines[3].split(), dtype=np.int64)
        cbas = np.array(lines[5].split(), dtype=np.int64)
        assert vbas.shape == (n,) and cbas.shape == (m,) and int((vbas == 1).sum() + (cbas == 1).sum()) == m
        with torch.no_grad():
            lc, lv = ref(uni.x[:m], uni.x[m:], port.TorchGraph(g))
        exp = port.inference_gnn_np(torch.cat((lc, lv)).numpy(), m)
        agree = np.mean(np.concatenate([cbas, vbas]) == exp)
        assert agree >= 0.995, agree
        assert os.path.exists(f"{log_dir}/pred-basis/{fn}.bas.sort")
    # val.py entry point (val.py:238-281 -> validation, 43-69) with the same weights: per-LP accuracy vs the oracle's
    from lpgnn_b200 import val
    assert all(0.0 <= h["acc"] <= 1.0 for h in history)                      # train.py:131-137 accuracy meter
    was = model.training
    avg_loss, avg_acc = val.run(args2)
    rows = json.load(open(log_dir + "val_metrics.json"))
    assert avg_loss == 0.0 and len(rows) == 3 and abs(avg_acc - np.mean([r["acc"] for r in rows.values()])) < 1e-12
    for i in val_ds.indices():
        uni = ds.get(i)
        fn = os.path.basename(uni.processed_path).replace(".pk", "")
        g = port.to_bipartite(uni.edge_index.numpy(), uni.edge_attr.numpy(), uni.is_vars.numpy())
        m = g.m
        with torch.no_grad():
            lc, lv = ref(uni.x[:m], uni.x[m:], port.TorchGraph(g))
        exp = port.inference_gnn_np(torch.cat((lc, lv)).numpy(), m)
        y = uni.y.numpy()
        exp_acc = ((exp[:m] == y[:m]).mean() + (exp[m:] == y[m:]).mean()) / 2
        assert abs(rows[fn]["acc"] - exp_acc) < 0.02, (fn, rows[fn], exp_acc)
        assert 0.0 <= rows[fn]["prec"] <= 1.0 and 0.0 <= rows[fn]["recl"] <= 1.0
    # validation() restores the training flag of the model it is given
    model.train()
    val.validation(model, [], torch.device(cuda))
    assert model.training
    model.train(was)
    # sweep mode (block-diagonal packs, segmented basis decision) writes the same .bas files
    log2 = str(tmp_path / "run_packed") + "/"
    args3 = train.parse_args([], arch="GCN_FC(8,8,hids=64,depth=3)", load_from=log_dir + "mdl.pth", packed=1,
                             dataset_processed_prefix=dataset_root, log_dir=log2, split="val", num_workers=0)
    pred_basis.run(args3)
    for i in val_ds.indices():
        fn = os.path.basename(ds.get(i).processed_path).replace(".pk", "")
        assert open(f"{log2}/pred-basis/{fn}.bas").read() == open(f"{log_dir}/pred-basis/{fn}.bas").read()


def test_pred_basis_handles_lps_above_edge_num_thresh(cuda, dataset_root, tmp_path):
    """LPs above ``edge_num_thresh`` stay unipartite in the loader (dataset.py:280-281); the reference routes every batch
    through model_inference_with_batch (scripts/pred_basis.py:79), i.e. the sampled full-neighbourhood path.  Both modes
    of ``pred_basis.run`` must write the same .bas files as with the threshold out of the way."""
    from lpgnn_b200 import arch, pred_basis, train
    torch.manual_seed(3)
    model = arch.GCN_FC(8, 8, hids=64, depth=3)
    ckpt = str(tmp_path / "mdl.pth")
    model.save(ckpt)
    outs = {}
    for tag, extra in (("full", {}), ("above", dict(edge_num_thresh=100.0, batch_size=200)),
                       ("above_packed", dict(edge_num_thresh=100.0, batch_size=200, packed=1))):
        log_dir = str(tmp_path / tag) + "/"
        args = train.parse_args([], arch="GCN_FC(8,8,hids=64,depth=3)", load_from=ckpt, dataset_processed_prefix=dataset_root,
                                log_dir=log_dir, split="val", num_workers=0, **extra)
        times = pred_basis.run(args)
        files = sorted(f for f in os.listdir(log_dir + "pred-basis") if f.endswith(".bas"))
        assert len(files) == 3
        if not extra.get("packed"):
            assert len(times) == 3 and all(t > 0 for t in times.values())
        outs[tag] = {f: np.array(open(f"{log_dir}/pred-basis/{f}").read().splitlines()[3].split() +
                                 open(f"{log_dir}/pred-basis/{f}").read().splitlines()[5].split(), dtype=np.int64) for f in files}
    for f, full in outs["full"].items():
        for tag in ("above", "above_packed"):
            got = outs[tag][f]
            assert got.shape == full.shape and np.mean(got == full) >= 0.995, (tag, f, np.mean(got == full))


def test_packed_training_step_equals_the_average_of_per_lp_steps(cuda, dataset_root):
    """Mini-batches of LP graphs (train.py --pack): the gradient of one packed step (block-diagonal graph, per-LP balanced
    loss, mean over the pack) equals the average of the gradients of the single-LP steps, and train.run_exp trains with it."""
    from lpgnn_b200 import arch, dataset, losses, train
    ds = dataset.LPDataset(dataset_root, dataset.MyToBipartite())
    batches = [ds[i] for i in range(4)]
    torch.manual_seed(5)
    model = arch.GCN_FC(8, 8, hids=64, depth=3, dp=0.0).to(cuda).train()
    params = list(model.parameters())

    def grads_of(loss):
        for p in params:
            p.grad = None
        loss.backward()
        return [p.grad.detach().clone() for p in params]

    singles = []
    for i in range(4):
        b = ds[i].to(cuda)
        lc, lv = model(b)
        singles.append(grads_of(losses.balanced(lc, lv, b.y_s, b.y_t)))
    pack = dataset.pack_bipartite(batches).to(cuda)
    assert pack.n_lps == 4 and pack.cons_ptr.dtype == torch.int32 and int(pack.cons_ptr[-1]) == pack.x_s.shape[0]
    pack.edge_index.check()
    lc, lv = model(pack)
    gp = grads_of(losses.balanced_packed(lc, lv, pack.y_s, pack.y_t, pack.cons_ptr, pack.vars_ptr))
    for k, g in enumerate(gp):
        mean = sum(s[k] for s in singles) / 4
        assert float((g - mean).abs().max()) <= 2e-4 * max(1e-6, float(mean.abs().max())), k


def test_train_entry_point_with_packs(cuda, dataset_root, tmp_path):
    from lpgnn_b200 import train
    log_dir = str(tmp_path / "run_pack") + "/"
    args = train.parse_args([], arch="GCN_FC(8,8,hids=64,depth=3)", epochs=6, lr=2e-3, loss="balanced", pack=2,
                            dataset_processed_prefix=dataset_root, log_dir=log_dir, num_workers=0, log_every=1)
    model, history = train.run_exp(args)
    assert os.path.exists(log_dir + "mdl.pth") and len(history) >= 12            # 5 training LPs -> 3 packs per epoch
    assert all(np.isfinite(h["loss"]) and 0.0 <= h["acc"] <= 1.0 for h in history)
    assert np.mean([h["loss"] for h in history[-4:]]) < np.mean([h["loss"] for h in history[:4]])


def test_val_inference_gnn_and_accuracy_accept_cpu_logits(cuda):
    """pred_basis.py:81-85 and train.py:132-137 hand CPU / detached logits to inference_gnn / accuracy."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import val
    rng = np.random.default_rng(0)
    m, n = 300, 700
    logits = torch.from_numpy((rng.standard_normal((m + n, 3)) * 3).astype(np.float32))
    pred = val.inference_gnn(logits, m)
    assert pred.device.type == "cpu" and pred.dtype == torch.int64
    np.testing.assert_array_equal(pred.numpy(), port.inference_gnn_np(logits.numpy(), m))
    gt = torch.from_numpy(rng.integers(0, 3, m + n))
    acc, prec, recl = val.accuracy(logits, gt, m, return_pr=True)
    exp = port.inference_gnn_np(logits.numpy(), m)
    exp_acc = ((exp[:m] == gt.numpy()[:m]).mean() + (exp[m:] == gt.numpy()[m:]).mean()) / 2
    assert abs(acc - exp_acc) < 1e-12 and 0 <= prec <= 1 and 0 <= recl <= 1


def test_model_inference_with_batch_and_half_switch(cuda, dataset_root):
    from lpgnn_b200 import arch, dataset, val
    ds = dataset.LPDataset(dataset_root, dataset.MyToBipartite())
    batch = ds[0]
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=128, depth=3).to(cuda)
    lc, lv = val.model_inference_with_batch(model, batch, types.SimpleNamespace(fp16=0, arch="GCN_FC(8,8,hids=128,depth=3)"))
    assert lc.device.type == "cpu" and lc.shape == (batch.s_bs, 3) and lv.shape == (batch.t_bs, 3)
    batch.edge_index.check()
    model.half()
    lc2, lv2 = val.model_inference_with_batch(model, ds[0], types.SimpleNamespace(fp16=1))
    assert float((lc2 - lc).norm() / lc.norm()) < 2e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_data_parallel_training_matches_single_process(cuda, dataset_root, tmp_path):
    """torchrun-style 2-rank NCCL training: replicas stay identical and the loss goes down."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    log_dir = str(tmp_path / "ddp") + "/"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "ddp_train_entry.py"), dataset_root, log_dir]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert os.path.exists(log_dir + "mdl.pth")
    import json
    log = json.load(open(log_dir + "train_log.json"))
    assert log["world"] == 2 and log["history"][-1]["loss"] < log["history"][0]["loss"]


def test_basis_pipeline_matches_direct_calls(cuda):
    """Host-buffer pipeline (packed H2D on a side stream, uint8 D2H) == direct per-LP calls, for ragged sizes."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch, synth
    from lpgnn_b200.graph import BipartiteCSR
    from lpgnn_b200.pipeline import BasisPipeline, pack_lp
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=128, depth=3).to(cuda).eval().set_precision("bf16")
    sizes = [(300, 700, 3000), (50, 90, 300), (1200, 2000, 9000), (5, 9, 20), (700, 1500, 6000)]
    lps = [synth.processed_lp(m, n, z, seed=40 + i) for i, (m, n, z) in enumerate(sizes)]
    hosts = [pack_lp(lp.row, lp.col, lp.a_data, lp.c_feas, lp.v_feas) for lp in lps]
    pipe = BasisPipeline(model, cuda)
    got = {i: st.copy() for i, st in pipe.run(hosts)}
    assert sorted(got) == list(range(len(lps)))
    for i, lp in enumerate(lps):
        g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, cuda, is_sorted=True)
        batch = types.SimpleNamespace(x_s=torch.from_numpy(lp.c_feas).to(cuda), x_t=torch.from_numpy(lp.v_feas).to(cuda),
                                      edge_index=g)
        exp = model.predict_basis(batch, int64=False).cpu().numpy()
        np.testing.assert_array_equal(got[i], exp)
        assert int((got[i] == 1).sum()) == lp.m
    # a second pass over the same pipeline object reuses the slots
    again = {i: st.copy() for i, st in pipe.run(hosts[::-1])}
    np.testing.assert_array_equal(again[0], got[len(lps) - 1])
    # any number of LPs in flight (1 = the caller's stream), including more slots than LPs: same statuses, in order
    for k in (1, 2, 4, 8):
        p2 = BasisPipeline(model, cuda, compute_streams=k)
        order = []
        for i, st in p2.run(hosts):
            order.append(i)
            np.testing.assert_array_equal(st, got[i])
        assert order == list(range(len(lps)))
        assert [i for i, _ in p2.run(hosts[:2])] == [0, 1]


def test_packed_pipeline_matches_per_lp_prediction(cuda):
    """Block-diagonal packs (one forward per pack, segmented basis decision) == one call per LP."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch, synth
    from lpgnn_b200.pipeline import BasisPipeline, PackedBasisPipeline, pack_lp
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=128, depth=3).to(cuda).eval().set_precision("fp32")
    rng = np.random.default_rng(5)
    sizes = [(int(m), int(2 * m), int(10 * m)) for m in rng.integers(20, 3000, 37)] + [(5, 9, 20), (4000, 8000, 40_000)]
    lps = [synth.processed_lp(m, n, z, seed=70 + i) for i, (m, n, z) in enumerate(sizes)]
    hosts = [pack_lp(lp.row, lp.col, lp.a_data, lp.c_feas, lp.v_feas) for lp in lps]
    single = {i: st.copy() for i, st in BasisPipeline(model, cuda).run(hosts)}
    packed = PackedBasisPipeline(model, cuda, max_nodes=20_000, max_nnz=100_000, max_lps=8)
    plan = packed._plan(hosts)
    assert len(plan) > 4 and max(len(p) for p in plan) > 1 and sorted(sum(plan, [])) == list(range(len(lps)))
    got = dict(packed.run(hosts))
    assert sorted(got) == list(range(len(lps)))
    for i, lp in enumerate(lps):
        assert got[i].shape == (lp.m + lp.n,) and int((got[i] == 1).sum()) == lp.m      # per-LP top-m rule
        np.testing.assert_array_equal(got[i], single[i])                                # fp32: same arithmetic per row


def test_segmented_select_matches_single(cuda):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, ops
    rng = np.random.default_rng(9)
    ms = [1, 7, 300, 2500, 40]
    ns = [1, 20, 700, 4100, 90]
    lc = torch.from_numpy((rng.standard_normal((sum(ms), 3)) * 3).astype(np.float32)).to(cuda)
    lv = torch.from_numpy((rng.standard_normal((sum(ns), 3)) * 3).astype(np.float32)).to(cuda)
    lv[5:25] = 0.0                                                  # ties inside segment 1
    cptr = torch.tensor(np.concatenate([[0], np.cumsum(ms)]), dtype=torch.int32, device=cuda)
    vptr = torch.tensor(np.concatenate([[0], np.cumsum(ns)]), dtype=torch.int32, device=cuda)
    lib = _lib.load()
    M, N = sum(ms), sum(ns)
    status = torch.empty(M + N, dtype=torch.uint8, device=cuda)
    nb = lib.lpgnn_basis_select_workspace_bytes(M + N)
    ws = torch.empty(nb, dtype=torch.uint8, device=cuda)
    rc = lib.lpgnn_basis_select_segmented(lc.data_ptr(), lv.data_ptr(), cptr.data_ptr(), vptr.data_ptr(), len(ms), M, N,
                                          status.data_ptr(), 0, ws.data_ptr(), nb, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    status = status.cpu().numpy()
    c, v = cptr.cpu().numpy(), vptr.cpu().numpy()
    for b in range(len(ms)):
        exp = ops.basis_select(lc[c[b]:c[b + 1]], lv[v[b]:v[b + 1]], int64=False).cpu().numpy()
        got = np.concatenate([status[c[b]:c[b + 1]], status[M + v[b]:M + v[b + 1]]])
        np.testing.assert_array_equal(got, exp)
