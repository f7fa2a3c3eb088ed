"""(f-3) Fused training losses and device-side accuracy counters vs the framework spellings of the reference's
train.py:18-53 / val.py:199-237 (those are checked against the verbatim reference on CPU in
test_oracle_vs_reference.py::test_training_losses_match_reference_train_py)."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _case(m, n, seed, cuda, classes_s=(0, 1, 2), classes_t=(0, 1, 2)):
    rng = np.random.default_rng(seed)
    lc = torch.from_numpy((rng.standard_normal((m, 3)) * 4).astype(np.float32))
    lv = torch.from_numpy((rng.standard_normal((n, 3)) * 4).astype(np.float32))
    ys = torch.from_numpy(rng.choice(classes_s, m)) if m else torch.zeros(0, dtype=torch.int64)
    yt = torch.from_numpy(rng.choice(classes_t, n)) if n else torch.zeros(0, dtype=torch.int64)
    return lc, lv, ys.long(), yt.long()


@pytest.mark.parametrize("name", ["balanced", "unbalanced", "focal"])
@pytest.mark.parametrize("m,n,cs,ct", [(7, 9, (0, 1, 2), (0, 1, 2)), (300, 700, (0, 1, 2), (1, 2)), (1000, 2000, (1, 2), (0, 1)),
                                       (50_000, 100_000, (0, 1, 2), (0, 1, 2)), (257, 1, (0, 1), (1,))])
def test_fused_losses_match_framework_ops(cuda, name, m, n, cs, ct):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import losses
    lc, lv, ys, yt = _case(m, n, m + n, cuda, cs, ct)
    ref_fn = {"balanced": losses.balanced_torch, "unbalanced": losses.unbalanced_torch, "focal": losses.focal_torch}[name]

    def run(fn, dev, scale):
        a, b = lc.detach().clone().to(dev).requires_grad_(), lv.detach().clone().to(dev).requires_grad_()
        loss = fn(a, b, ys.to(dev), yt.to(dev))
        (loss * scale).backward()                    # a non-trivial upstream gradient
        return float(loss.detach()), a.grad.cpu(), b.grad.cpu()

    ref = run(ref_fn, "cpu", 0.7)
    got = run(losses.LOSSES[name], cuda, 0.7)
    assert abs(ref[0] - got[0]) <= 2e-6 * max(1.0, abs(ref[0])), (name, ref[0], got[0])
    for r, g in ((ref[1], got[1]), (ref[2], got[2])):
        assert torch.allclose(r, g, rtol=2e-5, atol=1e-9 + 2e-6 * float(r.abs().max())), name
    # bit-reproducible (fixed summation order, integer atomics only)
    again = run(losses.LOSSES[name], cuda, 0.7)
    assert got[0] == again[0] and torch.equal(got[1], again[1]) and torch.equal(got[2], again[2])


def test_fused_losses_take_the_native_path_without_host_sync(cuda):
    """The three losses are native calls on CUDA tensors: the launch counter moves, and nothing reads back to the host
    (torch.cuda.set_sync_debug_mode('error') would raise on a sync)."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, losses
    lc, lv, ys, yt = (t.to(cuda) for t in _case(500, 900, 3, cuda))
    lib = _lib.load()
    for name in ("balanced", "unbalanced", "focal"):
        a, b = lc.clone().requires_grad_(), lv.clone().requires_grad_()
        torch.cuda.synchronize()
        before = lib.lpgnn_launch_count()
        torch.cuda.set_sync_debug_mode("error")
        try:
            loss = losses.LOSSES[name](a, b, ys, yt)
            loss.backward()
        finally:
            torch.cuda.set_sync_debug_mode("default")
        assert lib.lpgnn_launch_count() > before, name
        assert bool(torch.isfinite(loss))


@pytest.mark.parametrize("m,n,seed", [(300, 700, 0), (1, 5, 1), (50_000, 100_000, 2), (64, 64, 3)])
@pytest.mark.parametrize("stoch", [False, True])
def test_accuracy_counters_match_sklearn_on_the_reference_formula(cuda, m, n, seed, stoch):
    """val.accuracy (reference val.py:199-237) from twelve device integers == the reference's numpy + sklearn spelling on
    the oracle's basis decision."""
    from sklearn import metrics
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import val
    rng = np.random.default_rng(seed)
    logits = torch.from_numpy((rng.standard_normal((m + n, 3)) * 3).astype(np.float32))
    gt = torch.from_numpy(rng.integers(0, 3, m + n))
    name = "stoch-x" if stoch else "mirp"
    acc, prec, recl = val.accuracy(logits.to(cuda), gt.to(cuda), m, return_pr=True, dataset_name=name)
    pred = port.inference_gnn_np(logits.numpy(), m)
    g = gt.numpy()
    a1, a2 = (g[:m] == pred[:m]).mean(), (g[m:] == pred[m:]).mean()
    kw = dict(labels=[1], average="macro", zero_division=0)
    p1, p2 = metrics.precision_score(g[:m], pred[:m], **kw), metrics.precision_score(g[m:], pred[m:], **kw)
    r1, r2 = metrics.recall_score(g[:m], pred[:m], **kw), metrics.recall_score(g[m:], pred[m:], **kw)
    if stoch:
        a1, p1, r1 = a2, p2, r2
    assert abs(acc - (a1 + a2) / 2) < 1e-12 and abs(prec - (p1 + p2) / 2) < 1e-12 and abs(recl - (r1 + r2) / 2) < 1e-12
    assert val.accuracy(logits, gt, m, dataset_name=name) == acc                      # CPU-resident inputs, scalar form
    counts = val.accuracy_counts(logits.to(cuda), gt.to(cuda), m)
    assert counts.is_cuda and counts.dtype == torch.int32 and counts.shape == (12,)


@pytest.mark.parametrize("sizes", [[(7, 9)], [(300, 700), (1, 5), (257, 64)], [(1000, 2000)] * 5 + [(64, 3)]])
def test_packed_balanced_loss_is_the_mean_of_the_per_lp_losses(cuda, sizes):
    """losses.balanced_packed (one CTA per LP and side, per-LP class weights) vs the framework spelling: value, gradients,
    bit-reproducibility."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import losses
    parts = [_case(m, n, 10 * i + m, cuda, (0, 1, 2) if i % 2 else (1, 2), (0, 1, 2)) for i, (m, n) in enumerate(sizes)]
    lc, lv = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    ys, yt = torch.cat([p[2] for p in parts]), torch.cat([p[3] for p in parts])
    cptr = torch.tensor(np.concatenate([[0], np.cumsum([m for m, _ in sizes])]), dtype=torch.int32)
    vptr = torch.tensor(np.concatenate([[0], np.cumsum([n for _, n in sizes])]), dtype=torch.int32)

    def run(dev):
        a, b = lc.clone().to(dev).requires_grad_(), lv.clone().to(dev).requires_grad_()
        loss = losses.balanced_packed(a, b, ys.to(dev), yt.to(dev), cptr.to(dev), vptr.to(dev))
        (loss * 1.3).backward()
        return float(loss.detach()), a.grad.cpu(), b.grad.cpu()

    ref, got, again = run("cpu"), run(cuda), run(cuda)
    assert abs(ref[0] - got[0]) <= 2e-6 * max(1.0, abs(ref[0])), (ref[0], got[0])
    for r, g in ((ref[1], got[1]), (ref[2], got[2])):
        assert torch.allclose(r, g, rtol=2e-5, atol=1e-9 + 2e-6 * float(r.abs().max()))
    assert got[0] == again[0] and torch.equal(got[1], again[1]) and torch.equal(got[2], again[2])
