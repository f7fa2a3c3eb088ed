"""Training step parity: gradients of every parameter through the CUDA backward kernels vs the gradients
the REFERENCE computed (golden vectors produced by /root/reference's own autograd, oracle/make_golden.py),
plus unit checks of each backward kernel against torch autograd on the same inputs."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _golden_setup(name, dev, precision):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch
    from lpgnn_b200.graph import BipartiteCSR
    zg = np.load(os.path.join(GOLD, f"graph_{name}.npz"))
    zm = np.load(os.path.join(GOLD, f"model_{name}.npz"))
    model = arch.GCN_FC(8, 8, hids=int(zm["hids"]), depth=int(zm["depth"]))
    model.load_state_dict({k[3:]: torch.from_numpy(zm[k]) for k in zm.files if k.startswith("w::")})
    model = model.to(dev).eval().set_precision(precision)          # eval: the golden gradients were taken without dropout
    m, n = int(zg["m"]), int(zg["n"])
    row = np.repeat(np.arange(m), np.diff(zg["rowptr"]))
    g = BipartiteCSR.from_coo_arrays(row, zg["col"], zg["val"], m, n, dev, is_sorted=True).check()
    batch = types.SimpleNamespace(x_s=torch.from_numpy(zg["x_s"]).to(dev), x_t=torch.from_numpy(zg["x_t"]).to(dev),
                                  edge_index=g)
    y_s, y_t = torch.from_numpy(zg["y_s"]).to(dev), torch.from_numpy(zg["y_t"]).to(dev)
    return model, batch, y_s, y_t, zm


@pytest.mark.parametrize("name", ["tiny_5x7", "small_300x600", "c1_1000x2000"])
def test_loss_and_gradients_match_reference_fp32(cuda, name):
    from lpgnn_b200.losses import balanced
    model, batch, y_s, y_t, zm = _golden_setup(name, cuda, "fp32")
    lc, lv = model(batch)                                           # grad enabled -> autograd.Function path
    assert lc.requires_grad
    np.testing.assert_allclose(lc.detach().cpu().numpy(), zm["logits_cons"], atol=1e-3)
    assert np.abs(lv.detach().cpu().numpy() - zm["logits_vars"]).max() / 10 < 1e-4
    loss = balanced(lc, lv, y_s, y_t)
    assert abs(float(loss) - float(zm["loss"])) < 1e-4 * max(1.0, abs(float(zm["loss"])))
    loss.backward()
    for k, p in model.named_parameters():
        ref = zm[f"g::{k}"]
        got = p.grad.cpu().numpy()
        assert got.shape == ref.shape, k
        scale = max(np.abs(ref).max(), 1e-6)
        assert np.abs(got - ref).max() / scale < 2e-3, (k, np.abs(got - ref).max() / scale)


@pytest.mark.parametrize("name", ["small_300x600", "c1_1000x2000"])
def test_gradients_bf16_close_to_reference(cuda, name):
    from lpgnn_b200.losses import balanced
    model, batch, y_s, y_t, zm = _golden_setup(name, cuda, "bf16")
    lc, lv = model(batch)
    loss = balanced(lc, lv, y_s, y_t)
    assert abs(float(loss) - float(zm["loss"])) < 2e-2 * max(1.0, abs(float(zm["loss"])))
    loss.backward()
    for k, p in model.named_parameters():
        ref = zm[f"g::{k}"]
        got = p.grad.cpu().numpy()
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-9)
        assert rel < 8e-2, (k, rel)                                 # bf16 activations + bf16 tensor-core GEMMs


def test_backward_is_bit_reproducible(cuda):
    from lpgnn_b200.losses import balanced
    grads = []
    for _ in range(2):
        model, batch, y_s, y_t, _ = _golden_setup("small_300x600", cuda, "fp32")
        lc, lv = model(batch)
        balanced(lc, lv, y_s, y_t).backward()
        grads.append([p.grad.clone() for p in model.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b)                                    # atomics-free: identical bits run to run


def test_training_reduces_loss_with_dropout(cuda):
    from lpgnn_b200.losses import balanced
    model, batch, y_s, y_t, _ = _golden_setup("small_300x600", cuda, "fp32")
    model.train()
    model.dp = 0.1
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    losses = []
    for _ in range(30):
        lc, lv = model(batch)
        loss = balanced(lc, lv, y_s, y_t)
        assert not torch.isnan(loss).item()                          # train.py:126
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert min(losses[-5:]) < losses[0]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,dp", [("small_300x600", 0.0), ("c1_1000x2000", 0.2)])
def test_native_step_equals_op_by_op_step(cuda, precision, name, dp):
    """lpgnn_train_forward/backward enqueue the same kernels as the op-by-op orchestration: identical bits."""
    from lpgnn_b200 import training
    from lpgnn_b200.losses import balanced
    out = []
    for native in (True, False):
        model, batch, y_s, y_t, _ = _golden_setup(name, cuda, precision)
        model.train()
        model.dp, model.native_train = dp, native
        training._step_counter[0] = 41                               # same dropout seed for both paths
        lc, lv = model(batch)
        assert (type(lc.grad_fn).__name__.startswith("_NativeTrainFunction")) == native
        loss = balanced(lc, lv, y_s, y_t)
        loss.backward()
        out.append((lc.detach().clone(), lv.detach().clone(), [p.grad.clone() for p in model.parameters()]))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    for (k, _), a, b in zip(model.named_parameters(), out[0][2], out[1][2]):
        assert torch.equal(a, b), (k, (a - b).abs().max().item())


@pytest.mark.parametrize("precision,name", [("fp32", "small_300x600"), ("bf16", "c1_1000x2000")])
def test_two_phase_backward_with_gradient_sync_hook(cuda, precision, name):
    """Data-parallel form of the native backward (TAIL, hook, REST, hook): same bits as the one-call form; the hook sees
    the tail of the flat gradient buffer (last hidden layer + head) first, then the head of it, and the scale is applied."""
    from lpgnn_b200 import training
    from lpgnn_b200.losses import balanced
    from lpgnn_b200.train import allreduce_gradients
    res, seen = [], []

    class _H:
        def wait(self):
            seen.append("wait")

    def hook(t):
        seen.append((t.data_ptr(), t.numel()))
        return _H()

    for sync in (False, True):
        model, batch, y_s, y_t, _ = _golden_setup(name, cuda, precision)
        model.train()
        model.dp = 0.1
        training._step_counter[0] = 7
        training.set_gradient_sync(hook if sync else None, 0.5)
        try:
            lc, lv = model(batch)
            balanced(lc, lv, y_s, y_t).backward()
        finally:
            training.set_gradient_sync(None)
        res.append([p.grad.clone() for p in model.parameters()])
    params = list(model.parameters())
    for (k, _), a, b in zip(model.named_parameters(), res[0], res[1]):
        assert torch.equal(a * 0.5, b), k
    n_tail = sum(p.numel() for p in params[-(10 if len(params) > 10 else 4):])   # last hidden layer (6 tensors) + head (4)
    (p1, c1), (p0, c0) = seen[0], seen[1]
    assert seen[2:] == ["wait", "wait"]
    assert c1 >= n_tail and c1 < n_tail + 4 * 10 and p0 + 4 * c0 == p1      # tail first; the two slices tile the buffer
    # the training loop's all-reduce skips a backward pass that was reduced by the hook (flag consumed once)
    assert training.consume_synced_backward() and not training.consume_synced_backward()
    allreduce_gradients(params, 1)


def test_native_step_rejects_second_backward(cuda):
    from lpgnn_b200.losses import balanced
    model, batch, y_s, y_t, _ = _golden_setup("small_300x600", cuda, "fp32")
    lc, lv = model(batch)
    loss = balanced(lc, lv, y_s, y_t)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        loss.backward()


# ------------------------------------------------------------------------------------------- kernel units
@pytest.mark.parametrize("classes_s,classes_t", [((0, 1, 2), (0, 1, 2)), ((0, 1), (1, 2)), ((1,), (0, 1, 2)), ((0, 2), (0,)),
                                                 ((0, 1, 2), (1,))])
@pytest.mark.parametrize("m,n", [(5, 7), (1000, 2000), (50_000, 100_001)])
def test_fused_balanced_ce_matches_framework_ops(cuda, classes_s, classes_t, m, n):
    """lpgnn_balanced_ce == reference balanced() (train.py:39-46) spelled with torch ops, value and gradient, for every
    class-presence pattern of labels_to_balanced_weights (utils.py:286-299)."""
    from lpgnn_b200 import losses
    gen = torch.Generator(device="cpu").manual_seed(m * 31 + n)
    ys = torch.tensor(classes_s)[torch.randint(0, len(classes_s), (m,), generator=gen)]
    yt = torch.tensor(classes_t)[torch.randint(0, len(classes_t), (n,), generator=gen)]
    ys[:len(classes_s)] = torch.tensor(classes_s)                     # every listed class occurs
    yt[:len(classes_t)] = torch.tensor(classes_t)
    ls = (torch.randn(m, 3, generator=gen) * 6).to(cuda).requires_grad_()
    lt = (torch.randn(n, 3, generator=gen) * 6).to(cuda).requires_grad_()
    ys, yt = ys.to(cuda), yt.to(cuda)
    ref = losses.balanced_torch(ls, lt, ys, yt)
    gs_ref, gt_ref = torch.autograd.grad(ref, (ls, lt))
    got = losses.balanced(ls, lt, ys, yt)
    assert type(got.grad_fn).__name__.startswith("_BalancedCE")
    gs, gt = torch.autograd.grad(got * 1.0, (ls, lt))
    assert abs(got.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item()))
    for a, b in ((gs, gs_ref), (gt, gt_ref)):
        assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-12) + 1e-12
    got2 = losses.balanced(ls, lt, ys, yt)
    assert torch.equal(got, got2)                                    # fixed summation order


@pytest.mark.parametrize("M,N,K", [(1000, 256, 128), (4097, 1024, 1024), (300, 64, 64)])
def test_transform_epilogue_dropout_and_mask_bf16(cuda, M, N, K):
    """Fused keep-masks of the tensor-core transform == fp32-output transform + torch scale/mask, rounded once; the
    dropout pattern is exactly lpgnn_dropout's."""
    from lpgnn_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(M + N)
    bf = torch.bfloat16
    a1, a2 = torch.randn(M, K, generator=gen).to(cuda).to(bf), torch.randn(M, K, generator=gen).to(cuda).to(bf)
    w1, w2 = (torch.randn(N, K, generator=gen) / K ** 0.5).to(cuda).to(bf), (torch.randn(N, K, generator=gen) / K ** 0.5).to(cuda).to(bf)
    bias = torch.randn(N, generator=gen).to(cuda)
    ref32 = ops.node_transform(a1, w1, a2, w2, bias, relu=True, out_dtype=torch.float32)
    p, seed = 0.25, 123456789
    keep = torch.ones(M, N, device=cuda, dtype=bf)
    ops.dropout_(keep, p, seed)
    keep = keep > 0
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    got = ops.node_transform(a1, w1, a2, w2, bias, relu=True, dropout=(p, seed))
    want = (ref32 * torch.tensor(1.0 / (1.0 - p), dtype=torch.float32, device=cuda)).to(bf) * keep
    assert torch.equal(got, want)
    # backward-style mask: no bias, no relu, scale * (act > 0)
    act = torch.relu(torch.randn(M, N, generator=gen)).to(cuda).to(bf)
    ref32 = ops.node_transform(a1, w1, a2, w2, out_dtype=torch.float32)
    got = ops.node_transform(a1, w1, a2, w2, mask=(act, 1.25))
    want = (ref32 * 1.25).to(bf) * (act > 0)
    assert torch.equal(got, want)
    # both at once
    got = ops.node_transform(a1, w1, None, None, bias, relu=True, dropout=(p, seed), mask=(act, 2.0))
    ref32 = ops.node_transform(a1, w1, None, None, bias, relu=True, out_dtype=torch.float32)
    want = (ref32 * (torch.tensor(2.0, device=cuda) * torch.tensor(1.0 / (1.0 - p), dtype=torch.float32, device=cuda))).to(bf) * keep * (act > 0)
    assert torch.equal(got, want)


def test_transform_epilogue_fp32_matches_separate_kernels(cuda):
    from lpgnn_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 777, 64, 64
    a, w = torch.randn(M, K, generator=gen).to(cuda), torch.randn(N, K, generator=gen).to(cuda)
    act = torch.relu(torch.randn(M, N, generator=gen)).to(cuda)
    got = ops.node_transform(a, w, relu=True, dropout=(0.5, 77), mask=(act, 1.5))
    ref = ops.node_transform(a, w, relu=True)
    ref = ops.relu_bwd(ref, None, act, 1.5)
    ops.dropout_(ref, 0.5, 77)
    assert torch.equal(got, ref)

def test_bf16_operands_of_the_narrow_weight_gradients(cuda):
    """gather_cat's bf16 output carries a ones column after the features (bias gradient through the tensor-core
    weight gradient); head_mask_bwd's bf16 draw is [draw | 0]; wgrad over them matches fp32 matmuls."""
    from conftest import make_graph_arrays
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    m, n, H = 700, 1300, 128
    row, col, val = make_graph_arrays(m, n, 6000, seed=3)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda)
    csr, csc = g.views()
    gen = torch.Generator(device="cpu").manual_seed(0)
    x_s, x_t = torch.randn(m, 8, generator=gen).to(cuda), torch.randn(n, 8, generator=gen).to(cuda)
    z32, zb = ops.gather_cat(csc, x_s, x_t, want_f32=True, want_bf16=True)
    assert torch.equal(zb[:, :16], z32[:, :16].to(torch.bfloat16))
    assert torch.all(zb[:, 16] == 1) and torch.all(zb[:, 17:] == 0)
    d_pre = (torch.randn(n, H, generator=gen) * 0.1).to(cuda).to(torch.bfloat16)
    gw = ops.wgrad(d_pre, zb)                                        # [H,64]
    ref = d_pre.float().t() @ zb.float()
    assert (gw - ref).abs().max() / ref.abs().max() < 1e-4
    assert (gw[:, 16] - d_pre.float().sum(0)).abs().max() < 1e-3 * d_pre.float().abs().sum(0).max()
    assert torch.all(gw[:, 17:] == 0)
    # head: draw as a bf16 operand
    h_act = torch.relu(torch.randn(n, H, generator=gen)).to(cuda).to(torch.bfloat16)
    w = torch.randn(3, H, generator=gen).to(cuda)
    raw, dlog = torch.randn(n, 3, generator=gen).to(cuda), torch.randn(n, 3, generator=gen).to(cuda)
    dH, draw, draw_b = ops.head_mask_bwd(dlog, raw, h_act, w, 1.0, want_bf16=True)
    dH2, draw2 = ops.head_mask_bwd(dlog, raw, h_act, w, 1.0)
    assert torch.equal(dH, dH2) and torch.equal(draw, draw2)
    assert torch.equal(draw_b[:, :3], draw.to(torch.bfloat16)) and torch.all(draw_b[:, 3:] == 0)
    gh = ops.wgrad(h_act, draw_b)[:, :3]
    ref = h_act.float().t() @ draw_b[:, :3].float()
    assert (gh - ref).abs().max() / ref.abs().max() < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_relu_bwd_and_transpose_and_colsum(cuda, dtype):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N = 1000, 192
    a = torch.randn(M, N, device=cuda, generator=g).to(dtype)
    b = torch.randn(M, N, device=cuda, generator=g).to(dtype)
    act = torch.randn(M, N, device=cuda, generator=g).relu().to(dtype)
    out = ops.relu_bwd(a, b, act, scale=1.25)
    exp = ((a.float() + b.float()) * 1.25 * (act > 0)).to(dtype)
    assert torch.equal(out, exp)
    assert torch.equal(ops.relu_bwd(a, None, act), (a.float() * (act > 0)).to(dtype))
    t = ops.transpose(a)
    assert t.shape == (N, 1024) and torch.equal(t[:, :M], a.t()) and not t[:, M:].any()
    cs = ops.colsum(a)
    assert float((cs - a.float().sum(0)).abs().max()) < 1e-2
    assert torch.equal(cs, ops.colsum(a))                           # deterministic


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,ldz", [(3, 3), (16, 16), (12, 16), (40, 64)])
def test_small_wgrad(cuda, dtype, K, ldz):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(K)
    M, N = 3000, 320
    dy = torch.randn(M, N, device=cuda, generator=g).to(dtype)
    z = torch.randn(M, ldz, device=cuda, generator=g)
    dW, dB = ops.small_wgrad(dy, z, K, want_bias=True)
    exp = dy.double().t() @ z[:, :K].double()
    assert float((dW.double() - exp).abs().max()) < 2e-3
    assert float((dB.double() - dy.double().sum(0)).abs().max()) < 2e-3


@pytest.mark.parametrize("dtype,H", [(torch.float32, 64), (torch.float32, 1024), (torch.bfloat16, 1024), (torch.bfloat16, 128)])
def test_head_mask_bwd_vs_autograd(cuda, dtype, H):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(H)
    rows = 2000
    h = torch.randn(rows, H, device=cuda, generator=g).relu().to(dtype)
    w = torch.randn(3, H, device=cuda, generator=g) / H ** 0.5
    b = torch.randn(3, device=cuda, generator=g)
    feas = torch.randint(-1, 2, (rows, 8), device=cuda, generator=g).float()
    dl = torch.randn(rows, 3, device=cuda, generator=g)
    logits, raw = ops.head_mask(h, w, b, feas, want_raw=True)
    dH, draw = ops.head_mask_bwd(dl, raw, h, w, scale=1.0)
    hh = h.float().clone().requires_grad_(True)
    ww = w.clone().requires_grad_(True)
    r = hh @ ww.t() + b
    y = torch.nn.functional.normalize(r) * 10
    (y * dl).sum().backward()
    exp_dH = hh.grad * (h.float() > 0)
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    assert float((dH.float() - exp_dH).abs().max()) < tol * max(1.0, float(exp_dH.abs().max()))
    gw, _ = ops.small_wgrad(h, draw, 3)
    assert float((gw.t() - ww.grad).abs().max()) < 1e-2 * max(1.0, float(ww.grad.abs().max()))


@pytest.mark.parametrize("dtype,H,rows", [(torch.float32, 64, 2000), (torch.float32, 1024, 777), (torch.bfloat16, 1024, 30011),
                                          (torch.bfloat16, 128, 3), (torch.bfloat16, 2048, 1500)])
def test_head_mask_bwd_fused_column_sums(cuda, dtype, H, rows):
    """head_mask_bwd(want_colsum) = same dH / draw bit for bit + colsum(dH) taken in fp32 before the rounding
    (bias gradient of the layer under the head, GraphConv.lin_rel.bias); deterministic."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(H + rows)
    h = torch.randn(rows, H, device=cuda, generator=g).relu().to(dtype)
    w = torch.randn(3, H, device=cuda, generator=g) / H ** 0.5
    b = torch.randn(3, device=cuda, generator=g)
    feas = torch.randint(-1, 2, (rows, 8), device=cuda, generator=g).float()
    dl = torch.randn(rows, 3, device=cuda, generator=g)
    _, raw = ops.head_mask(h, w, b, feas, want_raw=True)
    dH, draw = ops.head_mask_bwd(dl, raw, h, w, scale=1.25)
    dH2, draw2, cs = ops.head_mask_bwd(dl, raw, h, w, scale=1.25, want_colsum=True)
    assert torch.equal(dH, dH2) and torch.equal(draw, draw2)
    # exact (unrounded) dH from the kernel's own draw, in float64
    exact = (draw.double() @ w.double()) * 1.25 * (h.double() > 0)
    exp = exact.sum(0)
    assert float((cs.double() - exp).abs().max()) < 1e-4 * max(1.0, float(exact.abs().sum(0).max()))
    _, _, cs2 = ops.head_mask_bwd(dl, raw, h, w, scale=1.25, want_colsum=True)
    assert torch.equal(cs, cs2)
    # the head's own bias gradient colsum(draw) from the same pass, with or without the dH sums
    dH3, draw3, cs3, db3 = ops.head_mask_bwd(dl, raw, h, w, scale=1.25, want_colsum=True, want_bias_grad=True)
    _, _, db4 = ops.head_mask_bwd(dl, raw, h, w, scale=1.25, want_bias_grad=True)
    assert torch.equal(dH3, dH) and torch.equal(cs3, cs) and torch.equal(db3, db4) and db3.shape == (3,)
    expd = draw.double().sum(0)
    assert float((db3.double() - expd).abs().max()) < 1e-5 * max(1.0, float(draw.double().abs().sum(0).max()))
    # canary: the partial-sum workspace is the only scratch; the [H] output has no neighbours overwritten
    assert cs.shape == (H,) and torch.isfinite(cs).all()


def test_dropout_statistics_and_determinism(cuda):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    x = torch.ones(4096, 1024, device=cuda)
    a = ops.dropout_(x.clone(), 0.1, seed=123)
    b = ops.dropout_(x.clone(), 0.1, seed=123)
    c = ops.dropout_(x.clone(), 0.1, seed=124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    keep = float((a > 0).float().mean())
    assert abs(keep - 0.9) < 2e-3
    assert abs(float(a.mean()) - 1.0) < 3e-3                        # inverted scaling keeps the mean
    assert set(torch.unique(a).cpu().tolist()) == {0.0, float(torch.tensor(1.0 / 0.9, dtype=torch.float32))}


@pytest.mark.parametrize("M,N,K", [(1024, 2048, 100_032), (64, 128, 1024), (128, 256, 64 * 9), (1024, 1024, 50_048)])
def test_gemm_tn_splitk(cuda, M, N, K):
    """Split-K tensor-core GEMM (weight gradients): exact bf16 products, fp32 accumulate, deterministic."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(K % 1000)
    a = torch.randn(M, K, device=cuda, generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, device=cuda, generator=g).to(torch.bfloat16)
    out = ops.gemm_tn(a, b)
    splits = _lib.load().lpgnn_gemm_tn_splits(M, N, K)
    assert splits >= 1 and (K // 64 + splits - 1) // splits * (splits - 1) < K // 64      # no empty slice
    rows = torch.randint(0, M, (16,), device=cuda, generator=g)
    exp = a[rows].double() @ b.double().T
    err = (out[rows].double() - exp).abs().max()
    assert float(err) < 2e-4 * K ** 0.5 * 4, float(err)
    assert torch.equal(out, ops.gemm_tn(a, b))


@pytest.mark.parametrize("Mn,N,K", [(100_000, 1024, 1024), (50_001, 1024, 1024), (777, 64, 64), (3000, 128, 256), (64, 256, 64)])
def test_wgrad_mn_major_tensor_core(cuda, Mn, N, K):
    """dW = dY^T X with MN-major tcgen05 operands (no transposes) vs float64, ragged node counts included."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(Mn % 1000 + N)
    dy = torch.randn(Mn, N, device=cuda, generator=g).to(torch.bfloat16)
    x = torch.randn(Mn, K, device=cuda, generator=g).to(torch.bfloat16)
    out = ops.wgrad(dy, x)
    rows = torch.randint(0, N, (8,), device=cuda, generator=g)
    exp = dy[:, rows].double().t() @ x.double()
    err = float((out[rows].double() - exp).abs().max())
    assert err < 2e-4 * Mn ** 0.5 * 4, err
    assert torch.equal(out, ops.wgrad(dy, x))                       # deterministic
