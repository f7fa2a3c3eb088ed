"""Training path: GCN_FC forward + hand-derived backward over the CUDA kernels.

One ``torch.autograd.Function`` covers the whole network (reference arch.py:179-193), so autograd sees
the parameters and the two logit tensors only; every activation-sized step of the backward pass is a
kernel of liblpgnn (``ops``).  The loss (reference train.py:32-53) and the optimiser (train.py:85-89)
stay in PyTorch, as in the reference.

Backward algebra for one GraphConvTwoDirection layer (reference arch.py:65-81; A is m x n):
    R' = relu(drop(T W_rel^{l2r T} + R W_root^{l2r T} + b)),  T = A^T L        (variables)
    L' = relu(drop(S W_rel^{r2l T} + L W_root^{r2l T} + b)),  S = A  R        (constraints)
    dPre_t = dR' * mask(R'), dPre_s = dL' * mask(L')                         -> ops.relu_bwd / head_mask_bwd
    dW_rel^{l2r} = dPre_t^T T, dW_root^{l2r} = dPre_t^T R, db = colsum(dPre_t)   -> transposes + GEMM / colsum
    dL = (dPre_s W_root^{r2l} + (A   dPre_t) W_rel^{l2r}) * mask(L)          -> ops.spmm(csr) + two-operand GEMM
    dR = (dPre_t W_root^{l2r} + (A^T dPre_s) W_rel^{r2l}) * mask(R)          -> ops.spmm(csc) + two-operand GEMM
Everything is atomics-free, so a step is bit-reproducible.
"""
from __future__ import annotations

import torch

from . import ops


def _param_list(model):
    """Flat parameter order shared by forward() and backward()."""
    out = []
    for conv in [model.conv1, *model.layers]:
        for gc in (conv.left2right, conv.right2left):
            out += [gc.lin_rel.weight, gc.lin_rel.bias, gc.lin_root.weight]
    out += [model.lin_left.weight, model.lin_left.bias, model.lin_right.weight, model.lin_right.bias]
    return out


def _wgrad_pair(d_pre_t, x_rel, x_root):
    """(dW_rel, dW_root) = d_pre^T [x_rel | x_root] with ONE split-K GEMM over the node dimension: the two
    operands are transposed into the halves of one [K_rel + K_root, M_pad] buffer."""
    k_rel, k_root = x_rel.shape[1], x_root.shape[1]
    xt = torch.empty((k_rel + k_root, d_pre_t.shape[1]), dtype=x_rel.dtype, device=x_rel.device)   # [2K, M_pad]
    ops.transpose(x_rel, out=xt[:k_rel])
    ops.transpose(x_root, out=xt[k_rel:])
    g = ops.gemm_tn(d_pre_t, xt)                                             # [N, 2K] fp32
    k = x_rel.shape[1]
    return g[:, :k].contiguous(), g[:, k:].contiguous()


class _GCNFCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_s, x_t, csr, csc, cfg, *params):
        dt, dp, training, seed = cfg["dtype"], cfg["dp"], cfg["training"], cfg["seed"]
        n_layers = cfg["n_hidden"]
        P = [p.detach() for p in params]
        cast = (lambda w: w.to(dt)) if dt != torch.float32 else (lambda w: w)
        x_s, x_t = x_s.float().contiguous(), x_t.float().contiguous()
        # conv1 (p,q -> H), relu fused.  bf16 mode: gather -> tensor-core transform over one zero-padded K block
        # (same arithmetic as the inference path); fp32 mode: CUDA-core fused kernel.
        if dt == torch.bfloat16 and x_s.shape[1] == 8 and x_t.shape[1] == 8 and P[0].shape[0] % 64 == 0 and P[0].shape[0] <= 4096:
            # the reference's 8 + 8 input features: one kernel per direction, which also emits the bf16 operand
            # [z | 1 | 0] of the layer's tensor-core weight gradient
            right, z_t = ops.conv_in_16(csc, x_s, x_t, P[0], P[1], P[2], dt, relu=True, want_z16=True)
            left, z_s = ops.conv_in_16(csr, x_t, x_s, P[3], P[4], P[5], dt, relu=True, want_z16=True)
        elif dt == torch.bfloat16 and x_s.shape[1] + x_t.shape[1] <= 64 and P[0].shape[0] % 64 == 0:
            def wcat(w_rel, w_root):
                w = torch.zeros((w_rel.shape[0], 64), dtype=dt, device=w_rel.device)
                w[:, :w_rel.shape[1]] = w_rel
                w[:, w_rel.shape[1]:w_rel.shape[1] + w_root.shape[1]] = w_root
                return w
            _, zb_t = ops.gather_cat(csc, x_s, x_t, want_f32=False, want_bf16=True)
            _, zb_s = ops.gather_cat(csr, x_t, x_s, want_f32=False, want_bf16=True)
            right = ops.node_transform(zb_t, wcat(P[0], P[2]), bias=P[1], relu=True)
            left = ops.node_transform(zb_s, wcat(P[3], P[5]), bias=P[4], relu=True)
            z_s, z_t = zb_s, zb_t            # the bf16 operands (with their ones column) feed the tensor-core weight gradient
        else:
            right, z_t = ops.conv_in_fused(csc, x_s, x_t, P[0], P[1], P[2], dt, relu=True)
            left, z_s = ops.conv_in_fused(csr, x_t, x_s, P[3], P[4], P[5], dt, relu=True)
        saved = [z_s, z_t, left, right]
        scale = 1.0
        head_done = False
        for i in range(n_layers):
            w = P[6 + 6 * i: 12 + 6 * i]
            agg_t, agg_s = ops.spmm(csc, left), ops.spmm(csr, right)
            # reference order is dropout then relu_ (arch.py:186-188); relu(drop(x)) == drop(relu(x)), and both are
            # fused into the transform's epilogue
            drop = training and dp > 0
            if dt == torch.bfloat16 and i == n_layers - 1:
                # last hidden layer, bf16: the head rides in the transform's epilogue (as in the native step)
                hw = P[6 + 6 * n_layers:]
                right_new, logit_t, raw_t = ops.node_transform_head_train(agg_t, cast(w[0]), right, cast(w[2]), w[1], hw[2], hw[3],
                                                                          x_t, relu=True, dropout=(dp, seed + 2 * i) if drop else None)
                left_new, logit_s, raw_s = ops.node_transform_head_train(agg_s, cast(w[3]), left, cast(w[5]), w[4], hw[0], hw[1],
                                                                         x_s, relu=True, dropout=(dp, seed + 2 * i + 1) if drop else None)
                head_done = True
            else:
                right_new = ops.node_transform(agg_t, cast(w[0]), right, cast(w[2]), w[1], relu=True,
                                               dropout=(dp, seed + 2 * i) if drop else None)
                left_new = ops.node_transform(agg_s, cast(w[3]), left, cast(w[5]), w[4], relu=True,
                                              dropout=(dp, seed + 2 * i + 1) if drop else None)
            if drop:
                scale = 1.0 / (1.0 - dp)
            saved += [agg_s, agg_t, left_new, right_new]
            left, right = left_new, right_new
        if not head_done:
            hw = P[6 + 6 * n_layers:]
            logit_s, raw_s = ops.head_mask(left, hw[0], hw[1], x_s, want_raw=True)
            logit_t, raw_t = ops.head_mask(right, hw[2], hw[3], x_t, want_raw=True)
        ctx.csr, ctx.csc, ctx.cfg = csr, csc, cfg
        ctx.scale = scale if (training and dp > 0) else 1.0
        ctx.save_for_backward(x_s, x_t, raw_s, raw_t, *saved, *P)
        ctx.n_saved_act = len(saved)
        return logit_s, logit_t

    @staticmethod
    def backward(ctx, d_logit_s, d_logit_t):
        cfg = ctx.cfg
        dt, n_layers = cfg["dtype"], cfg["n_hidden"]
        csr, csc = ctx.csr, ctx.csc
        t = ctx.saved_tensors
        x_s, x_t, raw_s, raw_t = t[:4]
        acts = t[4:4 + ctx.n_saved_act]
        P = t[4 + ctx.n_saved_act:]
        z_s, z_t = acts[0], acts[1]
        cast = (lambda w: w.to(dt)) if dt != torch.float32 else (lambda w: w)
        grads = [None] * len(P)
        hbase = 6 + 6 * n_layers
        # ---- heads: dPre of the last layer's activations (relu/dropout mask fused)
        left, right = acts[2 + 4 * n_layers], acts[3 + 4 * n_layers]
        last_scale = ctx.scale if n_layers > 0 else 1.0
        fuse_cs = n_layers > 0                              # bias gradients of the layer under the head from the same pass
        tc_small = z_s.dtype == torch.bfloat16              # bf16 mode: narrow weight gradients on the tensor cores too
        if tc_small:
            d_pre_s, draw_s, drawb_s, *cs_s = ops.head_mask_bwd(d_logit_s.contiguous(), raw_s, left, P[hbase], last_scale, want_bf16=True, want_colsum=fuse_cs, want_bias_grad=True)
            d_pre_t, draw_t, drawb_t, *cs_t = ops.head_mask_bwd(d_logit_t.contiguous(), raw_t, right, P[hbase + 2], last_scale, want_bf16=True, want_colsum=fuse_cs, want_bias_grad=True)
            grads[hbase] = ops.wgrad(left, drawb_s)[:, :3].t().contiguous()       # [H,64] = left^T [draw | 0]
            grads[hbase + 2] = ops.wgrad(right, drawb_t)[:, :3].t().contiguous()
        else:
            d_pre_s, draw_s, *cs_s = ops.head_mask_bwd(d_logit_s.contiguous(), raw_s, left, P[hbase], last_scale, want_colsum=fuse_cs, want_bias_grad=True)
            d_pre_t, draw_t, *cs_t = ops.head_mask_bwd(d_logit_t.contiguous(), raw_t, right, P[hbase + 2], last_scale, want_colsum=fuse_cs, want_bias_grad=True)
            gw, _ = ops.small_wgrad(left, draw_s, 3)            # [H,3] = left^T draw
            grads[hbase] = gw.t().contiguous()
            gw, _ = ops.small_wgrad(right, draw_t, 3)
            grads[hbase + 2] = gw.t().contiguous()
        grads[hbase + 1], grads[hbase + 3] = cs_s.pop(), cs_t.pop()      # colsum(draw): same pass as head_mask_bwd
        # ---- hidden layers, last to first
        for i in reversed(range(n_layers)):
            w = P[6 + 6 * i: 12 + 6 * i]
            agg_s, agg_t = acts[4 + 4 * i], acts[5 + 4 * i]
            left_in, right_in = acts[2 + 4 * i], acts[3 + 4 * i]       # inputs of this layer (outputs of the previous)
            g = 6 + 6 * i
            if dt == torch.bfloat16:
                # MN-major tensor-core operands: dW = dPre^T X straight from the row-major activations
                grads[g + 0], grads[g + 2] = ops.wgrad(d_pre_t, agg_t), ops.wgrad(d_pre_t, right_in)
                grads[g + 3], grads[g + 5] = ops.wgrad(d_pre_s, agg_s), ops.wgrad(d_pre_s, left_in)
            else:
                dps_t, dpt_t = ops.transpose(d_pre_s), ops.transpose(d_pre_t)
                grads[g + 0], grads[g + 2] = _wgrad_pair(dpt_t, agg_t, right_in)   # l2r lin_rel / lin_root weights
                grads[g + 3], grads[g + 5] = _wgrad_pair(dps_t, agg_s, left_in)    # r2l lin_rel / lin_root weights
                del dps_t, dpt_t
            if i == n_layers - 1:      # the layer under the head: column sums fused into head_mask_bwd (fp32, pre-rounding)
                grads[g + 1], grads[g + 4] = cs_t[0], cs_s[0]
            else:
                grads[g + 1] = ops.colsum(d_pre_t)
                grads[g + 4] = ops.colsum(d_pre_s)
            # data gradients (weights transposed once per step: [K,N] K-major for the TN kernel)
            w_rel_l2r_t, w_root_l2r_t = cast(w[0]).t().contiguous(), cast(w[2]).t().contiguous()
            w_rel_r2l_t, w_root_r2l_t = cast(w[3]).t().contiguous(), cast(w[5]).t().contiguous()
            # dL = A (dPre_t W_rel^{l2r}) + dPre_s W_root^{r2l} = [A dPre_t | dPre_s] [W_rel^{l2r} ; W_root^{r2l}]: aggregate
            # first, then ONE two-operand transform per side (the forward kernel) whose epilogue applies the
            # ReLU / dropout mask of the layer input -- no separate sum / mask pass over the activations
            a_dpre_t = ops.spmm(csr, d_pre_t)                           # A   . dPre_t  [m,H]
            at_dpre_s = ops.spmm(csc, d_pre_s)                          # A^T . dPre_s  [n,H]
            prev_scale = ctx.scale if i > 0 else 1.0                    # conv1 output has no dropout
            d_pre_s, d_pre_t = (
                ops.node_transform(a_dpre_t, w_rel_l2r_t, d_pre_s, w_root_r2l_t, mask=(left_in, prev_scale)),
                ops.node_transform(at_dpre_s, w_rel_r2l_t, d_pre_t, w_root_l2r_t, mask=(right_in, prev_scale)))
        # ---- conv1: weight gradients only (inputs are data)
        k_s, k_t = x_s.shape[1], x_t.shape[1]
        K = k_s + k_t
        if tc_small:                                                        # z = [agg | x_dst | 1 | 0] bf16 [rows,64]
            gw = ops.wgrad(d_pre_t, z_t)
            gb = gw[:, K].contiguous() if K < 64 else ops.colsum(d_pre_t)
            grads[0], grads[1], grads[2] = gw[:, :k_s].contiguous(), gb, gw[:, k_s:K].contiguous()
            gw = ops.wgrad(d_pre_s, z_s)
            gb = gw[:, K].contiguous() if K < 64 else ops.colsum(d_pre_s)
            grads[3], grads[4], grads[5] = gw[:, :k_t].contiguous(), gb, gw[:, k_t:K].contiguous()
        else:
            gw, gb = ops.small_wgrad(d_pre_t, z_t, K, want_bias=True)       # z_t = [A^T x_s | x_t]
            grads[0], grads[1], grads[2] = gw[:, :k_s].contiguous(), gb, gw[:, k_s:].contiguous()
            gw, gb = ops.small_wgrad(d_pre_s, z_s, K, want_bias=True)       # z_s = [A x_t | x_s]
            grads[3], grads[4], grads[5] = gw[:, :k_t].contiguous(), gb, gw[:, k_t:].contiguous()
        return (None, None, None, None, None, *grads)


_ws_high_water = {}          # device -> bytes requested for the native step's workspace (see _NativeTrainFunction.forward)


class _NativeTrainFunction(torch.autograd.Function):
    """The same step through ``lpgnn_train_forward`` / ``lpgnn_train_backward``: two C calls enqueue every
    kernel; activations live in one workspace tensor that the autograd node keeps alive; the parameter
    gradients are views of one flat fp32 buffer."""

    @staticmethod
    def forward(ctx, x_s, x_t, csr, csc, cfg, *params):
        import ctypes as C

        from . import _lib
        lib = _lib.load()
        w = cfg["weights"]
        dev = x_s.device
        x_s, x_t = x_s.float().contiguous(), x_t.float().contiguous()
        m, n = x_s.shape[0], x_t.shape[0]
        ws_bytes = lib.lpgnn_train_workspace_bytes(m, n, w.p, w.q, w.hids, w.depth, w.precision)
        # Sampled mini-batches differ by a few per cent in size: ask the caching allocator for the SAME size every step (a
        # high-water mark with 6 % headroom when it has to grow), so that a batch slightly larger than all before it reuses
        # the cached block instead of triggering a multi-GB cudaMalloc in the middle of training.
        high = _ws_high_water.get(dev, 0)
        if ws_bytes > high or ws_bytes < high // 2:
            high = _ws_high_water[dev] = ws_bytes + ws_bytes // 16 if ws_bytes > high else ws_bytes
        ws_bytes = max(ws_bytes, high)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        logits = torch.empty((m + n, 3), dtype=torch.float32, device=dev)
        dp = cfg["dp"] if cfg["training"] else 0.0
        with torch.cuda.device(dev):
            rc = lib.lpgnn_train_forward(C.byref(w), csr[0].data_ptr(), csr[1].data_ptr(), csr[2].data_ptr(),
                                         csc[0].data_ptr(), csc[1].data_ptr(), csc[2].data_ptr(), m, n,
                                         x_s.data_ptr(), x_t.data_ptr(), dp, cfg["seed"], logits.data_ptr(),
                                         logits[m:].data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "lpgnn_train_forward")
        ctx.csr, ctx.csc, ctx.cfg, ctx.dp, ctx.mn = csr, csc, cfg, dp, (m, n)
        ctx.ws = ws
        ctx.shapes = [p.shape for p in params]
        ctx.keep = (x_s, x_t)
        return logits[:m], logits[m:]

    @staticmethod
    def backward(ctx, d_logit_s, d_logit_t):
        import ctypes as C

        from . import _lib
        lib = _lib.load()
        w, (m, n), csr, csc = ctx.cfg["weights"], ctx.mn, ctx.csr, ctx.csc
        if ctx.ws is None:
            raise RuntimeError("the native training step frees its activations after backward (no retain_graph)")
        dev = ctx.ws.device
        d_logit_s = torch.zeros((m, 3), device=dev) if d_logit_s is None else d_logit_s.float().contiguous()
        d_logit_t = torch.zeros((n, 3), device=dev) if d_logit_t is None else d_logit_t.float().contiguous()
        sizes = [int(torch.Size(s).numel()) for s in ctx.shapes]
        offs = [0]
        for k in sizes:
            offs.append(offs[-1] + (k + 3) // 4 * 4)                       # 16-byte aligned views
        flat = torch.empty(offs[-1], dtype=torch.float32, device=dev)
        grads = [flat[o:o + k].view(s) for o, k, s in zip(offs, sizes, ctx.shapes)]
        g = _lib.GcnFcGrads()
        names = ["c1_l2r_wrel", "c1_l2r_b", "c1_l2r_wroot", "c1_r2l_wrel", "c1_r2l_b", "c1_r2l_wroot"]
        for k, nm in enumerate(names):
            setattr(g, nm, grads[k].data_ptr())
        nh = w.depth - 2
        for i in range(nh):
            for k, nm in enumerate(("l2r_wrel", "l2r_b", "l2r_wroot", "r2l_wrel", "r2l_b", "r2l_wroot")):
                getattr(g, nm)[i] = grads[6 + 6 * i + k].data_ptr()
        for k, nm in enumerate(("head_left_w", "head_left_b", "head_right_w", "head_right_b")):
            setattr(g, nm, grads[6 + 6 * nh + k].data_ptr())
        ws = ctx.ws

        def run(phases):
            with torch.cuda.device(dev):
                rc = lib.lpgnn_train_backward_ex(C.byref(w), csr[0].data_ptr(), csr[1].data_ptr(), csr[2].data_ptr(),
                                                 csc[0].data_ptr(), csc[1].data_ptr(), csc[2].data_ptr(), m, n, ctx.dp,
                                                 d_logit_s.data_ptr(), d_logit_t.data_ptr(), C.byref(g), phases,
                                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "lpgnn_train_backward")

        sync = _gradient_sync[0]
        if sync is None:
            run(_lib.BWD_TAIL | _lib.BWD_REST)
        else:
            # data parallel: the head + last-hidden-layer gradients (the tail of the flat buffer, ~all of the parameters at
            # depth 3) are reduced while the rest of the backward pass runs; the small remainder follows
            cut = offs[6 + 6 * (nh - 1)] if nh > 0 else offs[6]
            run(_lib.BWD_TAIL)
            h1 = sync(flat[cut:])
            run(_lib.BWD_REST)
            h0 = sync(flat[:cut])
            for h in (h1, h0):
                if h is not None:
                    h.wait()
            if _gradient_sync[1] != 1.0:
                flat.mul_(_gradient_sync[1])
            _gradient_sync[2] = 1
        ctx.ws = None
        return (None, None, None, None, None, *grads)


# [reduce(flat_slice) -> handle with .wait() | None, scale applied to the whole buffer afterwards, #backward passes synced]
_gradient_sync = [None, 1.0, 0]


def set_gradient_sync(reduce_fn=None, scale=1.0):
    """Data-parallel hook of the native training step: ``reduce_fn(flat_slice)`` is called on the tail of the flat
    gradient buffer right after it is complete and on the head at the end (it starts an in-place SUM all-reduce and
    returns a handle with ``.wait()``, or None), then the buffer is multiplied by ``scale`` (1 / world).  ``None``
    restores the one-call backward.  `train.allreduce_gradients` skips gradients that were reduced here."""
    _gradient_sync[0], _gradient_sync[1] = reduce_fn, float(scale)


def enable_overlapped_allreduce(world):
    """Mean over ranks with the all-reduce of the tail gradients overlapped with the rest of the backward pass
    (torch.distributed, async_op: NCCL runs it on its own stream behind the kernels enqueued so far)."""
    if world <= 1:
        return set_gradient_sync(None)
    import torch.distributed as dist
    set_gradient_sync(lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True), 1.0 / world)


def consume_synced_backward() -> bool:
    """True once per backward pass whose gradients were already reduced by the hook above."""
    done = _gradient_sync[2] > 0
    _gradient_sync[2] = 0
    return done


def _train_weights(model, params):
    """``lpgnn_gcn_fc_weights`` over the fp32 MASTER parameters (no copies: the optimiser updates them in place, so
    the struct stays valid until a parameter is re-allocated).  None when the native step does not cover the model."""
    from . import _lib
    bf16 = model.precision == "bf16"
    c1 = model.conv1.left2right
    p_, q_ = c1.in_channels[0], c1.in_channels[1]
    nh = len(model.layers)
    if nh > _lib.MAX_HIDDEN_LAYERS or any(t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda for t in params):
        return None
    if bf16 and (model.hids % 64 != 0 or p_ + q_ > 64):
        return None
    key = tuple(t.data_ptr() for t in params) + (model.precision,)
    hit = getattr(model, "_train_cache", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    w = _lib.GcnFcWeights()
    w.p, w.q, w.hids, w.depth = p_, q_, model.hids, nh + 2
    w.precision = _lib.BF16 if bf16 else _lib.F32
    for k, nm in enumerate(("c1_l2r_wrel", "c1_l2r_b", "c1_l2r_wroot", "c1_r2l_wrel", "c1_r2l_b", "c1_r2l_wroot")):
        setattr(w, nm, params[k].data_ptr())
    for i in range(nh):
        for k, nm in enumerate(("l2r_wrel", "l2r_b", "l2r_wroot", "r2l_wrel", "r2l_b", "r2l_wroot")):
            getattr(w, nm)[i] = params[6 + 6 * i + k].data_ptr()
    for k, nm in enumerate(("head_left_w", "head_left_b", "head_right_w", "head_right_b")):
        setattr(w, nm, params[6 + 6 * nh + k].data_ptr())
    model._train_cache = (key, w)
    return w


_step_counter = [0]


def gcn_fc_train(model, x_s, x_t, csr, csc):
    dt = torch.bfloat16 if model.precision == "bf16" else torch.float32      # fp32 / fp32_tc train in fp32
    _step_counter[0] += 1
    cfg = dict(dtype=dt, dp=float(model.dp), training=bool(model.training), n_hidden=len(model.layers),
               seed=(int(torch.initial_seed()) * 1_000_003 + _step_counter[0] * 7919) & (2 ** 62 - 1))
    params = _param_list(model)
    if getattr(model, "native_train", True):
        w = _train_weights(model, params)
        if w is not None:
            cfg["weights"] = w
            return _NativeTrainFunction.apply(x_s, x_t, csr, csc, cfg, *params)
    return _GCNFCFunction.apply(x_s, x_t, csr, csc, cfg, *params)


def conv_train(conv, left, right, csr, csc, relu, dropout_p=0.0, training=False):
    raise NotImplementedError("stand-alone GraphConvTwoDirection training is not wired up; train through GCN_FC")


class AddKnowledgeFn(torch.autograd.Function):
    """add_knowledge with gradient (reference arch.py:129-141) for callers that own the head."""

    @staticmethod
    def forward(ctx, logits, feas):
        ctx.save_for_backward(logits)
        return ops.add_knowledge_kernel(logits, feas)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        x = x.float()
        nrm = x.norm(dim=1, keepdim=True).clamp_min(1e-12)      # [rows,3]: tiny, not an activation-sized tensor
        u = x / nrm
        return (10.0 / nrm) * (g - u * (u * g).sum(1, keepdim=True)), None


def smoke_step(dev):
    """One training step on a small LP (forward, balanced loss, backward, Adam) -- used by smoke()."""
    import types

    import numpy as np

    from . import arch, synth
    from .graph import BipartiteCSR
    from .losses import balanced
    lp = synth.processed_lp(400, 800, 4000, seed=5)
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=64, depth=3).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    batch = types.SimpleNamespace(
        x_s=torch.from_numpy(lp.c_feas).to(dev), x_t=torch.from_numpy(lp.v_feas).to(dev),
        edge_index=BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True))
    y_s, y_t = torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev)
    losses = []
    for _ in range(3):
        lc, lv = model(batch)
        loss = balanced(lc, lv, y_s, y_t)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)), losses
    print(f"smoke[train]: losses {losses}")
