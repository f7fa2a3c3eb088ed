"""Forward / backward orchestration of the GCN_FC hot path over the CUDA kernels.

Inference (no grad) goes straight through ``ops``.  Training wraps the same kernels in
``torch.autograd.Function``s whose backward passes are again kernels of liblpgnn (the transposed
SpMM orientation, the data/weight-gradient GEMMs, the head/mask backward) -- never PyTorch-eager
math on the activations.
"""
from __future__ import annotations

import torch

from . import ops
from .graph import BipartiteCSR


_ACT_DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16}
_HALF_TYPES = (torch.bfloat16, torch.float16)


def _act_dtype(model) -> torch.dtype:
    return _ACT_DTYPES.get(getattr(model, "precision", "fp32"), torch.float32)


def _needs_grad(module) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())


def _check_graph(g):
    if not isinstance(g, BipartiteCSR):
        raise TypeError(f"batch.edge_index must be a BipartiteCSR (got {type(g).__name__}); build it with "
                        "BipartiteCSR.from_edge_index(...).to('cuda')")
    return g.views()


# --------------------------------------------------------------------------------------------------
# inference path
# --------------------------------------------------------------------------------------------------
def x2_weights_cached(conv_cache, gc):
    """x2 form of one GraphConv's hidden weights, cached per parameter version: ``((rel_hi, rel_lo), (root_hi, root_lo),
    colscale[N])`` with one power-of-two scale per output feature shared by W_rel and W_root."""
    key = ("x2", id(gc))
    ver = (gc.lin_rel.weight._version, gc.lin_root.weight._version, gc.lin_rel.weight.device)
    hit = conv_cache._c.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    trip = ops.split_x2(gc.lin_rel.weight.detach(), gc.lin_root.weight.detach())
    conv_cache._c[key] = (ver, trip)
    return trip


def use_x2(conv) -> bool:
    """'fp32' mode on the tensor cores: hidden transforms as three half x half passes over x2 operands with chunked
    accumulation (``ops.node_transform_x2``) when the shape allows it; ``conv.fp32_cuda_cores`` forces the CUDA-core
    kernel ('fp32_simt')."""
    gc = conv.left2right
    return gc.in_channels[0] % 64 == 0 and gc.in_channels[1] % 64 == 0 and gc.out_channels % 64 == 0 \
        and not getattr(conv, "fp32_cuda_cores", False)


def _conv_hidden_x2(conv, left, right, csr, csc, relu, heads=None, pre=None):
    """One hidden layer in the fp32 tensor-core mode.  ``heads = ((w, b, feas)_left, (w, b, feas)_right)`` fuses the
    basis-status heads (last layer): returns the logits instead of the activations.  ``pre = ((parts, scale)_left,
    (parts, scale)_right)``: the inputs already exist as x2 operands (written by the input layer), so the aggregation
    writes x2 operands directly and no split pass runs."""
    l2r, r2l = conv.left2right, conv.right2left
    if pre is not None:
        (xs, sxs), (xt, sxt) = pre
        at, st = ops.spmm_x2(csc, left, sxs)      # [n,H]  A^T . left   (sources: constraint rows, scales sxs)
        as_, ss = ops.spmm_x2(csr, right, sxt)    # [m,H]  A   . right
    else:
        agg_t = ops.spmm(csc, left)     # [n,H]  A^T . left
        agg_s = ops.spmm(csr, right)    # [m,H]  A   . right
        at, xt, st = ops.split_x2(agg_t, right)
        as_, xs, ss = ops.split_x2(agg_s, left)
        sxt = sxs = None                # operands split together share their row scale
    wrel_t, wroot_t, cs_t = x2_weights_cached(conv._cache, l2r)
    wrel_s, wroot_s, cs_s = x2_weights_cached(conv._cache, r2l)
    if heads is not None:
        _, logit_t = ops.node_transform_x2(at, wrel_t, xt, wroot_t, st, cs_t, l2r.lin_rel.bias.detach(), relu=relu,
                                           head=heads[1], want_out=False, rowscale2=sxt)
        _, logit_s = ops.node_transform_x2(as_, wrel_s, xs, wroot_s, ss, cs_s, r2l.lin_rel.bias.detach(), relu=relu,
                                           head=heads[0], want_out=False, rowscale2=sxs)
        return logit_s, logit_t
    right_new = ops.node_transform_x2(at, wrel_t, xt, wroot_t, st, cs_t, l2r.lin_rel.bias.detach(), relu=relu, rowscale2=sxt)
    left_new = ops.node_transform_x2(as_, wrel_s, xs, wroot_s, ss, cs_s, r2l.lin_rel.bias.detach(), relu=relu, rowscale2=sxs)
    return left_new, right_new


def _conv_hidden_infer(conv, left, right, csr, csc, relu):
    dt = left.dtype
    cast = conv._cache.get
    l2r, r2l = conv.left2right, conv.right2left
    if dt == torch.float32 and use_x2(conv):
        return _conv_hidden_x2(conv, left, right, csr, csc, relu)
    agg_t = ops.spmm(csc, left)     # [n,H]  A^T . left
    agg_s = ops.spmm(csr, right)    # [m,H]  A   . right
    right_new = ops.node_transform(agg_t, cast(l2r.lin_rel.weight, dt), right, cast(l2r.lin_root.weight, dt),
                                   l2r.lin_rel.bias.detach(), relu=relu)
    left_new = ops.node_transform(agg_s, cast(r2l.lin_rel.weight, dt), left, cast(r2l.lin_root.weight, dt),
                                  r2l.lin_rel.bias.detach(), relu=relu)
    return left_new, right_new


def wcat_bf16(conv_cache, gc, dtype=torch.bfloat16):
    """[W_rel | W_root | 0] as bf16 (or half) [N,64] (cached per parameter version): B operand of the 16-bit input layer."""
    key = ("wcat", id(gc), dtype)
    ver = (gc.lin_rel.weight._version, gc.lin_root.weight._version, gc.lin_rel.weight.device)
    hit = conv_cache._c.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    w_rel, w_root = gc.lin_rel.weight.detach(), gc.lin_root.weight.detach()
    w = torch.zeros((w_rel.shape[0], 64), dtype=dtype, device=w_rel.device)
    w[:, :w_rel.shape[1]] = w_rel
    w[:, w_rel.shape[1]:w_rel.shape[1] + w_root.shape[1]] = w_root
    conv_cache._c[key] = (ver, w)
    return w


def conv_in_bf16(conv, x_left, x_right, csr, csc, relu, want_f32=False, dtype=torch.bfloat16):
    """16-bit input layer: gather_cat (fp32 accumulate, bf16 / half out) -> tensor-core transform with one K block."""
    l2r, r2l = conv.left2right, conv.right2left
    z32_t, zb_t = ops.gather_cat(csc, x_left, x_right, want_f32=want_f32, want_bf16=True, dtype16=dtype)
    z32_s, zb_s = ops.gather_cat(csr, x_right, x_left, want_f32=want_f32, want_bf16=True, dtype16=dtype)
    right_new = ops.node_transform(zb_t, wcat_bf16(conv._cache, l2r, dtype), bias=l2r.lin_rel.bias.detach(), relu=relu)
    left_new = ops.node_transform(zb_s, wcat_bf16(conv._cache, r2l, dtype), bias=r2l.lin_rel.bias.detach(), relu=relu)
    return left_new, right_new, z32_s, z32_t


def _conv_in_infer(conv, x_left, x_right, csr, csc, dt, relu):
    l2r, r2l = conv.left2right, conv.right2left
    k_tot = l2r.in_channels[0] + l2r.in_channels[1]
    if dt in _HALF_TYPES and l2r.in_channels == (8, 8) and r2l.in_channels == (8, 8) and l2r.out_channels % 32 == 0 \
            and l2r.out_channels <= 4096:
        # the reference's shape: aggregate + transform + ReLU + 16-bit store in one kernel per direction
        right_new, _ = ops.conv_in_16(csc, x_left, x_right, l2r.lin_rel.weight.detach(), l2r.lin_rel.bias.detach(),
                                      l2r.lin_root.weight.detach(), dt, relu=relu)
        left_new, _ = ops.conv_in_16(csr, x_right, x_left, r2l.lin_rel.weight.detach(), r2l.lin_rel.bias.detach(),
                                     r2l.lin_root.weight.detach(), dt, relu=relu)
        return left_new, right_new
    if dt in _HALF_TYPES and k_tot <= 64 and l2r.out_channels % 64 == 0:
        left_new, right_new, _, _ = conv_in_bf16(conv, x_left, x_right, csr, csc, relu, dtype=dt)
        return left_new, right_new
    right_new, _ = ops.conv_in_fused(csc, x_left, x_right, l2r.lin_rel.weight.detach(), l2r.lin_rel.bias.detach(),
                                     l2r.lin_root.weight.detach(), dt, relu=relu)
    left_new, _ = ops.conv_in_fused(csr, x_right, x_left, r2l.lin_rel.weight.detach(), r2l.lin_rel.bias.detach(),
                                    r2l.lin_root.weight.detach(), dt, relu=relu)
    return left_new, right_new


def two_direction_forward(conv, left, right, graph, relu=False):
    """GraphConvTwoDirection.forward (reference arch.py:65-81) for callers that use the layer
    on its own."""
    csr, csc = _check_graph(graph)
    if _needs_grad(conv) or left.requires_grad or right.requires_grad:
        from .training import conv_train
        return conv_train(conv, left, right, csr, csc, relu, dropout_p=0.0, training=False)
    narrow = conv.left2right.in_channels[0] + conv.left2right.in_channels[1] <= 64
    if narrow and left.dtype == torch.float32:
        return _conv_in_infer(conv, left, right, csr, csc, torch.float32, relu)
    return _conv_hidden_infer(conv, left, right, csr, csc, relu)


def gcn_fc_forward(model, x_s, x_t, graph):
    """GCN_FC.forward (reference arch.py:179-193)."""
    csr, csc = _check_graph(graph)
    if _needs_grad(model):
        if getattr(model, "precision", "fp32") == "fp16":
            raise NotImplementedError("precision 'fp16' is the reference's inference switch (--fp16, val.py:269): train in "
                                      "'fp32' or 'bf16', or call the model under torch.no_grad()")
        if getattr(graph, "normalize", None):
            raise NotImplementedError("degree-normalised graphs (normalize='mean') are forward / inference only: the "
                                      "backward pass needs the transposes of both normalised orientations")
        from .training import gcn_fc_train
        return gcn_fc_train(model, x_s, x_t, csr, csc)
    dt = _act_dtype(model)
    n_layers = len(model.layers)
    pre = None
    if dt == torch.float32 and n_layers and use_x2(model.layers[0]):
        # fp32 on the tensor cores: the input layer writes its output as fp32 (gather source of the aggregation) AND as x2
        # operands (lin_root side of the first hidden transform)
        l2r, r2l = model.conv1.left2right, model.conv1.right2left
        right, p_t, s_t = ops.conv_in_fused_x2(csc, x_s, x_t, l2r.lin_rel.weight.detach(), l2r.lin_rel.bias.detach(),
                                               l2r.lin_root.weight.detach(), relu=True)
        left, p_s, s_s = ops.conv_in_fused_x2(csr, x_t, x_s, r2l.lin_rel.weight.detach(), r2l.lin_rel.bias.detach(),
                                              r2l.lin_root.weight.detach(), relu=True)
        pre = ((p_s, s_s), (p_t, s_t))
    else:
        left, right = _conv_in_infer(model.conv1, x_s, x_t, csr, csc, dt, relu=True)
    for li, conv in enumerate(model.layers):
        # eval mode: dropout is the identity; relu is fused into the transform epilogue
        if li == n_layers - 1 and dt in _HALF_TYPES:
            # last layer: the head is fused into the transform epilogue; the hidden activation never reaches HBM
            cast = conv._cache.get
            l2r, r2l = conv.left2right, conv.right2left
            agg_t, agg_s = ops.spmm(csc, left), ops.spmm(csr, right)
            logit_t, _ = ops.node_transform_head(agg_t, cast(l2r.lin_rel.weight, dt), right, cast(l2r.lin_root.weight, dt),
                                                 l2r.lin_rel.bias.detach(), model.lin_right.weight.detach(),
                                                 model.lin_right.bias.detach(), x_t)
            logit_s, _ = ops.node_transform_head(agg_s, cast(r2l.lin_rel.weight, dt), left, cast(r2l.lin_root.weight, dt),
                                                 r2l.lin_rel.bias.detach(), model.lin_left.weight.detach(),
                                                 model.lin_left.bias.detach(), x_s)
            return logit_s, logit_t
        if li == n_layers - 1 and dt == torch.float32 and use_x2(conv):
            heads = ((model.lin_left.weight.detach(), model.lin_left.bias.detach(), x_s),
                     (model.lin_right.weight.detach(), model.lin_right.bias.detach(), x_t))
            return _conv_hidden_x2(conv, left, right, csr, csc, True, heads=heads, pre=pre if li == 0 else None)
        if li == 0 and pre is not None:
            left, right = _conv_hidden_x2(conv, left, right, csr, csc, True, pre=pre)
            continue
        left, right = _conv_hidden_infer(conv, left, right, csr, csc, relu=True)
    logit_s, _ = ops.head_mask(left, model.lin_left.weight.detach(), model.lin_left.bias.detach(), x_s)
    logit_t, _ = ops.head_mask(right, model.lin_right.weight.detach(), model.lin_right.bias.detach(), x_t)
    return logit_s, logit_t


def add_knowledge_fn(logits, feas):
    if torch.is_grad_enabled() and logits.requires_grad:
        from .training import AddKnowledgeFn
        return AddKnowledgeFn.apply(logits, feas)
    return ops.add_knowledge_kernel(logits, feas)
