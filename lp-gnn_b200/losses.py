"""Training losses of the reference (train.py:18-53, utils.py:286-299).  On CUDA fp32 logits with int64 labels all
three (``balanced``, ``unbalanced``, ``focal``) are single native calls that return the loss and d(loss)/d(logits)
(csrc/loss.cu); the framework spellings below them are what CPU tensors and the parity tests use."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


@torch.no_grad()
def labels_to_balanced_weights(labels, merge_lu=True):
    """Inverse-frequency class weights; for two-sided problems the lower/upper weights are averaged
    (utils.py:286-299)."""
    # torch.unique(return_counts) of the reference == per-class counts; computed with a comparison + sum so that
    # there is neither a device sort nor a host sync (unique and bincount both read a size back to the host)
    cls = torch.arange(3, device=labels.device)
    cnt = (labels.view(-1, 1) == cls).sum(0).to(torch.float32)
    present = cnt > 0
    res = torch.where(present, cnt.sum() / cnt.clamp_min(1.0), torch.zeros_like(cnt))
    if merge_lu:
        merged = (res[0] + res[2]) / 2.
        two_sided = present.sum() != 2            # the reference merges lower/upper unless exactly 2 classes occur
        res = torch.where(two_sided & (cls != 1), merged, res)     # (no host-built mask: a pageable H2D copy would sync)
    return res


class FocalLoss(nn.modules.loss._WeightedLoss):
    """train.py:18-28 (note: the focal factor is applied to the batch-mean CE, as in the reference)."""

    def __init__(self, weight=None, gamma=2, reduction="mean"):
        super().__init__(weight, reduction=reduction)
        self.gamma = gamma
        self.weight = weight

    def forward(self, input, target):
        ce = F.cross_entropy(input, target, reduction=self.reduction, weight=self.weight)
        pt = torch.exp(-ce)
        return ((1 - pt) ** self.gamma * ce).mean()


cri_focal = FocalLoss()


def unbalanced_torch(logpr_cons, logpr_vars, y_s, y_t):
    """train.py:30-37 with framework ops."""
    return F.cross_entropy(torch.cat((logpr_cons, logpr_vars), dim=0), torch.cat((y_s, y_t), dim=0))


def focal_torch(logpr_cons, logpr_vars, y_s, y_t):
    """train.py:49-53 with framework ops."""
    return cri_focal(torch.cat((logpr_cons, logpr_vars), dim=0), torch.cat((y_s, y_t), dim=0))


def balanced_torch(logpr_cons, logpr_vars, y_s, y_t):
    """balanced() spelled with framework ops, as the reference does (train.py:39-46): CPU tensors and the tests."""
    m, n = len(y_s), len(y_t)
    loss = (m + n) / m * F.cross_entropy(logpr_cons, y_s, weight=labels_to_balanced_weights(y_s))
    loss = loss + (m + n) / n * F.cross_entropy(logpr_vars, y_t, weight=labels_to_balanced_weights(y_t))
    return loss


class _BalancedCE(torch.autograd.Function):
    """``lpgnn_balanced_ce``: the loss value and d(loss)/d(logits) from two kernels."""

    @staticmethod
    def forward(ctx, logit_s, logit_t, y_s, y_t):
        from . import _lib
        lib = _lib.load()
        logit_s, logit_t = logit_s.contiguous(), logit_t.contiguous()
        y_s, y_t = y_s.contiguous(), y_t.contiguous()
        m, n = logit_s.shape[0], logit_t.shape[0]
        dev = logit_s.device
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d = torch.empty((m + n, 3), dtype=torch.float32, device=dev) if need else None
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws_bytes = lib.lpgnn_balanced_ce_workspace_bytes(m, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.lpgnn_balanced_ce(logit_s.data_ptr(), y_s.data_ptr(), m, logit_t.data_ptr(), y_t.data_ptr(), n, 1,
                                       loss.data_ptr(), _lib.ptr(d), d[m:].data_ptr() if need else None, ws.data_ptr(),
                                       ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "lpgnn_balanced_ce")
        ctx.d, ctx.m = d, m
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d * g                       # one small [m+n,3] kernel; g is the scalar upstream gradient
        return d[:ctx.m], d[ctx.m:], None, None


def balanced(logpr_cons, logpr_vars, y_s, y_t):
    """train.py:39-46.  CUDA fp32 logits with int64 labels take the fused kernel; anything else the framework ops."""
    if (logpr_cons.is_cuda and logpr_vars.is_cuda and logpr_cons.dtype == torch.float32 and logpr_vars.dtype == torch.float32
            and y_s.dtype == torch.int64 and y_t.dtype == torch.int64 and len(y_s) > 0 and len(y_t) > 0
            and logpr_cons.shape[1] == 3 and logpr_vars.shape[1] == 3):
        return _BalancedCE.apply(logpr_cons, logpr_vars, y_s, y_t)
    return balanced_torch(logpr_cons, logpr_vars, y_s, y_t)


class _BalancedCEPacked(torch.autograd.Function):
    """``lpgnn_balanced_ce_segmented``: mean over the LPs of a pack of the per-LP balanced loss, and its gradient."""

    @staticmethod
    def forward(ctx, logit_s, logit_t, y_s, y_t, cons_ptr, vars_ptr):
        from . import _lib
        lib = _lib.load()
        logit_s, logit_t = logit_s.contiguous(), logit_t.contiguous()
        y_s, y_t = y_s.contiguous(), y_t.contiguous()
        m, n, B = logit_s.shape[0], logit_t.shape[0], int(cons_ptr.shape[0]) - 1
        dev = logit_s.device
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d = torch.empty((m + n, 3), dtype=torch.float32, device=dev) if need else None
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws_bytes = lib.lpgnn_balanced_ce_segmented_workspace_bytes(B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.lpgnn_balanced_ce_segmented(logit_s.data_ptr(), y_s.data_ptr(), cons_ptr.data_ptr(), logit_t.data_ptr(),
                                                 y_t.data_ptr(), vars_ptr.data_ptr(), B, 1, loss.data_ptr(), _lib.ptr(d),
                                                 d[m:].data_ptr() if need else None, ws.data_ptr(), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "lpgnn_balanced_ce_segmented")
        ctx.d, ctx.m = d, m
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d * g
        return d[:ctx.m], d[ctx.m:], None, None, None, None


def balanced_packed_torch(logpr_cons, logpr_vars, y_s, y_t, cons_ptr, vars_ptr):
    """Mean over the LPs of a pack of ``balanced`` (train.py:39-46 applied per graph), framework ops."""
    c, v = cons_ptr.tolist(), vars_ptr.tolist()
    terms = [balanced_torch(logpr_cons[c[b]:c[b + 1]], logpr_vars[v[b]:v[b + 1]], y_s[c[b]:c[b + 1]], y_t[v[b]:v[b + 1]])
             for b in range(len(c) - 1)]
    return torch.stack(terms).mean()


def balanced_packed(logpr_cons, logpr_vars, y_s, y_t, cons_ptr, vars_ptr):
    """``balanced`` for a block-diagonal pack of LPs (``dataset.pack_bipartite``): every LP keeps its own class weights
    and size factors; the pack's loss is the mean over its LPs, so its gradient is the average of the per-LP gradients."""
    if _native_ok(logpr_cons, logpr_vars, y_s, y_t) and cons_ptr.is_cuda and vars_ptr.is_cuda \
            and cons_ptr.dtype == torch.int32 and vars_ptr.dtype == torch.int32:
        return _BalancedCEPacked.apply(logpr_cons, logpr_vars, y_s, y_t, cons_ptr.contiguous(), vars_ptr.contiguous())
    return balanced_packed_torch(logpr_cons, logpr_vars, y_s, y_t, cons_ptr, vars_ptr)


def _native_ok(logpr_cons, logpr_vars, y_s, y_t):
    return (logpr_cons.is_cuda and logpr_vars.is_cuda and logpr_cons.dtype == torch.float32 and logpr_vars.dtype == torch.float32
            and y_s.dtype == torch.int64 and y_t.dtype == torch.int64 and len(y_s) + len(y_t) > 0
            and logpr_cons.shape[1] == 3 and logpr_vars.shape[1] == 3)


class _FlatCE(torch.autograd.Function):
    """``lpgnn_flat_ce``: unbalanced / focal loss and the un-scaled d(loss)/d(logits) from one kernel; the scalar
    d(loss)/d(mean CE) stays on the device and is folded into the upstream gradient in backward."""

    @staticmethod
    def forward(ctx, logit_s, logit_t, y_s, y_t, focal_gamma):
        from . import _lib
        lib = _lib.load()
        logit_s, logit_t = logit_s.contiguous(), logit_t.contiguous()
        y_s, y_t = y_s.contiguous(), y_t.contiguous()
        m, n = logit_s.shape[0], logit_t.shape[0]
        dev = logit_s.device
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d = torch.empty((m + n, 3), dtype=torch.float32, device=dev) if need else None
        out = torch.empty(2, dtype=torch.float32, device=dev)               # [loss, d loss / d ce]
        ws_bytes = lib.lpgnn_flat_ce_workspace_bytes(m, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.lpgnn_flat_ce(logit_s.data_ptr(), y_s.data_ptr(), m, logit_t.data_ptr(), y_t.data_ptr(), n,
                                   0 if focal_gamma is None else 1, 2.0 if focal_gamma is None else float(focal_gamma),
                                   out.data_ptr(), out[1:].data_ptr(), _lib.ptr(d), d[m:].data_ptr() if need else None,
                                   ws.data_ptr(), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "lpgnn_flat_ce")
        ctx.d, ctx.m, ctx.scale = d, m, out[1]
        return out[0]

    @staticmethod
    def backward(ctx, g):
        d = ctx.d * (g * ctx.scale)
        return d[:ctx.m], d[ctx.m:], None, None, None


def unbalanced(logpr_cons, logpr_vars, y_s, y_t):
    """train.py:30-37.  CUDA fp32 logits with int64 labels take the fused kernel; anything else the framework ops."""
    if _native_ok(logpr_cons, logpr_vars, y_s, y_t):
        return _FlatCE.apply(logpr_cons, logpr_vars, y_s, y_t, None)
    return unbalanced_torch(logpr_cons, logpr_vars, y_s, y_t)


def focal(logpr_cons, logpr_vars, y_s, y_t):
    """train.py:49-53 (gamma = 2, focal factor on the batch-mean CE as in the reference)."""
    if _native_ok(logpr_cons, logpr_vars, y_s, y_t):
        return _FlatCE.apply(logpr_cons, logpr_vars, y_s, y_t, cri_focal.gamma)
    return focal_torch(logpr_cons, logpr_vars, y_s, y_t)


LOSSES = {"balanced": balanced, "unbalanced": unbalanced, "focal": focal}
