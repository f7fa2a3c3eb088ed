"""Drop-in for the reference ``arch.py`` (GCN_FC stack): same class names, constructor signatures,
``forward(batch) -> (logit_cons[m,3], logit_vars[n,3])`` and state_dict keys, so
``model = eval(args.arch)`` after ``from arch import *`` (reference train.py:9,79; val.py:6,266;
scripts/pred_basis.py:5,128) and an existing ``mdl.pth`` keep working -- but every numeric step runs
in the hand-written sm_100a kernels of ``liblpgnn.so``:

    reference (arch.py)                                   here
    --------------------------------------------------    ------------------------------------------
    GraphConvTwoDirection.forward, conv1   (65-81, 181)   ops.conv_in_fused  x2   (aggregate+transform+relu)
    GraphConvTwoDirection.forward, hidden  (65-81, 185)   ops.spmm x2 + ops.node_transform x2 (tcgen05 / fp32)
    F.dropout + relu_                      (186-188)      fused into the transform epilogue (relu) + mask
    lin_left / lin_right + add_knowledge   (190-191)      ops.head_mask x2

Precision: ``model.precision = 'fp32'`` (default; logits within 1e-4 of the reference), ``'bf16'`` (bf16
activations / tensor-core GEMMs, fp32 accumulate; within 2e-2) or ``'fp16'`` (IEEE half storage, same kernels and
speed, 8x smaller rounding error; inference only).  ``.half()`` selects ``'fp16'`` -- the reference's ``--fp16``
inference switch (``model.half()``, val.py:269, pred_basis.py:146) -- and ``.bfloat16()`` the bf16 mode, while the
parameters stay fp32 master copies.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import ops
from .graph import BipartiteCSR

__all__ = ["GraphConv", "GraphConvTwoDirection", "GCNBase", "GCN_FC", "add_knowledge"]


class _Linear(nn.Module):
    """Parameter container with torch_geometric ``Linear``'s layout and default init
    (weight [out,in] kaiming_uniform(a=sqrt 5); bias U(+-1/sqrt(in)))."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_features) if self.in_features > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)


class GraphConv(nn.Module):
    """PyG ``GraphConv((in_src, in_dst), out, aggr='add')`` parameter layout: ``lin_rel`` (with
    bias) acts on the aggregated source features, ``lin_root`` (no bias) on the destination's own."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel = _Linear(in_channels[0], out_channels, bias=True)
        self.lin_root = _Linear(in_channels[1], out_channels, bias=False)


class _CastCache:
    """bf16 copies of fp32 master weights, refreshed when the parameter changes (version counter)."""

    def __init__(self):
        self._c = {}

    def get(self, p: torch.Tensor, dtype):
        if p.dtype == dtype:
            return p.detach()
        key = id(p)
        hit = self._c.get(key)
        if hit is not None and hit[0] == p._version and hit[1].device == p.device and hit[1].dtype == dtype:
            return hit[1]
        c = p.detach().to(dtype).contiguous()
        self._c[key] = (p._version, c)
        return c


class GraphConvTwoDirection(nn.Module):
    """Reference arch.py:51-81.  ``left`` = constraints, ``right`` = variables; both directions read
    the OLD features (synchronous update):
        right' = lin_rel_l2r(A^T left) + lin_root_l2r(right)
        left'  = lin_rel_r2l(A  right) + lin_root_r2l(left)
    """

    def __init__(self, left_dim, right_dim, out_dim):
        super().__init__()
        self.left2right = GraphConv((left_dim, right_dim), out_dim)
        self.right2left = GraphConv((right_dim, left_dim), out_dim)
        self._cache = _CastCache()

    def forward(self, left_feas, right_feas, edge_index: BipartiteCSR, edge_weight=None, relu=False):
        from .autograd import two_direction_forward
        return two_direction_forward(self, left_feas, right_feas, edge_index, relu)


class GCNBase(nn.Module):
    """Reference arch.py:107-114: weights-only checkpointing."""

    def save(self, pn):
        torch.save(self.state_dict(), pn)

    def load(self, pn):
        st = torch.load(pn, map_location=torch.device("cpu"))
        self.load_state_dict(st)


def add_knowledge(left_logit, right_logit, left_feas, right_feas, bound=10):
    """Reference arch.py:129-141 on the device: row L2-normalise x10, then -10 on class 0 / class 2
    where the lower / upper bound tag of the node is non-zero."""
    if bound != 10:
        raise ValueError("the masking kernel implements the reference's bound=10")
    from .autograd import add_knowledge_fn
    return add_knowledge_fn(left_logit, left_feas), add_knowledge_fn(right_logit, right_feas)


class GCN_FC(GCNBase):
    """Reference arch.py:167-193: conv1 (p,q -> hids) + (depth-2) hidden convs + two Linear(hids,3)
    heads + knowledge masking.  Same constructor argument order as the reference, so a seeded
    construction draws the same initial weights."""

    def __init__(self, p, q, hids=128, depth=3, dp=.1, *args, **kwargs):
        super().__init__()
        self.conv1 = GraphConvTwoDirection(p, q, hids)
        self.layers = nn.ModuleList()
        for _ in range(depth - 2):
            self.layers.append(GraphConvTwoDirection(hids, hids, hids))
        self.lin_left = nn.Linear(hids, 3)
        self.lin_right = nn.Linear(hids, 3)
        self.dp = dp
        self.hids = hids
        self.set_precision("fp32")

    # -- precision switches ----------------------------------------------------------------
    def set_precision(self, precision: str):
        """'fp32'      the reference's default arithmetic (`--fp16 0`, utils.py:770): fp32 storage; hidden transforms on the
                    tensor cores from x2 operands (three half x half passes, chunked fp32 accumulation: logits within
                    1e-4 of the reference, csrc/gemm_x2.cu) when hids % 64 == 0, else the CUDA-core kernel;
        'fp32_simt' fp32 storage, every transform on the CUDA cores (csrc/gemm_simt.cu; the cross-check of 'fp32');
        'fp32_tc'   alias of 'fp32' (the name of round 1's tensor-core fp32 mode);
        'bf16'      bf16 storage and single-pass tensor-core transforms (within 2e-2);
        'fp16'      IEEE half storage, the same tensor-core kernels at the same rate (fp32 accumulate): logits within
                    ~2e-3, statuses agree with fp32 on >= 99.9 % of the nodes.  Inference only, like the reference's
                    `--fp16` (val.py:269): activations must stay below 65504, which scaled LPs (|A|, |c| <= 1) do."""
        if precision == "fp32_tc":
            precision = "fp32"
        if precision not in ("fp32", "fp32_simt", "bf16", "fp16"):
            raise ValueError("precision must be 'fp32', 'fp32_simt', 'bf16' or 'fp16'")
        if precision in ("bf16", "fp16") and self.hids % 64 != 0:
            raise ValueError("16-bit modes need hids to be a multiple of 64 (tensor-core tile)")
        self.precision = precision
        for conv in self.layers:
            conv.fp32_cuda_cores = precision == "fp32_simt"
        return self

    def half(self):       # reference `--fp16`: model.half() (val.py:269, pred_basis.py:146)
        return self.set_precision("fp16")

    def bfloat16(self):
        return self.set_precision("bf16")

    def float(self):
        return self.set_precision("fp32")

    def forward(self, batch):
        from .autograd import gcn_fc_forward
        return gcn_fc_forward(self, batch.x_s, batch.x_t, batch.edge_index)

    # -- native one-call prediction ---------------------------------------------------------
    def _native_weights(self):
        """``lpgnn_gcn_fc_weights`` for the current parameters (rebuilt when a parameter changes)."""
        from . import _lib
        from .autograd import wcat_bf16
        # (walking the module tree costs ~100 us per call: the Parameter objects are listed once, then only their
        # version counters and storage addresses are compared)
        plist = getattr(self, "_param_list", None)
        if plist is None or len(plist) != 6 * (len(self.layers) + 1) + 4:
            plist = self._param_list = list(self.parameters())
        ver = tuple(p._version for p in plist) + (plist[0].data_ptr(), plist[-1].data_ptr(), self.precision)
        hit = getattr(self, "_native_cache", None)
        if hit is not None and hit[0] == ver:
            return hit[1]
        bf16 = self.precision in ("bf16", "fp16")            # 16-bit storage: the tensor-core path
        dt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(self.precision, torch.float32)
        keep = []                                           # tensors the struct points into

        def f32(t):
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def cd(t):
            t = t.detach().to(dt).contiguous()
            keep.append(t)
            return t.data_ptr()

        w = _lib.GcnFcWeights()
        c1 = self.conv1
        w.p, w.q = c1.left2right.in_channels[0], c1.left2right.in_channels[1]
        w.hids, w.depth, w.precision = self.hids, len(self.layers) + 2, _lib.dtype_code(dt)
        if w.depth - 2 > _lib.MAX_HIDDEN_LAYERS:
            raise ValueError("native prediction supports at most 8 hidden layers")
        for tag, gc in (("l2r", c1.left2right), ("r2l", c1.right2left)):
            setattr(w, f"c1_{tag}_wrel", f32(gc.lin_rel.weight))
            setattr(w, f"c1_{tag}_b", f32(gc.lin_rel.bias))
            setattr(w, f"c1_{tag}_wroot", f32(gc.lin_root.weight))
            if bf16:
                wc = wcat_bf16(c1._cache, gc, dt)
                keep.append(wc)
                setattr(w, f"c1_{tag}_wcat", wc.data_ptr())
        from .autograd import use_x2, x2_weights_cached
        for i, conv in enumerate(self.layers):
            for tag, gc in (("l2r", conv.left2right), ("r2l", conv.right2left)):
                getattr(w, f"{tag}_wrel")[i] = cd(gc.lin_rel.weight)
                getattr(w, f"{tag}_wroot")[i] = cd(gc.lin_root.weight)
                getattr(w, f"{tag}_b")[i] = f32(gc.lin_rel.bias)
                if not bf16 and use_x2(conv):
                    wrel, wroot, wscale = x2_weights_cached(conv._cache, gc)
                    keep.extend([*wrel, *wroot, wscale])
                    getattr(w, f"{tag}_wrel_hi")[i], getattr(w, f"{tag}_wrel_lo")[i] = wrel[0].data_ptr(), wrel[1].data_ptr()
                    getattr(w, f"{tag}_wroot_hi")[i], getattr(w, f"{tag}_wroot_lo")[i] = wroot[0].data_ptr(), wroot[1].data_ptr()
                    getattr(w, f"{tag}_wscale")[i] = wscale.data_ptr()
        w.head_left_w, w.head_left_b = f32(self.lin_left.weight), f32(self.lin_left.bias)
        w.head_right_w, w.head_right_b = f32(self.lin_right.weight), f32(self.lin_right.bias)
        self._native_cache = (ver, w, keep)
        return w

    @torch.no_grad()
    def predict_basis_coo(self, row, col, val, m, n, x_s, x_t, is_sorted=False, want_logits=False):
        """Graph build + forward + basis decision as ONE native call (``lpgnn_predict_basis``): device int32
        ``row/col``, fp32 ``val`` COO of the m x n matrix and fp32 features -> uint8 statuses [m+n] (constraints
        first) [+ fp32 logits [m+n,3]].  The sorted claim / index range are reported in ``self.last_graph_status``
        (device int32, bit 0 = not sorted, bit 1 = out of range)."""
        import ctypes as C

        from . import _lib
        from .graph import _status_slot
        _lib.require_cuda(row, col, val, x_s, x_t)
        lib = _lib.load()
        w = self._native_weights()
        dev = x_s.device
        z = int(row.shape[0])
        x2 = _lib.WS_X2 if (w.precision == _lib.F32 and w.depth > 2 and w.l2r_wrel_hi[0]) else 0
        ws_bytes = lib.lpgnn_predict_workspace_bytes(z, m, n, w.p, w.q, w.hids, w.depth, w.precision | x2)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        status = torch.empty(m + n, dtype=torch.uint8, device=dev)
        logits = torch.empty((m + n, 3), dtype=torch.float32, device=dev) if want_logits else None
        gstat = _status_slot(dev)
        with torch.cuda.device(dev):
            rc = lib.lpgnn_predict_basis(C.byref(w), row.data_ptr(), col.data_ptr(), val.data_ptr(), z, m, n,
                                         _lib.COO_SORTED if is_sorted else 0, x_s.data_ptr(), x_t.data_ptr(),
                                         status.data_ptr(), _lib.ptr(logits), gstat.data_ptr(), ws.data_ptr(), ws_bytes,
                                         _lib.stream_ptr())
        _lib.check(rc, "lpgnn_predict_basis")
        self.last_graph_status = gstat
        return (status, logits) if want_logits else status

    @torch.no_grad()
    def predict_basis_packed(self, row, col, val, m, n, x_s, x_t, cons_ptr, vars_ptr, is_sorted=True, want_logits=False,
                             buffers=None, lp_major=False):
        """Same for a block-diagonal pack of LPs (``row/col`` already in the pack's numbering, ``m/n`` the pack
        totals, ``cons_ptr/vars_ptr`` device int32 [B+1]): one forward pass over the pack, basis decision per LP.
        Returns uint8 statuses [m+n] in the packed layout (all constraints, then all variables), or with ``lp_major`` LP by
        LP (constraints of LP b, then its variables, starting at ``cons_ptr[b] + vars_ptr[b]``)."""
        import ctypes as C

        from . import _lib
        from .graph import _status_slot
        _lib.require_cuda(row, col, val, x_s, x_t, cons_ptr, vars_ptr)
        lib = _lib.load()
        w = self._native_weights()
        dev = x_s.device
        z = int(row.shape[0])
        x2 = _lib.WS_X2 if (w.precision == _lib.F32 and w.depth > 2 and w.l2r_wrel_hi[0]) else 0
        ws_bytes = lib.lpgnn_predict_workspace_bytes(z, m, n, w.p, w.q, w.hids, w.depth, w.precision | x2)
        if buffers is not None:
            # caller-owned grow-only buffers (a sweep must not go through cudaMalloc for every new pack size):
            # buffers = {"ws": uint8 tensor or None, "status": uint8 tensor or None}, replaced in place when too small
            if buffers.get("ws") is None or buffers["ws"].numel() < ws_bytes:
                buffers["ws"] = torch.empty(int(ws_bytes * 1.25) + 256, dtype=torch.uint8, device=dev)
            if buffers.get("status") is None or buffers["status"].numel() < m + n:
                buffers["status"] = torch.empty(int((m + n) * 1.25) + 256, dtype=torch.uint8, device=dev)
            ws, status = buffers["ws"], buffers["status"][:m + n]
        else:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            status = torch.empty(m + n, dtype=torch.uint8, device=dev)
        logits = torch.empty((m + n, 3), dtype=torch.float32, device=dev) if want_logits else None
        gstat = _status_slot(dev)
        with torch.cuda.device(dev):
            rc = lib.lpgnn_predict_basis_packed(C.byref(w), row.data_ptr(), col.data_ptr(), val.data_ptr(), z, m, n,
                                                (_lib.COO_SORTED if is_sorted else 0) | (_lib.STATUS_LP_MAJOR if lp_major else 0),
                                                x_s.data_ptr(), x_t.data_ptr(),
                                                cons_ptr.data_ptr(), vars_ptr.data_ptr(), int(cons_ptr.shape[0]) - 1,
                                                status.data_ptr(), _lib.ptr(logits), gstat.data_ptr(), ws.data_ptr(),
                                                ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "lpgnn_predict_basis_packed")
        self.last_graph_status = gstat
        return (status, logits) if want_logits else status

    @torch.no_grad()
    def predict_basis(self, batch, int64=True):
        """forward + ``val.inference_gnn`` without leaving the device (reference
        scripts/pred_basis.py:113-118 ``inference_only``)."""
        lc, lv = self.forward(batch)
        return ops.basis_select(lc, lv, k_basic=lc.shape[0], int64=int64)
