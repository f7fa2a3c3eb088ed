"""Python entry points of the CUDA kernels (thin wrappers over the C ABI, ``include/lpgnn.h``).

Each function validates shapes/dtypes, allocates outputs with torch (the library never allocates
user-visible memory), and enqueues the kernel on torch's CURRENT stream.  No function here has a
CPU or PyTorch-eager fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import EPI_NONE, EPI_RELU, check, dtype_code, ptr, require_cuda, stream_ptr  # noqa: F401


def _contig(t):
    return t if t.is_contiguous() else t.contiguous()


def spmm(view, x: torch.Tensor, slab_bytes: int = 0, unroll: int = 0) -> torch.Tensor:
    """``Y[i] = sum_e val[e] * X[idx[e]]`` for one orientation ``view = (ptr, idx, val, rows)``
    (``BipartiteCSR.views()``).  Replaces ``torch_sparse.matmul(adj_t, x, reduce='add')``
    (reference arch.py:75-80 through PyG GraphConv).  ``slab_bytes`` / ``unroll`` select the kernel explicitly
    (``lpgnn_spmm_ex``; 0 = automatic, ``slab_bytes=-1`` = row-per-warp kernel); results do not depend on them."""
    ptr_, idx, val, rows = view
    require_cuda(ptr_, x)
    x = _contig(x)
    F = x.shape[1]
    y = torch.empty((rows, F), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().lpgnn_spmm_ex(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x.data_ptr(), y.data_ptr(), F,
                                       dtype_code(x.dtype), slab_bytes, unroll, stream_ptr())
    check(rc, "lpgnn_spmm")
    return y


def spmm_pair(csr, csc, left, right, nnz=-1):
    """Both aggregations of a layer in one launch where both take the banded sweep (``lpgnn_spmm_pair``):
    ``agg_s[m,F] = A @ right`` (CSR view), ``agg_t[n,F] = A^T @ left`` (CSC view).  Returns ``(agg_s, agg_t)``, the bits of
    two ``spmm`` calls.  ``nnz`` only steers how the SMs are split between the two sides (-1: unknown)."""
    rowptr, col, val, m = csr
    colptr, row_csc, val_csc, n = csc
    require_cuda(rowptr, colptr, left, right)
    left, right = _contig(left), _contig(right)
    if left.dtype != right.dtype or left.shape[1] != right.shape[1]:
        raise TypeError("spmm_pair: both feature matrices must share dtype and width")
    F = left.shape[1]
    agg_s = torch.empty((m, F), dtype=left.dtype, device=left.device)
    agg_t = torch.empty((n, F), dtype=left.dtype, device=left.device)
    with torch.cuda.device(left.device):
        rc = _lib.load().lpgnn_spmm_pair(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), m, colptr.data_ptr(), row_csc.data_ptr(),
                                         val_csc.data_ptr(), n, int(nnz), left.data_ptr(), right.data_ptr(), agg_s.data_ptr(),
                                         agg_t.data_ptr(), F, dtype_code(left.dtype), stream_ptr())
    check(rc, "lpgnn_spmm_pair")
    return agg_s, agg_t


def conv_in_fused(view, x_src, x_dst, w_rel, b_rel, w_root, out_dtype, relu=True):
    """conv1 of GCN_FC for one direction: ``relu(lin_rel(A_view @ x_src) + lin_root(x_dst))`` (reference
    arch.py:75-80, 181-182).  Returns ``(out[rows,N], z_cat[rows,KT])`` where ``z_cat = [A_view@x_src | x_dst | 0]``
    (fp32) is the transform input, kept for the weight gradient."""
    ptr_, idx, val, rows = view
    require_cuda(ptr_, x_src, x_dst, w_rel, w_root)
    x_src, x_dst = _contig(x_src.float()), _contig(x_dst.float())
    w_rel, w_root, b_rel = _contig(w_rel.float()), _contig(w_root.float()), _contig(b_rel.float())
    N = w_rel.shape[0]
    lib = _lib.load()
    kt = lib.lpgnn_conv_in_zcat_width(x_src.shape[1], x_dst.shape[1])
    out = torch.empty((rows, N), dtype=out_dtype, device=x_src.device)
    z_cat = torch.empty((rows, kt), dtype=torch.float32, device=x_src.device)
    with torch.cuda.device(x_src.device):
        rc = lib.lpgnn_conv_in_fused(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x_src.data_ptr(),
                                     x_src.shape[1], x_dst.data_ptr(), x_dst.shape[1], w_rel.data_ptr(),
                                     b_rel.data_ptr(), w_root.data_ptr(), N, out.data_ptr(), dtype_code(out_dtype),
                                     EPI_RELU if relu else EPI_NONE, z_cat.data_ptr(), stream_ptr())
    check(rc, "lpgnn_conv_in_fused")
    return out, z_cat


def conv_in_16(view, x_src, x_dst, w_rel, b_rel, w_root, out_dtype, relu=True, want_z16=False):
    """conv1 of GCN_FC(8, 8, ...) for one direction in the 16-bit modes, ONE kernel (``lpgnn_conv_in_16``: aggregate,
    16-wide MMA with register accumulators, bias, ReLU, 16-bit store).  Returns ``(out[rows,N], z16[rows,64] | None)``."""
    ptr_, idx, val, rows = view
    require_cuda(ptr_, x_src, x_dst, w_rel, w_root, b_rel)
    x_src, x_dst = _contig(x_src.float()), _contig(x_dst.float())
    w_rel, w_root, b_rel = _contig(w_rel.float()), _contig(w_root.float()), _contig(b_rel.float())
    if x_src.shape[1] != 8 or x_dst.shape[1] != 8:
        raise ValueError("conv_in_16 covers the reference's 8 + 8 input features")
    N = w_rel.shape[0]
    out = torch.empty((rows, N), dtype=out_dtype, device=x_src.device)
    z16 = torch.empty((rows, 64), dtype=out_dtype, device=x_src.device) if want_z16 else None
    with torch.cuda.device(x_src.device):
        rc = _lib.load().lpgnn_conv_in_16(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x_src.data_ptr(),
                                          x_dst.data_ptr(), w_rel.data_ptr(), b_rel.data_ptr(), w_root.data_ptr(), N,
                                          out.data_ptr(), dtype_code(out_dtype), EPI_RELU if relu else EPI_NONE, ptr(z16),
                                          stream_ptr())
    check(rc, "lpgnn_conv_in_16")
    return out, z16


def conv_in_16_pair(csr, csc, x_s, x_t, l2r, r2l, out_dtype, relu=True, want_z16=False):
    """Both directions of conv1 in ONE launch (``lpgnn_conv_in_16_pair``).  ``csr`` / ``csc`` = the two views of the
    graph, ``l2r`` / ``r2l`` = ``(w_rel, b_rel, w_root)`` of the two GraphConvs.  Returns
    ``(left[m,N], right[n,N], z16_s | None, z16_t | None)``, bit-identical to two ``conv_in_16`` calls."""
    rowptr, col, val, m = csr
    colptr, row_csc, val_csc, n = csc
    require_cuda(rowptr, colptr, x_s, x_t, *l2r, *r2l)
    x_s, x_t = _contig(x_s.float()), _contig(x_t.float())
    l2r = [_contig(t.float()) for t in l2r]
    r2l = [_contig(t.float()) for t in r2l]
    if x_s.shape[1] != 8 or x_t.shape[1] != 8:
        raise ValueError("conv_in_16_pair covers the reference's 8 + 8 input features")
    N = l2r[0].shape[0]
    dev = x_s.device
    left = torch.empty((m, N), dtype=out_dtype, device=dev)
    right = torch.empty((n, N), dtype=out_dtype, device=dev)
    z_s = torch.empty((m, 64), dtype=out_dtype, device=dev) if want_z16 else None
    z_t = torch.empty((n, 64), dtype=out_dtype, device=dev) if want_z16 else None
    with torch.cuda.device(dev):
        rc = _lib.load().lpgnn_conv_in_16_pair(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), colptr.data_ptr(),
                                               row_csc.data_ptr(), val_csc.data_ptr(), m, n, x_s.data_ptr(), x_t.data_ptr(),
                                               l2r[0].data_ptr(), l2r[1].data_ptr(), l2r[2].data_ptr(), r2l[0].data_ptr(),
                                               r2l[1].data_ptr(), r2l[2].data_ptr(), N, left.data_ptr(), right.data_ptr(),
                                               dtype_code(out_dtype), EPI_RELU if relu else EPI_NONE, ptr(z_s), ptr(z_t),
                                               stream_ptr())
    check(rc, "lpgnn_conv_in_16_pair")
    return left, right, z_s, z_t


def gather_cat(view, x_src, x_dst, want_f32=True, want_bf16=False, dtype16=torch.bfloat16):
    """``z = [A_view @ x_src | x_dst | 0]``: fp32 ``[rows,KT]`` and/or 16-bit ``[rows,64]`` (``dtype16``: bf16 or
    half; input of the tensor-core transform in the 16-bit modes)."""
    ptr_, idx, val, rows = view
    require_cuda(ptr_, x_src, x_dst)
    x_src, x_dst = _contig(x_src.float()), _contig(x_dst.float())
    lib = _lib.load()
    kt = lib.lpgnn_conv_in_zcat_width(x_src.shape[1], x_dst.shape[1])
    z32 = torch.empty((rows, kt), dtype=torch.float32, device=x_src.device) if want_f32 else None
    zb = torch.empty((rows, 64), dtype=dtype16, device=x_src.device) if want_bf16 else None
    with torch.cuda.device(x_src.device):
        rc = lib.lpgnn_gather_cat_ex(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x_src.data_ptr(), x_src.shape[1],
                                     x_dst.data_ptr(), x_dst.shape[1], ptr(z32), ptr(zb), dtype_code(dtype16), stream_ptr())
    check(rc, "lpgnn_gather_cat_ex")
    return z32, zb


def node_transform(a1, w1, a2=None, w2=None, bias=None, relu=False, out_dtype=None, dropout=None, mask=None) -> torch.Tensor:
    """``epi(a1 @ w1.T + a2 @ w2.T + bias)``: bf16 operands -> tcgen05 kernel, fp32 -> CUDA-core
    kernel.  Replaces ``lin_rel(agg) + lin_root(x_dst)`` (+ relu_) (reference arch.py:75-80, 188).
    Training epilogues (``lpgnn_node_transform_ex``): ``dropout=(p, seed)`` applies inverted dropout to the output,
    ``mask=(act, scale)`` multiplies it by ``scale * (act > 0)``."""
    require_cuda(a1, w1, a2, w2, bias)
    dt = a1.dtype
    a1, w1 = _contig(a1), _contig(w1)
    if w1.dtype != dt:
        raise TypeError(f"node_transform: weight dtype {w1.dtype} != activation dtype {dt}")
    M, K1 = a1.shape
    N = w1.shape[0]
    K2 = 0
    if a2 is not None:
        a2, w2 = _contig(a2), _contig(w2)
        K2 = a2.shape[1]
        if a2.dtype != dt or w2.dtype != dt:
            raise TypeError("node_transform: all operands must share one dtype")
    if bias is not None:
        bias = _contig(bias.float())
    out_dtype = dt if out_dtype is None else out_dtype
    out = torch.empty((M, N), dtype=out_dtype, device=a1.device)
    if (dropout is not None and dropout[0] > 0) or mask is not None:
        import ctypes as C
        if out_dtype != dt:
            raise TypeError("node_transform: dropout / mask epilogues write the operand dtype")
        epi = _lib.EpilogueArgs()
        epi.epilogue = EPI_RELU if relu else EPI_NONE
        if dropout is not None:
            epi.dropout_p, epi.dropout_seed = float(dropout[0]), int(dropout[1])
        if mask is not None:
            act = mask[0]
            require_cuda(act)
            if act.shape != out.shape or act.dtype != dt or not act.is_contiguous():
                raise ValueError("node_transform: mask activation must be a contiguous tensor of the output's shape and dtype")
            epi.mask_act, epi.mask_scale = act.data_ptr(), float(mask[1])
        with torch.cuda.device(a1.device):
            rc = _lib.load().lpgnn_node_transform_ex(a1.data_ptr(), K1, w1.data_ptr(), ptr(a2), K2, ptr(w2), ptr(bias), M, N,
                                                     out.data_ptr(), dtype_code(dt), C.byref(epi), stream_ptr())
        check(rc, "lpgnn_node_transform_ex")
        return out
    with torch.cuda.device(a1.device):
        rc = _lib.load().lpgnn_node_transform(a1.data_ptr(), K1, w1.data_ptr(), ptr(a2), K2, ptr(w2), ptr(bias), M, N,
                                              out.data_ptr(), dtype_code(dt), dtype_code(out_dtype),
                                              EPI_RELU if relu else EPI_NONE, stream_ptr())
    check(rc, "lpgnn_node_transform")
    return out


def node_transform_head(a1, w1, a2, w2, bias, head_w, head_b, feas, relu=True, want_out=False):
    """Last hidden transform fused with the basis-status head and knowledge masking (bf16 tensor-core kernel):
    ``logits = add_knowledge(relu(a1 w1^T + a2 w2^T + bias) head_w^T + head_b)`` without writing the hidden
    activation (unless ``want_out``).  Returns ``(logits[M,3] f32, out | None)``."""
    require_cuda(a1, w1, a2, w2, bias, head_w, head_b, feas)
    a1, w1, a2, w2 = _contig(a1), _contig(w1), _contig(a2), _contig(w2)
    bias, head_w, head_b, feas = _contig(bias.float()), _contig(head_w.float()), _contig(head_b.float()), _contig(feas.float())
    M, K1 = a1.shape
    N, K2 = w1.shape[0], a2.shape[1]
    lib = _lib.load()
    nparts = lib.lpgnn_node_transform_head_parts(N)
    partial = torch.empty((nparts, M, 3), dtype=torch.float32, device=a1.device)
    out = torch.empty((M, N), dtype=a1.dtype, device=a1.device) if want_out else None
    logits = torch.empty((M, 3), dtype=torch.float32, device=a1.device)
    with torch.cuda.device(a1.device):
        rc = lib.lpgnn_node_transform_head_ex(a1.data_ptr(), K1, w1.data_ptr(), a2.data_ptr(), K2, w2.data_ptr(),
                                              bias.data_ptr(), M, N, ptr(out), dtype_code(a1.dtype),
                                              EPI_RELU if relu else EPI_NONE, head_w.data_ptr(), partial.data_ptr(),
                                              stream_ptr())
        check(rc, "lpgnn_node_transform_head_ex")
        rc = lib.lpgnn_head_finish(partial.data_ptr(), nparts, M, head_b.data_ptr(), feas.data_ptr(), feas.shape[1],
                                   logits.data_ptr(), stream_ptr())
    check(rc, "lpgnn_head_finish")
    return logits, out


def node_transform_head_train(a1, w1, a2, w2, bias, head_w, head_b, feas, relu=True, dropout=None):
    """Training forward of the last hidden layer (bf16): ``out = dropout(relu(a1 w1^T + a2 w2^T + bias))`` stored in 16 bits,
    with the basis-status head accumulated in the same epilogue on the stored (post-dropout) values.  ``dropout=(p, seed)``
    or None.  Returns ``(out[M,N], logits[M,3], raw[M,3])`` -- ``raw`` = un-normalised logits for ``head_mask_bwd``."""
    import ctypes as C
    require_cuda(a1, w1, a2, w2, bias, head_w, head_b, feas)
    a1, w1, a2, w2 = _contig(a1), _contig(w1), _contig(a2), _contig(w2)
    bias, head_w, head_b, feas = _contig(bias.float()), _contig(head_w.float()), _contig(head_b.float()), _contig(feas.float())
    if a1.dtype != torch.bfloat16:
        raise TypeError("node_transform_head_train: bf16 operands (training mode of the tensor-core path)")
    M, K1 = a1.shape
    N, K2 = w1.shape[0], a2.shape[1]
    lib = _lib.load()
    nparts = lib.lpgnn_node_transform_head_parts(N)
    partial = torch.empty((nparts, M, 3), dtype=torch.float32, device=a1.device)
    out = torch.empty((M, N), dtype=a1.dtype, device=a1.device)
    logits = torch.empty((M, 3), dtype=torch.float32, device=a1.device)
    raw = torch.empty((M, 3), dtype=torch.float32, device=a1.device)
    epi = _lib.EpilogueArgs()
    epi.epilogue = EPI_RELU if relu else EPI_NONE
    epi.dropout_p, epi.dropout_seed = (float(dropout[0]), int(dropout[1]) & (2 ** 64 - 1)) if dropout else (0.0, 0)
    epi.mask_act, epi.mask_scale = None, 1.0
    with torch.cuda.device(a1.device):
        rc = lib.lpgnn_node_transform_head_train(a1.data_ptr(), K1, w1.data_ptr(), a2.data_ptr(), K2, w2.data_ptr(),
                                                 bias.data_ptr(), M, N, out.data_ptr(), C.byref(epi), head_w.data_ptr(),
                                                 partial.data_ptr(), stream_ptr())
        check(rc, "lpgnn_node_transform_head_train")
        rc = lib.lpgnn_head_finish_ex(partial.data_ptr(), nparts, M, head_b.data_ptr(), feas.data_ptr(), feas.shape[1],
                                      logits.data_ptr(), raw.data_ptr(), stream_ptr())
    check(rc, "lpgnn_head_finish_ex")
    return out, logits, raw


def _ptr_array(tensors):
    import ctypes as C
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def split_bf16(x, parts=3):
    """fp32 -> ``parts`` bf16 tensors whose sum reproduces x to ~2^-17 (parts=2) / ~2^-25 (parts=3) relative
    (operands of ``node_transform_split``)."""
    require_cuda(x)
    x = _contig(x.float())
    outs = tuple(torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) for _ in range(parts))
    with torch.cuda.device(x.device):
        rc = _lib.load().lpgnn_split_bf16(x.data_ptr(), x.numel(), parts, _ptr_array(outs), stream_ptr())
    check(rc, "lpgnn_split_bf16")
    return outs


def node_transform_split(a1, w1, a2=None, w2=None, bias=None, relu=False):
    """fp32-accurate transform on the tensor cores: ``a*`` / ``w*`` are tuples of bf16 parts (``split_bf16``);
    returns fp32 ``epi(a1 w1^T + a2 w2^T + bias)``."""
    parts = len(a1)
    require_cuda(*a1, *w1, bias)
    M, K1 = a1[0].shape
    N = w1[0].shape[0]
    K2 = a2[0].shape[1] if a2 is not None else 0
    if bias is not None:
        bias = _contig(bias.float())
    out = torch.empty((M, N), dtype=torch.float32, device=a1[0].device)
    with torch.cuda.device(a1[0].device):
        rc = _lib.load().lpgnn_node_transform_split(parts, _ptr_array(a1), K1, _ptr_array(w1),
                                                    _ptr_array(a2) if a2 is not None else None, K2,
                                                    _ptr_array(w2) if w2 is not None else None, ptr(bias), M, N,
                                                    out.data_ptr(), EPI_RELU if relu else EPI_NONE, stream_ptr())
    check(rc, "lpgnn_node_transform_split")
    return out


def split_x2(x1, x2=None):
    """fp32 rows -> x2 operands of ``node_transform_x2``: ``x[i,k] = scale[i] * (hi[i,k] + 2^-11 lo[i,k])`` with IEEE-half
    ``hi`` / ``lo`` and one power-of-two scale per row shared by ``x1 [rows,K1]`` and ``x2 [rows,K2]``.
    Returns ``((hi1, lo1), (hi2, lo2) | None, scale[rows])``."""
    require_cuda(x1, x2)
    x1 = _contig(x1.float())
    rows, K1 = x1.shape
    K2 = 0
    if x2 is not None:
        x2 = _contig(x2.float())
        if x2.shape[0] != rows:
            raise ValueError("split_x2: x1 and x2 must have the same number of rows")
        K2 = x2.shape[1]
    h = lambda K: torch.empty((rows, K), dtype=torch.float16, device=x1.device)
    p1 = (h(K1), h(K1))
    p2 = (h(K2), h(K2)) if x2 is not None else None
    scale = torch.empty(rows, dtype=torch.float32, device=x1.device)
    with torch.cuda.device(x1.device):
        rc = _lib.load().lpgnn_split_x2(x1.data_ptr(), K1, ptr(x2), K2, rows, p1[0].data_ptr(), p1[1].data_ptr(),
                                        ptr(p2[0]) if p2 else None, ptr(p2[1]) if p2 else None, scale.data_ptr(),
                                        stream_ptr())
    check(rc, "lpgnn_split_x2")
    return p1, p2, scale


def node_transform_x2(a1, w1, a2=None, w2=None, rowscale=None, colscale=None, bias=None, relu=False, head=None,
                      want_out=True, rowscale2=None):
    """fp32-accurate ``epi(a1 w1^T + a2 w2^T + bias)`` on the tensor cores from x2 operands (``split_x2`` pairs
    ``(hi, lo)``; three half x half passes, chunked accumulation: csrc/gemm_x2.cu).  ``rowscale`` belongs to ``a1``,
    ``rowscale2`` to ``a2`` (default: the same array, operands split together).  ``head=(head_w, head_b, feas)``
    fuses the basis-status head + knowledge masking of the last layer and returns ``(out | None, logits[M,3])``."""
    require_cuda(*a1, *w1, rowscale, colscale, bias)
    M, K1 = a1[0].shape
    N = w1[0].shape[0]
    K2 = a2[0].shape[1] if a2 is not None else 0
    dev = a1[0].device
    if bias is not None:
        bias = _contig(bias.float())
    lib = _lib.load()
    out = torch.empty((M, N), dtype=torch.float32, device=dev) if want_out else None
    partial = head_w = None
    if head is not None:
        head_w, head_b, feas = (_contig(t.float()) for t in head)
        require_cuda(head_w, head_b, feas)
        nparts = lib.lpgnn_node_transform_head_parts(N)
        partial = torch.empty((nparts, M, 3), dtype=torch.float32, device=dev)
    elif not want_out:
        raise ValueError("node_transform_x2: nothing to compute (no output, no head)")
    with torch.cuda.device(dev):
        rc = lib.lpgnn_node_transform_x2(a1[0].data_ptr(), a1[1].data_ptr(), K1, w1[0].data_ptr(), w1[1].data_ptr(),
                                         ptr(a2[0]) if a2 else None, ptr(a2[1]) if a2 else None, K2,
                                         ptr(w2[0]) if w2 else None, ptr(w2[1]) if w2 else None, ptr(rowscale),
                                         ptr(rowscale2), ptr(colscale), ptr(bias), M, N, ptr(out),
                                         EPI_RELU if relu else EPI_NONE,
                                         ptr(head_w), ptr(partial), stream_ptr())
        check(rc, "lpgnn_node_transform_x2")
        if head is None:
            return out
        logits = torch.empty((M, 3), dtype=torch.float32, device=dev)
        rc = lib.lpgnn_head_finish(partial.data_ptr(), nparts, M, head_b.data_ptr(), feas.data_ptr(), feas.shape[1],
                                   logits.data_ptr(), stream_ptr())
    check(rc, "lpgnn_head_finish")
    return out, logits


def conv_in_fused_x2(view, x_src, x_dst, w_rel, b_rel, w_root, relu=True):
    """fp32 input layer (``conv_in_fused``) that also emits its output as x2 operands.  Returns ``(out[rows,N] f32,
    (hi, lo), scale[rows])`` with ``|out[r,:]| <= scale[r] * 2^12`` (scale from an a-priori bound, ``lpgnn_conv_in_fused_x2``)."""
    ptr_, idx, val, rows = view
    require_cuda(ptr_, x_src, x_dst, w_rel, w_root)
    x_src, x_dst = _contig(x_src.float()), _contig(x_dst.float())
    w_rel, w_root, b_rel = _contig(w_rel.float()), _contig(w_root.float()), _contig(b_rel.float())
    N = w_rel.shape[0]
    lib = _lib.load()
    dev = x_src.device
    kt = lib.lpgnn_conv_in_zcat_width(x_src.shape[1], x_dst.shape[1])
    out = torch.empty((rows, N), dtype=torch.float32, device=dev)
    z_cat = torch.empty((rows, kt), dtype=torch.float32, device=dev)
    hi, lo = (torch.empty((rows, N), dtype=torch.float16, device=dev) for _ in range(2))
    scale = torch.empty(rows, dtype=torch.float32, device=dev)
    wabs = torch.empty(80, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.lpgnn_conv_in_fused_x2(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x_src.data_ptr(), x_src.shape[1],
                                        x_dst.data_ptr(), x_dst.shape[1], w_rel.data_ptr(), b_rel.data_ptr(), w_root.data_ptr(),
                                        N, out.data_ptr(), EPI_RELU if relu else EPI_NONE, z_cat.data_ptr(), hi.data_ptr(),
                                        lo.data_ptr(), scale.data_ptr(), wabs.data_ptr(), stream_ptr())
    check(rc, "lpgnn_conv_in_fused_x2")
    return out, (hi, lo), scale


def spmm_x2(view, x, src_scale):
    """Aggregation of fp32 features straight into x2 operands (``lpgnn_spmm_x2``): ``((hi, lo), scale[rows])`` with
    ``A_view @ x = scale * (hi + 2^-11 lo)`` to 22 bits; ``src_scale[j]`` bounds source row j as ``conv_in_fused_x2`` reports."""
    ptr_, idx, val, rows = view
    require_cuda(ptr_, x, src_scale)
    x, src_scale = _contig(x.float()), _contig(src_scale.float())
    F = x.shape[1]
    dev = x.device
    hi, lo = (torch.empty((rows, F), dtype=torch.float16, device=dev) for _ in range(2))
    scale = torch.empty(rows, dtype=torch.float32, device=dev)
    scratch = torch.empty((rows, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().lpgnn_spmm_x2(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x.data_ptr(), F, src_scale.data_ptr(),
                                       hi.data_ptr(), lo.data_ptr(), scale.data_ptr(), scratch.data_ptr(), stream_ptr())
    check(rc, "lpgnn_spmm_x2")
    return (hi, lo), scale


def spmm_x2_pair(csr, csc, left, right, scale_left, scale_right, nnz=-1):
    """``spmm_x2`` for both directions in one launch (``lpgnn_spmm_x2_pair``).  Returns
    ``(((hi_s, lo_s), scale_s), ((hi_t, lo_t), scale_t))`` for ``A @ right`` and ``A^T @ left``."""
    rowptr, col, val, m = csr
    colptr, row_csc, val_csc, n = csc
    require_cuda(rowptr, colptr, left, right, scale_left, scale_right)
    left, right = _contig(left.float()), _contig(right.float())
    scale_left, scale_right = _contig(scale_left.float()), _contig(scale_right.float())
    F = left.shape[1]
    dev = left.device
    mk = lambda rows: (torch.empty((rows, F), dtype=torch.float16, device=dev), torch.empty((rows, F), dtype=torch.float16, device=dev),
                       torch.empty(rows, dtype=torch.float32, device=dev), torch.empty((rows, F), dtype=torch.float32, device=dev))
    hi_s, lo_s, sc_s, tmp_s = mk(m)
    hi_t, lo_t, sc_t, tmp_t = mk(n)
    with torch.cuda.device(dev):
        rc = _lib.load().lpgnn_spmm_x2_pair(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), m, colptr.data_ptr(), row_csc.data_ptr(),
                                            val_csc.data_ptr(), n, int(nnz), left.data_ptr(), right.data_ptr(), F, scale_left.data_ptr(),
                                            scale_right.data_ptr(), hi_s.data_ptr(), lo_s.data_ptr(), sc_s.data_ptr(), hi_t.data_ptr(),
                                            lo_t.data_ptr(), sc_t.data_ptr(), tmp_s.data_ptr(), tmp_t.data_ptr(), stream_ptr())
    check(rc, "lpgnn_spmm_x2_pair")
    return ((hi_s, lo_s), sc_s), ((hi_t, lo_t), sc_t)


def set_x2_chunk(kblocks: int) -> int:
    """Tuning knob of ``node_transform_x2`` (K-blocks of 64 per TMEM chunk); returns the previous value."""
    return _lib.load().lpgnn_set_x2_chunk(int(kblocks))


def head_mask(h, w, b, feas, want_raw=False):
    """Linear(H,3) + add_knowledge in one pass (reference arch.py:190-191, 129-141).
    Returns ``(logits[rows,3] f32, raw[rows,3] f32 | None)``."""
    require_cuda(h, w, b, feas)
    h, feas = _contig(h), _contig(feas.float())
    w, b = _contig(w.float()), _contig(b.float())
    rows, H = h.shape
    logits = torch.empty((rows, 3), dtype=torch.float32, device=h.device)
    raw = torch.empty((rows, 3), dtype=torch.float32, device=h.device) if want_raw else None
    with torch.cuda.device(h.device):
        rc = _lib.load().lpgnn_head_mask(h.data_ptr(), dtype_code(h.dtype), rows, H, w.data_ptr(), b.data_ptr(),
                                         feas.data_ptr(), feas.shape[1], logits.data_ptr(), ptr(raw), stream_ptr())
    check(rc, "lpgnn_head_mask")
    return logits, raw


def add_knowledge_kernel(logits, feas):
    """add_knowledge on given logits (reference arch.py:129-141)."""
    require_cuda(logits, feas)
    logits, feas = _contig(logits.float()), _contig(feas.float())
    out = torch.empty_like(logits)
    with torch.cuda.device(logits.device):
        rc = _lib.load().lpgnn_add_knowledge(logits.data_ptr(), logits.shape[0], feas.data_ptr(), feas.shape[1],
                                             out.data_ptr(), stream_ptr())
    check(rc, "lpgnn_add_knowledge")
    return out


def basis_select(logits_cons, logits_vars, k_basic=None, int64=True, want_counts=False):
    """``val.inference_gnn`` on the device (reference val.py:106-124): status per node, constraints
    first.  ``k_basic`` defaults to m (the reference's top-m rule)."""
    require_cuda(logits_cons, logits_vars)
    lc, lv = _contig(logits_cons.float()), _contig(logits_vars.float())
    m, n = lc.shape[0], lv.shape[0]
    k = m if k_basic is None else int(k_basic)
    dev = lc.device
    status = torch.empty(m + n, dtype=torch.int64 if int64 else torch.uint8, device=dev)
    counts = torch.empty(4, dtype=torch.int32, device=dev) if want_counts else None
    lib = _lib.load()
    ws_bytes = lib.lpgnn_basis_select_workspace_bytes(m + n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.lpgnn_basis_select(lc.data_ptr(), m, lv.data_ptr(), n, k, status.data_ptr(), 1 if int64 else 0,
                                    ptr(counts), ws.data_ptr(), ws_bytes, stream_ptr())
    check(rc, "lpgnn_basis_select")
    return (status, counts) if want_counts else status


# ------------------------------------------------------------------------------------------------ backward ops
def gemm_tn(a, b):
    """``a @ b.T`` -> fp32, for bf16 ``a [M,K]``, ``b [N,K]`` with a long K (split-K tensor-core GEMM);
    fp32 operands go through the CUDA-core kernel."""
    require_cuda(a, b)
    a, b = _contig(a), _contig(b)
    if a.dtype != torch.bfloat16:
        return node_transform(a, b, out_dtype=torch.float32)
    M, K = a.shape
    N = b.shape[0]
    lib = _lib.load()
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ws_bytes = lib.lpgnn_gemm_tn_workspace_bytes(M, N, K)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.lpgnn_gemm_tn(a.data_ptr(), b.data_ptr(), M, N, K, out.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
    check(rc, "lpgnn_gemm_tn")
    return out


def wgrad(dy, x):
    """``dy^T x`` -> fp32 ``[N_out, K_in]`` for bf16 row-major ``dy [Mn, N_out]``, ``x [Mn, K_in]`` (MN-major
    tensor-core operands, split-K over the nodes; no transposed copies)."""
    require_cuda(dy, x)
    dy, x = _contig(dy), _contig(x)
    Mn, N = dy.shape
    K = x.shape[1]
    lib = _lib.load()
    out = torch.empty((N, K), dtype=torch.float32, device=dy.device)
    ws_bytes = lib.lpgnn_wgrad_workspace_bytes(Mn, N, K)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dy.device)
    with torch.cuda.device(dy.device):
        rc = lib.lpgnn_wgrad(dy.data_ptr(), x.data_ptr(), Mn, N, K, out.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
    check(rc, "lpgnn_wgrad")
    return out


def head_mask_bwd(dlogits, raw, h_act, w, scale=1.0, want_bf16=False, want_colsum=False, want_bias_grad=False):
    """Backward of ``head_mask`` wrt the hidden activation, fused with its ReLU/dropout mask.
    Returns ``(dH[rows,H], draw[rows,3])`` [+ ``draw_bf16 [rows,64] = [draw | 0]``, the ``wgrad`` operand]
    [+ ``colsum(dH) [H]`` accumulated in fp32 before the output rounding: the bias gradient of the layer under the head]
    [+ ``colsum(draw) [3]``: the head's own bias gradient]."""
    require_cuda(dlogits, raw, h_act, w)
    dlogits, raw, h_act, w = _contig(dlogits.float()), _contig(raw), _contig(h_act), _contig(w.float())
    rows, H = h_act.shape
    lib = _lib.load()
    dH = torch.empty_like(h_act)
    draw = torch.empty((rows, 3), dtype=torch.float32, device=h_act.device)
    draw_b = torch.empty((rows, 64), dtype=torch.bfloat16, device=h_act.device) if want_bf16 else None
    cs = db = ws = None
    ws_bytes = 0
    if want_colsum or want_bias_grad:
        cs = torch.empty(H, dtype=torch.float32, device=h_act.device) if want_colsum else None
        db = torch.empty(3, dtype=torch.float32, device=h_act.device) if want_bias_grad else None
        ws_bytes = lib.lpgnn_head_mask_bwd_colsum_workspace_bytes(rows, H)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=h_act.device)
    with torch.cuda.device(h_act.device):
        rc = lib.lpgnn_head_mask_bwd_colsum(dlogits.data_ptr(), raw.data_ptr(), h_act.data_ptr(), dtype_code(h_act.dtype),
                                            rows, H, w.data_ptr(), float(scale), dH.data_ptr(), draw.data_ptr(),
                                            _lib.ptr(draw_b), _lib.ptr(cs), _lib.ptr(db), _lib.ptr(ws), ws_bytes, stream_ptr())
    check(rc, "lpgnn_head_mask_bwd_colsum")
    out = (dH, draw)
    if want_bf16:
        out += (draw_b,)
    if want_colsum:
        out += (cs,)
    if want_bias_grad:
        out += (db,)
    return out


def relu_bwd(a, b, act, scale=1.0, out=None):
    """``(a [+ b]) * scale * (act > 0)``; ``out`` may alias ``a``."""
    require_cuda(a, b, act)
    a, act = _contig(a), _contig(act)
    if b is not None:
        b = _contig(b)
    out = torch.empty_like(a) if out is None else out
    with torch.cuda.device(a.device):
        rc = _lib.load().lpgnn_relu_bwd(a.data_ptr(), ptr(b), act.data_ptr(), a.numel(), dtype_code(a.dtype), float(scale),
                                        out.data_ptr(), stream_ptr())
    check(rc, "lpgnn_relu_bwd")
    return out


def dropout_(x, p, seed):
    """In-place inverted dropout (training only)."""
    require_cuda(x)
    if p <= 0.0:
        return x
    with torch.cuda.device(x.device):
        rc = _lib.load().lpgnn_dropout(x.data_ptr(), x.numel(), dtype_code(x.dtype), float(p), int(seed) & (2 ** 64 - 1),
                                       stream_ptr())
    check(rc, "lpgnn_dropout")
    return x


def transpose(x, pad_to=64, out=None):
    """``x^T`` as ``[N, ld]`` with ``ld = M`` rounded up to ``pad_to`` and zero-filled padding columns.
    ``out`` may be a row-slice ``[N, ld]`` of a larger buffer (the halves of a concatenated GEMM operand)."""
    require_cuda(x)
    x = _contig(x)
    M, N = x.shape
    ld = (M + pad_to - 1) // pad_to * pad_to
    if out is None:
        out = torch.empty((N, ld), dtype=x.dtype, device=x.device)
    elif out.shape != (N, ld) or not out.is_contiguous() or out.dtype != x.dtype:
        raise ValueError("transpose: out must be a contiguous [N, ld] tensor of the input dtype")
    with torch.cuda.device(x.device):
        rc = _lib.load().lpgnn_transpose(x.data_ptr(), dtype_code(x.dtype), M, N, out.data_ptr(), ld, stream_ptr())
    check(rc, "lpgnn_transpose")
    return out


def colsum(x):
    require_cuda(x)
    x = _contig(x)
    M, N = x.shape
    lib = _lib.load()
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    ws_bytes = lib.lpgnn_colsum_workspace_bytes(M, N)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.lpgnn_colsum(x.data_ptr(), dtype_code(x.dtype), M, N, out.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
    check(rc, "lpgnn_colsum")
    return out


def small_wgrad(dy, z, k, want_bias=False):
    """``dW[N,k] = dy^T z[:, :k]`` (z fp32, any row stride) and optionally ``dB[N] = colsum(dy)``."""
    require_cuda(dy, z)
    dy, z = _contig(dy), _contig(z.float())
    M, N = dy.shape
    lib = _lib.load()
    dW = torch.empty((N, k), dtype=torch.float32, device=dy.device)
    dB = torch.empty(N, dtype=torch.float32, device=dy.device) if want_bias else None
    ws_bytes = lib.lpgnn_small_wgrad_workspace_bytes(M, N, k)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dy.device)
    with torch.cuda.device(dy.device):
        rc = lib.lpgnn_small_wgrad(dy.data_ptr(), dtype_code(dy.dtype), z.data_ptr(), z.shape[1], k, M, N, dW.data_ptr(),
                                   ptr(dB), ws.data_ptr(), ws_bytes, stream_ptr())
    check(rc, "lpgnn_small_wgrad")
    return dW, dB


def lp_features(g, a_csr, c, b_l, b_u, l, u):
    """(f-2) ``dataset.scaling`` + ``dataset.cvt_to_features`` (reference dataset.py:23-96) on the device.
    ``g``: a built ``BipartiteCSR`` holding the STRUCTURE of the raw A (its values are overwritten with the scaled
    coefficients of both orientations); ``a_csr`` float64 raw values in canonical CSR order; ``c, l, u`` [n] and
    ``b_l, b_u`` [m] float64.  Returns ``(x_s[m,8], x_t[n,8], scaled)`` with ``scaled`` a dict of the float64
    scaled LP (``A`` in CSR order, ``c, b_l, b_u, l, u``)."""
    g._require_built()
    require_cuda(g.rowptr, a_csr, c, b_l, b_u, l, u)
    f64 = lambda t: _contig(t.to(torch.float64))
    a_csr, c, b_l, b_u, l, u = f64(a_csr), f64(c), f64(b_l), f64(b_u), f64(l), f64(u)
    m, n, z = g.m, g.n, g.nnz()
    if a_csr.shape[0] != z or c.shape[0] != n or l.shape[0] != n or u.shape[0] != n or b_l.shape[0] != m or b_u.shape[0] != m:
        raise ValueError("lp_features: array lengths do not match the graph (m, n, nnz)")
    dev = g.rowptr.device
    lib = _lib.load()
    x_s = torch.empty((m, 8), dtype=torch.float32, device=dev)
    x_t = torch.empty((n, 8), dtype=torch.float32, device=dev)
    out = torch.empty(z + 3 * n + 2 * m, dtype=torch.float64, device=dev)
    o_a, o_c, o_bl, o_bu, o_l, o_u = torch.split(out, [z, n, m, m, n, n])
    ws_bytes = lib.lpgnn_lp_features_workspace_bytes(z, m, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.lpgnn_lp_features(g.rowptr.data_ptr(), g.col.data_ptr(), g.colptr.data_ptr(), g.row_csc.data_ptr(),
                                   g.csr2csc.data_ptr(), a_csr.data_ptr(), c.data_ptr(), b_l.data_ptr(), b_u.data_ptr(),
                                   l.data_ptr(), u.data_ptr(), z, m, n, g.val.data_ptr(), g.val_csc.data_ptr(),
                                   x_s.data_ptr(), x_t.data_ptr(), o_a.data_ptr(), o_c.data_ptr(), o_bl.data_ptr(),
                                   o_bu.data_ptr(), o_l.data_ptr(), o_u.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
    check(rc, "lpgnn_lp_features")
    return x_s, x_t, {"A": o_a, "c": o_c, "b_l": o_bl, "b_u": o_bu, "l": o_l, "u": o_u}
