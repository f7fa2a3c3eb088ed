/*
 * lpgnn.h -- C ABI of liblpgnn.so: the B200 (sm_100a) kernels behind the lp-gnn hot path.
 *
 * This is the drop-in boundary.  The reference (vbdai/lp-gnn) is pure Python and reaches its
 * numerics through torch_geometric / torch_sparse / ATen; each entry point below replaces one
 * of those third-party calls AT THE REFERENCE CALL SITE cited next to it, so a maintainer of
 * the reference binds this library with ctypes and keeps arch.py's module
 * classes, forward(batch) and state_dict keys unchanged.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless the name ends in _host.  The library never
 *    allocates user-visible memory: outputs and workspaces are caller-allocated (sizes from the
 *    *_workspace_bytes queries).  All work is enqueued on `stream` (a cudaStream_t); nothing
 *    synchronises unless stated.
 *  - Row-major dense matrices, leading dimension = number of columns.  Feature matrices are
 *    float32 (LPGNN_F32), bfloat16 (LPGNN_BF16) or IEEE half (LPGNN_F16, inference); accumulation is always float32.
 *  - Graph indices are int32 (max(m, n, nnz) < 2^31; BASELINE C4 has nnz = 1e7).
 *  - Return value: 0 on success, a negative LPGNN_E* code otherwise; lpgnn_last_error() gives a
 *    thread-local message.  There is no CPU fallback anywhere: without an sm_100 device every
 *    call fails with LPGNN_ENODEVICE.
 */
#ifndef LPGNN_H_
#define LPGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPGNN_VERSION 100 /* 0.1.0 */

#define LPGNN_OK 0
#define LPGNN_EINVAL (-1)    /* bad argument (shape, alignment, null pointer) */
#define LPGNN_ECUDA (-2)     /* CUDA runtime / driver error, see lpgnn_last_error() */
#define LPGNN_ENODEVICE (-3) /* no CUDA device, or device is not sm_100 */
#define LPGNN_EWORKSPACE (-4)/* workspace too small */

#define LPGNN_F32 0
#define LPGNN_BF16 1
/* IEEE half storage with fp32 accumulation: the reference's `--fp16 1` mode (model.half() val.py:269,
 * scripts/pred_basis.py:146; batch_to utils.py:909-915).  Same tensor-core rate as bf16 with 3 more mantissa bits (basis
 * statuses agree with fp32 on >= 99.9 % of the nodes where bf16 reaches ~99.8 %).  Inference entry points only
 * (spmm, gather_cat_ex, node_transform, node_transform_head_ex, head_mask, predict_basis*): the training-only
 * calls reject it. */
#define LPGNN_F16 2

/* epilogue flags of the node transforms */
/* lpgnn_graph_build flags */
#define LPGNN_COO_SORTED 1
#define LPGNN_GRAPH_MEAN 4   /* degree normalisation: val /= deg(row), val_csc /= deg(column) (mean aggregation) */
/* lpgnn_predict_basis_packed only: statuses are written per LP, [constraints of LP b | variables of LP b] back to back
 * (LP b starts at cons_ptr[b] + vars_ptr[b]), instead of [all constraints | all variables] */
#define LPGNN_STATUS_LP_MAJOR 8

#define LPGNN_EPI_NONE 0
#define LPGNN_EPI_RELU 1

typedef void* lpgnn_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define LPGNN_API __attribute__((visibility("default")))
#else
#define LPGNN_API
#endif

LPGNN_API int lpgnn_version(void);
LPGNN_API const char* lpgnn_last_error(void);
/* Number of CUDA kernels this library has enqueued so far in this process (all streams). */
LPGNN_API uint64_t lpgnn_launch_count(void);
/* Fills SM count and compute capability of the current device. */
LPGNN_API int lpgnn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * (a1) Graph construction.  Replaces torch_sparse.SparseTensor.from_edge_index
 * (reference dataset.py:301-304) and SparseTensor.t() (reference arch.py:71).
 *
 * Input: COO of the m x n LP matrix A in any order (coo_row[e] in [0,m), coo_col[e] in [0,n),
 * int64 if idx_is_i64 else int32).  Output: canonical CSR (stable sort by row, then column:
 * row-major, ascending column inside a row, duplicates kept in input order) and the CSC view
 * (stable sort of the CSR entries by column: column-major, ascending row inside a column) with
 * the permutation csr2csc (val_csc[k] = val[csr2csc[k]]).  Bit-exact with the reference's
 * ordering; all integer work, deterministic (no atomics on ordered data).
 *
 * flags: LPGNN_COO_SORTED = the caller asserts the COO is already in canonical (row, col) order
 * (torch_sparse's `is_sorted=True`; true for everything LPDataset.get produces, dataset.py:251-252),
 * which skips the COO sort.  status (device int32, optional, ZERO-INITIALISED BY THE CALLER) gets bits
 * OR-ed in by the kernels: bit 0 = the SORTED claim was false (outputs are then NOT canonical),
 * bit 1 = an index was out of range.  Written asynchronously; read it after synchronising the stream.
 * LPGNN_GRAPH_MEAN = degree normalisation (OFF in the reference: GraphConv aggr='add', arch.py:57,60 keep 'mean'
 * commented out): each orientation's values are divided by the degree of its destination node (IEEE division), so the
 * aggregation becomes the mean over the neighbours.  The two orientations then hold DIFFERENT values, i.e. the other
 * orientation is no longer the transpose: forward / inference only.
 * ------------------------------------------------------------------------------------------- */
LPGNN_API size_t lpgnn_graph_build_workspace_bytes(int64_t nnz, int32_t m, int32_t n);
LPGNN_API int lpgnn_graph_build(const void* coo_row, const void* coo_col, int idx_is_i64,
                      const float* coo_val, int64_t nnz, int32_t m, int32_t n, int flags,
                      int32_t* rowptr /*[m+1]*/, int32_t* col /*[nnz]*/, float* val /*[nnz]*/,
                      int32_t* colptr /*[n+1]*/, int32_t* row_csc /*[nnz]*/, float* val_csc /*[nnz]*/,
                      int32_t* csr2csc /*[nnz]*/, int32_t* status /*[1] device, optional*/,
                      void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* Tuning knob (process-wide): a LPGNN_COO_SORTED build runs as the plain chain (default: 2 + 3 launches per radix pass of
 * the CSC sort -- 8 at BASELINE C2) or, with enable = 1, as the compact chain (1 + 2 per pass: the preparation kernel
 * also counts the first digit and the columns, every scatter counts the next pass's digits, the last scatter writes the
 * CSC payload and colptr; measured no faster on B200, see csrc/graph_build.cu).  Bit-identical outputs.  Returns the
 * previous setting. */
LPGNN_API int lpgnn_set_graph_compact(int enable);

/* Tuning knob (process-wide): a LPGNN_COO_SORTED build runs as a chain of launches (default) or as ONE cooperative
 * launch (enable = 1: prep, the radix passes of the CSC sort and the finish, separated by grid-wide barriers; measured no
 * faster on B200, see csrc/graph_build.cu).  Bit-identical outputs.  Returns the previous setting. */
LPGNN_API int lpgnn_set_graph_fused(int enable);

/* `count` independent pinned-host -> device copies enqueued on `stream` by one native call (addresses and
 * sizes as host arrays of uint64).  Used to stage a pack of LPs: each LP's arrays go straight to their
 * offsets in the pack.  Replaces the per-LP `batch_to(batch, dev)` of the reference's prediction sweep
 * (val.py:31 via scripts/pred_basis.py:153-154, and pred_basis.py:170; utils.py:909-915: five tensor moves per LP). */
LPGNN_API int lpgnn_copy_many_h2d(const uint64_t* dst_ptrs, const uint64_t* src_ptrs, const uint64_t* nbytes,
                        int32_t count, lpgnn_stream_t stream);

/* Block-diagonal packing of several LPs into one graph (the direct sum of their matrices; what a PyG DataLoader
 * with batch_size > 1 would produce -- the reference runs batch_size=1, scripts/pred_basis.py:138): the COO entries
 * of LP b occupy [edge_ptr[b], edge_ptr[b+1]) with LP-local indices; this shifts them in place by the LP's
 * first constraint (cons_ptr[b]) / variable (vars_ptr[b]) of the pack.  A pack of row-major sorted LPs is
 * row-major sorted. */
LPGNN_API int lpgnn_pack_offsets(int32_t* row, int32_t* col, int64_t nnz, const int32_t* edge_ptr,
                       const int32_t* cons_ptr, const int32_t* vars_ptr, int32_t n_segments,
                       lpgnn_stream_t stream);

/* Same job for LPs that were STAGED whole: `staged` holds, LP after LP, each LP's host pack [row | col | val | x_s | x_t]
 * (4-byte words; LP b starts at word stage_off[b], stage_off [n_segments + 1]) exactly as one host -> device copy per LP
 * left it; one kernel moves every word to its place in the pack layout (row / col / val [nnz], x_s [M,p], x_t [N,q]) and
 * shifts the indices to the pack's numbering.  Replaces five copies per LP + lpgnn_pack_offsets (batch_to of the
 * reference's sweep, utils.py:909-915: five tensor moves per LP). */
LPGNN_API int lpgnn_pack_scatter(const int32_t* staged, const int32_t* stage_off, const int32_t* edge_ptr,
                       const int32_t* cons_ptr, const int32_t* vars_ptr, int32_t n_segments, int32_t p, int32_t q,
                       int64_t total_words, int32_t* row, int32_t* col, float* val, float* x_s, float* x_t,
                       lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (a2) Aggregation.  Replaces torch_sparse.matmul(adj_t, x, reduce='add') == spmm_sum, reached
 * from PyG GraphConv.message_and_aggregate (reference arch.py:75-80), forward AND backward
 * (backward wrt the dense operand is the same call on the other orientation).
 *
 *   Y[i, :] = sum_{e in [ptr[i], ptr[i+1])} val[e] * X[idx[e], :]     (e ascending, fp32 accumulate)
 *
 * X is [n_src, F], Y is [rows, F], both of `dtype`.  F * sizeof(elem) must be a multiple of 16.
 * Atomics-free and deterministic: one warp (or sub-warp) owns an output row.
 * ------------------------------------------------------------------------------------------- */
LPGNN_API int lpgnn_spmm(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
               const void* X, void* Y, int32_t F, int dtype, lpgnn_stream_t stream);
/* Same operation with explicit tuning (results are bit-identical for every setting).  Wide feature rows are
 * aggregated by a banded sweep: one CTA per SM owns one slab of the feature row and a contiguous range of output rows,
 * so the source rows shared by neighbouring output rows are served from the SM's L1 (see csrc/spmm.cu).
 * slab_bytes: 512 or 1024 (the row must be a multiple of it); 0 = automatic; -1 = the row-per-warp kernel.
 * unroll: neighbours gathered in flight per warp, 2 or 4; 0 = automatic. */
LPGNN_API int lpgnn_spmm_ex(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                  const void* X, void* Y, int32_t F, int dtype, int slab_bytes, int unroll,
                  lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (a2+a3, input layer) Fused aggregation + node transform for narrow inputs (conv1 of GCN_FC:
 * in = p or q = 8).  Replaces GraphConv.forward = lin_rel(spmm(adj_t, x_src)) + lin_root(x_dst)
 * (reference arch.py:75-80 with the (p,q)->hids layer built at arch.py:170) and the relu_ at
 * arch.py:182:
 *
 *   out[i,:] = epi( (sum_e val[e]*Xsrc[idx[e],:]) * W_rel^T + b_rel + Xdst[i,:] * W_root^T )
 *
 * Xsrc [n_src,k_src] f32, Xdst [rows,k_dst] f32, W_rel [N,k_src] f32, W_root [N,k_dst] f32,
 * b_rel [N] f32; k_src+k_dst <= 64; N even; out [rows,N] of out_dtype.  z_cat (required, f32
 * [rows, KT], KT = lpgnn_conv_in_zcat_width(k_src,k_dst) in {16,32,64}) receives the concatenated
 * transform input [aggregate | Xdst | 0-pad]: scratch for the forward pass, and exactly the operand
 * the weight gradient of this layer needs (lpgnn_small_wgrad).
 * ------------------------------------------------------------------------------------------- */
LPGNN_API int32_t lpgnn_conv_in_zcat_width(int32_t k_src, int32_t k_dst);
/* The gather half alone: z_cat (f32 [rows,KT], optional) and/or z_bf16 (bf16 [rows,64], optional: zero-padded
 * except column k_src+k_dst, which holds 1.0 when it exists -- its weight column is zero in the forward
 * transform, and the weight-gradient GEMM lpgnn_wgrad(dPre, z_bf16) finds the bias gradient there).  In bf16 mode the input layer is gather_cat -> lpgnn_node_transform(z_bf16, [W_rel|W_root|0]
 * as bf16 [N,64]): the tensor-core kernel with a single K block, bound by its epilogue (the output
 * write), which is ~4x faster than the CUDA-core transform. */
LPGNN_API int lpgnn_gather_cat(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                     const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst,
                     float* z_cat, void* z_bf16, lpgnn_stream_t stream);
/* Same with the 16-bit operand in z16_dtype (LPGNN_BF16 or LPGNN_F16). */
LPGNN_API int lpgnn_gather_cat_ex(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                        const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst,
                        float* z_cat, void* z16, int z16_dtype, lpgnn_stream_t stream);
LPGNN_API int lpgnn_conv_in_fused(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                        const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst,
                        const float* W_rel, const float* b_rel, const float* W_root, int32_t N,
                        void* out, int out_dtype, int epilogue, float* z_cat,
                        lpgnn_stream_t stream);

/* (a2+a3, input layer, 16-bit modes) conv1 of GCN_FC(8, 8, ...) in ONE kernel: out = epi( (A_view Xsrc) W_rel^T + b_rel +
 * Xdst W_root^T ) with Xsrc [n_src,8], Xdst [rows,8], W_rel / W_root [N,8] f32 and out [rows,N] of out_dtype (LPGNN_BF16
 * or LPGNN_F16; operands rounded to that type, fp32 accumulate).  The aggregate, the 16-wide product (one m16n8k16 MMA
 * step per 16 x 8 outputs, accumulators in registers, bias as their initial value), ReLU and the 16-byte stores share a
 * kernel whose only large traffic is the output write (csrc/conv_in_mma.cu).  z16 (optional, 16-bit [rows,64]) receives
 * [aggregate | Xdst | 1 | 0 ...]: the operand lpgnn_wgrad needs for this layer's weight / bias gradient.  N % 32 == 0.
 * Replaces GraphConv.forward + relu_ of the (p,q)->hids layer (reference arch.py:170, 75-80, 181-182). */
LPGNN_API int lpgnn_conv_in_16(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                     const float* Xsrc, const float* Xdst, const float* W_rel, const float* b_rel,
                     const float* W_root, int32_t N, void* out, int out_dtype, int epilogue, void* z16,
                     lpgnn_stream_t stream);
/* Both directions of that layer in ONE launch (arch.py:181: `left, right = self.conv1(x_left, x_right, edge_index)`):
 * out_t [n,N] = variables side (CSC view: colptr / row_csc / val_csc, sources x_s, own features x_t, l2r weights),
 * out_s [m,N] = constraints side (CSR view, sources x_t, own features x_s, r2l weights).  The launch's blocks are split
 * between the sides in proportion to their rows, so the second direction's weight staging and first gathers hide under
 * the first one's stores; results are bit-identical to two lpgnn_conv_in_16 calls.  z16_s / z16_t optional as above. */
LPGNN_API int lpgnn_conv_in_16_pair(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* colptr,
                          const int32_t* row_csc, const float* val_csc, int32_t m, int32_t n, const float* x_s,
                          const float* x_t, const float* l2r_wrel, const float* l2r_b, const float* l2r_wroot,
                          const float* r2l_wrel, const float* r2l_b, const float* r2l_wroot, int32_t N,
                          void* out_s, void* out_t, int out_dtype, int epilogue, void* z16_s, void* z16_t,
                          lpgnn_stream_t stream);
/* Tuning knob (process-wide): inputs of at least 2 x SM-count batches of 128 rows with N in {256, 512, 768, 1024} take a
 * warp-specialised form of the kernel (B fragments in registers, producer warps gathering the next batch) -- default on;
 * enable = 0 keeps every size on the shared-memory-B kernel.  Bit-identical results.  Returns the previous setting. */
LPGNN_API int lpgnn_set_conv_in_regb(int enable);

/* ---------------------------------------------------------------------------------------------
 * (a3) Dense node transform of a hidden GraphConv layer.  Replaces lin_rel(agg) + lin_root(x_dst)
 * (two torch Linear / cuBLAS sgemm calls + bias + add, PyG GraphConv.forward reached from
 * reference arch.py:75-80) and the following relu_ (arch.py:188) with ONE GEMM over the
 * concatenated reduction dimension:
 *
 *   out[M,N] = epi( A1[M,K1] * W1[N,K1]^T + A2[M,K2] * W2[N,K2]^T + bias[N] )
 *
 * dtype LPGNN_BF16 / LPGNN_F16: tcgen05 tensor-core kernel (TMA-fed, TMEM accumulators, fp32 accumulate);
 * operands are bf16 (or IEEE half, out_dtype = LPGNN_F16), out is bf16 or f32 (out_dtype; f32 is used by the weight-gradient GEMMs),
 * bias f32; K1,K2 multiples of 64, N multiple of 64 (A2/W2 may be NULL with K2 = 0).  dtype LPGNN_F32: fp32 CUDA-core kernel (the 1e-4 parity mode); operands,
 * bias and out are f32 (K1,K2 multiples of 4).
 * ------------------------------------------------------------------------------------------- */
LPGNN_API int lpgnn_node_transform(const void* A1, int32_t K1, const void* W1,
                         const void* A2, int32_t K2, const void* W2,
                         const float* bias, int32_t M, int32_t N,
                         void* out, int dtype, int out_dtype, int epilogue, lpgnn_stream_t stream);

/* lpgnn_node_transform with keep-masks fused into the epilogue (training; out is of `dtype`):
 *   out = keep ? epi(A1 W1^T + A2 W2^T + bias) * s : 0
 *   mask_act != NULL  keep &= (mask_act > 0), s *= mask_scale: the data gradient of a layer masked by the ReLU /
 *                     inverted-dropout pattern of that layer's input activation (reference arch.py:182,186-188
 *                     differentiated); mask_act has the shape and dtype of out.
 *   dropout_p > 0     keep &= lpgnn_dropout's hash of (dropout_seed, element index), s *= 1/(1-p): F.dropout
 *                     after the layer (arch.py:186-187; dropout and relu_ commute).
 * bf16: applied on the fp32 accumulators inside the tensor-core kernel, no extra pass over the activations.
 * fp32: the CUDA-core transform followed by lpgnn_relu_bwd / lpgnn_dropout (same result semantics). */
typedef struct lpgnn_epilogue_args {
  int32_t epilogue;        /* LPGNN_EPI_NONE | LPGNN_EPI_RELU */
  float dropout_p;
  uint64_t dropout_seed;
  const void* mask_act;
  float mask_scale;
} lpgnn_epilogue_args;
LPGNN_API int lpgnn_node_transform_ex(const void* A1, int32_t K1, const void* W1, const void* A2, int32_t K2,
                            const void* W2, const float* bias, int32_t M, int32_t N, void* out, int dtype,
                            const lpgnn_epilogue_args* epi, lpgnn_stream_t stream);

/* (a3, fp32 parity mode on the tensor cores) The same transform with fp32 accuracy from bf16 tensor-core
 * products: every fp32 operand x is given as `parts` bf16 tensors p0 = bf16(x), p1 = bf16(x - p0)
 * [, p2 = bf16(x - p0 - p1)] (lpgnn_split_bf16) and the significant cross products are accumulated in the
 * same fp32 TMEM tile: parts = 2 -> 3 passes (a0w0 + a0w1 + a1w0, ~2^-17 relative per product);
 * parts = 3 -> 6 passes (+ a0w2 + a2w0 + a1w1, ~2^-24: fp32 level; the mode that meets the 1e-4 logit bar).
 * A1/W1/A2/W2 are host arrays of `parts` device pointers; out is f32. */
LPGNN_API int lpgnn_split_bf16(const float* x, int64_t count, int parts, void* const* out_parts,
                     lpgnn_stream_t stream);
LPGNN_API int lpgnn_node_transform_split(int parts, const void* const* A1, int32_t K1, const void* const* W1,
                               const void* const* A2, int32_t K2, const void* const* W2,
                               const float* bias, int32_t M, int32_t N, float* out, int epilogue,
                               lpgnn_stream_t stream);

/* (a3, the reference's DEFAULT precision -- `--fp16 0`, utils.py:770 -- on the tensor cores)  fp32-accurate node
 * transform from "x2" operands.  lpgnn_split_x2 stores every fp32 row of x1 [rows,K1] (and, with the SAME row scale,
 * of x2 [rows,K2]; x2 may be NULL with K2 = 0) as two IEEE-half rows and a power-of-two scale:
 *     x[i,k] = scale[i] * ( hi[i,k] + 2^-11 * lo[i,k] )      (22 significant bits, no overflow for any fp32 row)
 * lpgnn_node_transform_x2 computes out = epi( A1 W1^T + A2 W2^T + bias ) from such operands (activations with row
 * scales `rowscale` [M] for A1 and `rowscale2` [M] for A2 -- NULL = the same array as A1's, for operands split together --,
 * weights with output-feature scales `colscale` [N] shared by W1 and W2; NULL = 1) as THREE half x half -> fp32
 * tcgen05 passes (hi*hi + 2^-11 (hi*lo + lo*hi)); a TMEM accumulator only holds a short chunk of the reduction and
 * the chunks are summed in fp32 registers with round-to-nearest, so the tensor core's truncating accumulation does
 * not drift (csrc/gemm_x2.cu).  out: f32 [M,N] or NULL; head_w [3,N] / head_partial [nparts][M][3] (optional, as
 * lpgnn_node_transform_head; finish with lpgnn_head_finish): the fused basis-status head of the last layer.
 * K1, K2, N multiples of 64.  Replaces lin_rel(agg) + lin_root(x) + relu_ (reference arch.py:75-80, 185-188). */
LPGNN_API int lpgnn_split_x2(const float* x1, int32_t K1, const float* x2, int32_t K2, int64_t rows,
                   void* hi1, void* lo1, void* hi2, void* lo2, float* scale, lpgnn_stream_t stream);
LPGNN_API int lpgnn_node_transform_x2(const void* A1_hi, const void* A1_lo, int32_t K1, const void* W1_hi, const void* W1_lo,
                            const void* A2_hi, const void* A2_lo, int32_t K2, const void* W2_hi, const void* W2_lo,
                            const float* rowscale, const float* rowscale2, const float* colscale, const float* bias,
                            int32_t M, int32_t N, float* out, int epilogue, const float* head_w, float* head_partial,
                            lpgnn_stream_t stream);
/* Producers that write x2 operands directly (no lpgnn_split_x2 pass over their fp32 output).  Both take their row scale
 * from an A-PRIORI bound instead of the row maximum -- any power of two s with |x[r,:]| <= s * 2^13 keeps the 22 bits --:
 *   lpgnn_conv_in_fused_x2   lpgnn_conv_in_fused (fp32 input layer, reference arch.py:170, 75-80, 181-182) that also emits
 *                            hi / lo [rows,N] and scale [rows] with |out[r,:]| <= scale[r] * 2^12 (bound: sum_k |z[r,k]| *
 *                            max_c |W[c,k]| + max |b|); wabs: float scratch [65]
 *   lpgnn_spmm_x2            lpgnn_spmm over fp32 features whose result is ONLY written as hi / lo [rows,F] + scale [rows]
 *                            (bound: sum_e |val[e]| * src_scale[idx[e]] * 2^12 with src_scale as written by the call
 *                            above); shapes too small for the banded sweep go through lpgnn_spmm into `scratch`
 *                            (fp32 [rows,F]) + lpgnn_split_x2 -- same outputs. */
LPGNN_API int lpgnn_conv_in_fused_x2(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                           const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst, const float* W_rel,
                           const float* b_rel, const float* W_root, int32_t N, float* out, int epilogue, float* z_cat,
                           void* hi, void* lo, float* scale, float* wabs, lpgnn_stream_t stream);
LPGNN_API int lpgnn_spmm_x2(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const float* X,
                  int32_t F, const float* src_scale, void* hi, void* lo, float* scale, float* scratch,
                  lpgnn_stream_t stream);
/* Both aggregations of a layer in ONE launch where both take the banded sweep (reference arch.py:71-80: the two
 * GraphConvs of a GraphConvTwoDirection read the same features): agg_t [n,F] = A^T L over the CSC view, agg_s [m,F] = A R
 * over the CSR view; each side gets the SMs its work (nnz + rows) asks for (nnz < 0: unknown to the caller, an LP-typical
 * density is assumed -- it only steers the split).  Other shapes: two lpgnn_spmm calls.  Same
 * bits either way.  lpgnn_spmm_x2_pair: the same for lpgnn_spmm_x2 (scale_L / scale_R = the sources' row-scale bounds). */
LPGNN_API int lpgnn_spmm_pair(const int32_t* rowptr, const int32_t* col, const float* val, int32_t m, const int32_t* colptr,
                    const int32_t* row_csc, const float* val_csc, int32_t n, int64_t nnz, const void* L,
                    const void* R, void* agg_s, void* agg_t, int32_t F, int dtype, lpgnn_stream_t stream);
LPGNN_API int lpgnn_spmm_x2_pair(const int32_t* rowptr, const int32_t* col, const float* val, int32_t m,
                       const int32_t* colptr, const int32_t* row_csc, const float* val_csc, int32_t n, int64_t nnz,
                       const float* L, const float* R, int32_t F, const float* scale_L, const float* scale_R,
                       void* hi_s, void* lo_s, float* scale_s, void* hi_t, void* lo_t, float* scale_t,
                       float* scratch_s, float* scratch_t, lpgnn_stream_t stream);
/* Tuning knob (process-wide): enable = 0 makes the two pair entry points above always run as two launches (A/B runs);
 * returns the previous setting. */
LPGNN_API int lpgnn_set_spmm_pair(int enable);
/* Tuning knob (process-wide): K-blocks of 64 accumulated inside TMEM before a chunk is added to the fp32 registers
 * (main passes; default 4; the 2^-11-weighted correction passes use 4x that, at least 16).  Returns the previous value. */
LPGNN_API int lpgnn_set_x2_chunk(int kblocks);

/* (a3+a4, inference) The LAST hidden transform fused with the basis-status head (reference
 * arch.py:185-190): the epilogue also forms, per row, the three dot products of the ReLU'd fp32
 * accumulator row with head_w [3,N] f32 over the tile's columns and writes them to
 * head_partial [nparts][M][3] f32 (nparts = lpgnn_node_transform_head_parts(N), one slice per column
 * tile; summed in a fixed order by lpgnn_head_finish).  out (bf16 [M,N]) may be NULL: inference
 * never needs the last hidden activation in HBM.  bf16 operands only. */
LPGNN_API int32_t lpgnn_node_transform_head_parts(int32_t N);
LPGNN_API int lpgnn_node_transform_head(const void* A1, int32_t K1, const void* W1,
                              const void* A2, int32_t K2, const void* W2,
                              const float* bias, int32_t M, int32_t N, void* out, int epilogue,
                              const float* head_w, float* head_partial, lpgnn_stream_t stream);
/* Same with the operand / output type given: dtype = LPGNN_BF16 or LPGNN_F16. */
LPGNN_API int lpgnn_node_transform_head_ex(const void* A1, int32_t K1, const void* W1,
                                 const void* A2, int32_t K2, const void* W2,
                                 const float* bias, int32_t M, int32_t N, void* out, int dtype, int epilogue,
                                 const float* head_w, float* head_partial, lpgnn_stream_t stream);
/* logits = add_knowledge(sum_p head_partial[p] + b): finishes the fused head (arch.py:190-191). */
LPGNN_API int lpgnn_head_finish(const float* head_partial, int32_t nparts, int32_t rows, const float* b,
                      const float* feas, int32_t q, float* logits, lpgnn_stream_t stream);
/* Same, also writing the un-normalised logits (raw_out [rows,3], may be null): what lpgnn_head_mask_bwd needs. */
LPGNN_API int lpgnn_head_finish_ex(const float* partial, int32_t nparts, int32_t rows, const float* b, const float* feas,
                         int32_t q, float* logits, float* raw_out, lpgnn_stream_t stream);
/* Training forward of the LAST hidden layer (bf16): lpgnn_node_transform_ex (ReLU + dropout + 16-bit store) with the head
 * of lpgnn_node_transform_head accumulated in the same epilogue on the values that are stored, i.e. after dropout
 * (reference arch.py:186-190: dropout, relu_, lin_left / lin_right).  head_partial as for lpgnn_node_transform_head. */
LPGNN_API int lpgnn_node_transform_head_train(const void* A1, int32_t K1, const void* W1, const void* A2, int32_t K2,
                                    const void* W2, const float* bias, int32_t M, int32_t N, void* out,
                                    const lpgnn_epilogue_args* epi, const float* head_w, float* head_partial,
                                    lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (a4+a5) Basis-status head + knowledge masking.  Replaces torch.nn.Linear(H,3) (reference
 * arch.py:190) and add_knowledge (reference arch.py:129-141):
 *
 *   raw[i,:]    = H[i,:] * W^T + b                       (W [3,Hdim] f32, b [3] f32)
 *   logits[i,:] = 10 * raw / max(||raw||_2, 1e-12)
 *   logits[i,0] -= 10 where feas[i, q-3] != 0 ;  logits[i,2] -= 10 where feas[i, q-1] != 0
 *
 * H [rows,Hdim] of h_dtype; feas [rows,q] f32 (the node features x_s / x_t); logits [rows,3] f32.
 * raw_out (optional, [rows,3] f32) keeps the un-normalised logits for the backward pass.
 * ------------------------------------------------------------------------------------------- */
LPGNN_API int lpgnn_head_mask(const void* H, int h_dtype, int32_t rows, int32_t Hdim,
                    const float* W, const float* b, const float* feas, int32_t q,
                    float* logits, float* raw_out, lpgnn_stream_t stream);

/* add_knowledge alone on given logits (reference arch.py:129-141; used by callers that own the
 * head GEMM, e.g. GCNRand-style baselines). in/out [rows,3] f32, may alias. */
LPGNN_API int lpgnn_add_knowledge(const float* logits_in, int32_t rows, const float* feas, int32_t q,
                        float* logits_out, lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (a6) Basis decision.  Replaces val.inference_gnn (reference val.py:106-124): softmax over the
 * 3 classes, NaN -> 0, the m largest P(basic) over all m+n nodes get status 1 (ties at the
 * threshold go to the lower node index), the others get 0 if p0 >= p2 else 2.
 * logits_cons [m,3] f32, logits_vars [n,3] f32 -> status [m+n] (int64 if status_is_i64 else
 * uint8), constraints first.  counts_out (optional, int32[4] device): number of nodes with
 * status 0,1,2 and number of basic variables (the two invariants asserted at val.py:118-122).
 * ------------------------------------------------------------------------------------------- */
LPGNN_API size_t lpgnn_basis_select_workspace_bytes(int64_t total_nodes);
LPGNN_API int lpgnn_basis_select(const float* logits_cons, int32_t m, const float* logits_vars, int32_t n,
                       int32_t k_basic, void* status, int status_is_i64, int32_t* counts_out,
                       void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* Tuning knob (process-wide): lpgnn_basis_select runs as ONE cooperative launch (keys in registers, 3-digit radix select
 * with grid-wide barriers) wherever the nodes fit the co-resident grid (default), or always as the 7-launch chain
 * (enable = 0).  Same statuses either way.  Returns the previous setting. */
LPGNN_API int lpgnn_set_select_fused(int enable);

/* Segmented variant for a block-diagonal PACK of LPs: segment b owns constraints
 * [cons_ptr[b], cons_ptr[b+1]) of logits_cons and variables [vars_ptr[b], vars_ptr[b+1]) of logits_vars and
 * gets exactly (cons_ptr[b+1]-cons_ptr[b]) basic nodes (val.inference_gnn applied per LP).  One CTA per
 * segment, one launch for the whole pack.  status: [total_cons + total_vars], all constraints first, in the
 * packed order.  Workspace: lpgnn_basis_select_workspace_bytes(total_cons + total_vars). */
LPGNN_API int lpgnn_basis_select_segmented(const float* logits_cons, const float* logits_vars,
                                 const int32_t* cons_ptr, const int32_t* vars_ptr, int32_t n_segments,
                                 int32_t total_cons, int32_t total_vars, void* status, int status_is_i64,
                                 void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* lp_major != 0: status position of local node i of segment b = cons_ptr[b] + vars_ptr[b] + i (each LP's statuses
 * contiguous, constraints first) instead of the packed layout. */
LPGNN_API int lpgnn_basis_select_segmented_ex(const float* logits_cons, const float* logits_vars,
                                    const int32_t* cons_ptr, const int32_t* vars_ptr, int32_t n_segments,
                                    int32_t total_cons, int32_t total_vars, void* status, int status_is_i64,
                                    int lp_major, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * One-call basis prediction: (a1) graph build -> (a2-a5) GCN_FC forward -> (a6) basis selection,
 * enqueued from native code on `stream`.  Replaces the per-LP body of the reference's prediction sweep
 * (scripts/pred_basis.py:113-118 `inference_only` = model(batch) + val.inference_gnn) together with the graph
 * construction of dataset.py:299-304.  Weights are passed by pointer (the caller owns them; bf16 mode
 * expects bf16 copies of the hidden-layer weights and the [W_rel|W_root|0] bf16 [hids,64] matrices of
 * the input layer).  status_out: uint8 [m+n] device, constraints first.  logits_out: optional f32
 * [m+n,3].  graph_status: optional zero-initialised device int32 (see lpgnn_graph_build); if NULL an
 * internal word is used.  All scratch memory comes from `workspace` (256-byte aligned).
 * ------------------------------------------------------------------------------------------- */
#define LPGNN_MAX_HIDDEN_LAYERS 8
typedef struct lpgnn_gcn_fc_weights {
  int32_t p, q, hids, depth, precision, reserved;  /* precision: LPGNN_F32 | LPGNN_BF16 | LPGNN_F16 (inference) */
  /* conv1, fp32: lin_rel.weight [hids,in_src], lin_rel.bias [hids], lin_root.weight [hids,in_dst] */
  const float *c1_l2r_wrel, *c1_l2r_b, *c1_l2r_wroot;   /* left2right: src = constraints (p), dst = variables (q) */
  const float *c1_r2l_wrel, *c1_r2l_b, *c1_r2l_wroot;   /* right2left: src = variables (q), dst = constraints (p) */
  const void *c1_l2r_wcat, *c1_r2l_wcat;                /* bf16 mode only */
  /* hidden layers i = 0 .. depth-3: weights [hids,hids] in the compute dtype, biases fp32 */
  const void* l2r_wrel[LPGNN_MAX_HIDDEN_LAYERS];
  const void* l2r_wroot[LPGNN_MAX_HIDDEN_LAYERS];
  const float* l2r_b[LPGNN_MAX_HIDDEN_LAYERS];
  const void* r2l_wrel[LPGNN_MAX_HIDDEN_LAYERS];
  const void* r2l_wroot[LPGNN_MAX_HIDDEN_LAYERS];
  const float* r2l_b[LPGNN_MAX_HIDDEN_LAYERS];
  /* heads, fp32: lin_left (constraints) / lin_right (variables): weight [3,hids], bias [3] */
  const float *head_left_w, *head_left_b, *head_right_w, *head_right_b;
  /* fp32 mode on the tensor cores (optional): x2 form of the hidden-layer weights, indexed [layer]: IEEE-half hi / lo
   * parts of W_rel and W_root [hids,hids] and ONE power-of-two scale per output feature shared by the two matrices
   * (lpgnn_split_x2(W_rel, hids, W_root, hids, rows = hids, ...)).  When l2r_wrel_hi[0] != NULL and hids % 64 == 0 the
   * hidden transforms of an LPGNN_F32 prediction run as lpgnn_node_transform_x2 (fp32-level accuracy from three
   * half x half tensor-core passes, chunked accumulation); otherwise the CUDA-core fp32 kernel is used. */
  const void* l2r_wrel_hi[LPGNN_MAX_HIDDEN_LAYERS];
  const void* l2r_wrel_lo[LPGNN_MAX_HIDDEN_LAYERS];
  const void* l2r_wroot_hi[LPGNN_MAX_HIDDEN_LAYERS];
  const void* l2r_wroot_lo[LPGNN_MAX_HIDDEN_LAYERS];
  const float* l2r_wscale[LPGNN_MAX_HIDDEN_LAYERS];
  const void* r2l_wrel_hi[LPGNN_MAX_HIDDEN_LAYERS];
  const void* r2l_wrel_lo[LPGNN_MAX_HIDDEN_LAYERS];
  const void* r2l_wroot_hi[LPGNN_MAX_HIDDEN_LAYERS];
  const void* r2l_wroot_lo[LPGNN_MAX_HIDDEN_LAYERS];
  const float* r2l_wscale[LPGNN_MAX_HIDDEN_LAYERS];
} lpgnn_gcn_fc_weights;

/* precision: LPGNN_F32, LPGNN_BF16, LPGNN_F16, or LPGNN_F32 | LPGNN_WS_X2 to reserve the x2 operand buffers of the
 * fp32 tensor-core mode. */
#define LPGNN_WS_X2 16
LPGNN_API size_t lpgnn_predict_workspace_bytes(int64_t nnz, int32_t m, int32_t n, int32_t p, int32_t q,
                                     int32_t hids, int32_t depth, int precision);
LPGNN_API int lpgnn_predict_basis(const lpgnn_gcn_fc_weights* w, const int32_t* coo_row, const int32_t* coo_col,
                        const float* coo_val, int64_t nnz, int32_t m, int32_t n, int flags,
                        const float* x_s, const float* x_t, uint8_t* status_out, float* logits_out,
                        int32_t* graph_status, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);
/* Same for a block-diagonal pack of n_segments LPs (the matrix is the direct sum of the LPs' matrices: rows
 * and columns already offset; m, n, nnz are the pack totals): the forward pass runs once over the pack and
 * the basis decision is taken per segment (lpgnn_basis_select_segmented).  cons_ptr / vars_ptr: device int32
 * [n_segments+1]. */
LPGNN_API int lpgnn_predict_basis_packed(const lpgnn_gcn_fc_weights* w, const int32_t* coo_row, const int32_t* coo_col,
                               const float* coo_val, int64_t nnz, int32_t m, int32_t n, int flags,
                               const float* x_s, const float* x_t, const int32_t* cons_ptr,
                               const int32_t* vars_ptr, int32_t n_segments, uint8_t* status_out,
                               float* logits_out, int32_t* graph_status, void* workspace,
                               size_t workspace_bytes, lpgnn_stream_t stream);

/* Tuning knob (process-wide): the two directions of a layer are independent (reference arch.py:183-184), so
 * lpgnn_predict_basis[_packed] can enqueue the constraint side's kernels on a library-owned side stream, forked from
 * and joined back into `stream` with events.  mode 0 = never, 1 = only LPs of up to 65536 nodes (default: larger LPs
 * fill the GPU with either side alone and measured no gain), 2 = always; LPGNN_PREDICT_FORK sets the initial mode.
 * Same results for every mode; everything the call enqueues is ordered before later work on `stream`.
 * Returns the previous mode. */
LPGNN_API int lpgnn_set_predict_fork(int mode);

/* =============================================================================================
 * Backward pass (training step: reference train.py:121-129 calls loss.backward(), which runs the
 * autograd formulas of PyG GraphConv / torch_sparse spmm_sum / F.normalize / relu_ / dropout).
 * The data-gradient GEMMs are lpgnn_node_transform with transposed weights, the aggregation
 * backward is lpgnn_spmm on the other orientation; the entry points below are the rest.
 * ============================================================================================= */

/* out[M,N] (f32) = A[M,K] * B[N,K]^T, bf16 operands, K a multiple of 64 and LARGE (the node dimension): the
 * weight-gradient GEMM dW = dPre^T X on transposed, zero-padded copies (lpgnn_transpose).  Few output
 * tiles and a long reduction, so the tcgen05 kernel runs split-K: (tile, K-slice) work items fill the SMs,
 * partial tiles go to `workspace` and are summed in a fixed order (deterministic). */
LPGNN_API int32_t lpgnn_gemm_tn_splits(int32_t M, int32_t N, int32_t K);
LPGNN_API size_t lpgnn_gemm_tn_workspace_bytes(int32_t M, int32_t N, int32_t K);
LPGNN_API int lpgnn_gemm_tn(const void* A, const void* B, int32_t M, int32_t N, int32_t K, float* out,
                  void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* dW[N_out, K_in] (f32) = dY[Mn, N_out]^T * X[Mn, K_in] (bf16, row-major, Mn = number of nodes): the weight
 * gradient of a node transform straight from the row-major activations -- the tcgen05 operands are
 * MN-major (TMA boxes of 64 nodes x 64 features), so no transposed copies exist; split-K over the nodes with
 * a fixed-order reduction (deterministic).  N_out, K_in multiples of 8 (K_in of 64). */
LPGNN_API size_t lpgnn_wgrad_workspace_bytes(int64_t Mn, int32_t N_out, int32_t K_in);
LPGNN_API int lpgnn_wgrad(const void* dY, const void* X, int64_t Mn, int32_t N_out, int32_t K_in, float* out,
                void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* Backward of lpgnn_head_mask wrt the hidden activation (reference arch.py:186-191, 129-141), fused with
 * the ReLU / inverted-dropout mask of that activation:
 *   draw[i,:] = d(10*raw/max(|raw|,1e-12))^T dlogits[i,:]        (mask offsets are constants)
 *   dH[i,:]   = (draw[i,:] * W) * scale * (Hact[i,:] > 0)
 * dlogits, raw, draw [rows,3] f32; Hact, dH [rows,Hdim] of h_dtype; W [3,Hdim] f32.  draw (optional
 * output) feeds the head's weight / bias gradients (lpgnn_small_wgrad, lpgnn_colsum).  draw_bf16 (optional
 * output, bf16 [rows,64] = [draw | 0]) is the same as an operand of the tensor-core weight gradient:
 * lpgnn_wgrad(Hact, draw_bf16) = [dW_head^T | 0]. */
LPGNN_API int lpgnn_head_mask_bwd(const float* dlogits, const float* raw, const void* Hact, int h_dtype,
                        int32_t rows, int32_t Hdim, const float* W, float scale, void* dH, float* draw,
                        void* draw_bf16, lpgnn_stream_t stream);
/* Same, and colsum_out[Hdim] (f32, optional) = column sums of dH accumulated in fp32 before the output rounding: the bias
 * gradient of the layer under the head (db = colsum(dPre), PyG GraphConv.lin_rel.bias) without a pass that reads dH back;
 * draw_colsum_out[3] (f32, optional) = column sums of draw: the head's own bias gradient (nn.Linear(H,3).bias).
 * Per-block partial sums in `workspace`, combined in a fixed order: deterministic. */
LPGNN_API size_t lpgnn_head_mask_bwd_colsum_workspace_bytes(int32_t rows, int32_t Hdim);
LPGNN_API int lpgnn_head_mask_bwd_colsum(const float* dlogits, const float* raw, const void* Hact, int h_dtype,
                               int32_t rows, int32_t Hdim, const float* W, float scale, void* dH, float* draw,
                               void* draw_bf16, float* colsum_out, float* draw_colsum_out, void* workspace,
                               size_t workspace_bytes, lpgnn_stream_t stream);

/* out = (a [+ b]) * scale * (act > 0), elementwise; b may be NULL; out may alias a.  Backward of
 * relu_ (reference arch.py:182,188) and of dropout followed by relu_ (arch.py:186-188: `act` is the
 * post-dropout activation, scale = 1/(1-p)); the optional b fuses the sum of the two gradient paths
 * of a node feature (root path + aggregated path).  count multiple of 16 bytes / sizeof(elem). */
LPGNN_API int lpgnn_relu_bwd(const void* a, const void* b, const void* act, int64_t count, int dtype,
                   float scale, void* out, lpgnn_stream_t stream);

/* In-place inverted dropout, x = keep ? x/(1-p) : 0 with keep drawn from a counter-based hash of
 * (seed, element index) (reference arch.py:186-187 F.dropout; the RNG stream necessarily differs). */
LPGNN_API int lpgnn_dropout(void* x, int64_t count, int dtype, float p, uint64_t seed, lpgnn_stream_t stream);

/* out[N, ld_out] = X[M,N]^T with the columns [M, ld_out) zero-filled (ld_out >= M; the weight-gradient
 * GEMM wants a reduction length that is a multiple of its K tile). */
LPGNN_API int lpgnn_transpose(const void* X, int dtype, int64_t M, int64_t N, void* out, int64_t ld_out,
                    lpgnn_stream_t stream);

/* out[c] = sum_r X[r,c] (bias gradients), two-stage with a fixed summation order. */
LPGNN_API size_t lpgnn_colsum_workspace_bytes(int64_t M, int32_t N);
LPGNN_API int lpgnn_colsum(const void* X, int dtype, int64_t M, int32_t N, float* out,
                 void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* dW[N,K] = dY[M,N]^T * Z[M,K] for narrow Z (K <= 64; Z f32 with row stride ldz) and optionally
 * dB[N] = column sums of dY: weight gradients of the input layer (Z = z_cat of lpgnn_conv_in_fused)
 * and of the head (dY = hidden activation, Z = draw).  Two-stage, fixed order (deterministic). */
LPGNN_API size_t lpgnn_small_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K);
LPGNN_API int lpgnn_small_wgrad(const void* dY, int dtype, const float* Z, int32_t ldz, int32_t K,
                      int64_t M, int32_t N, float* dW, float* dB,
                      void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (a1-a5, training) One training step of GCN_FC enqueued from native code.  Replaces `model(batch)` and the
 * activation-sized part of `loss.backward()` in the reference's training loop (train.py:121-128, the model of
 * arch.py:179-193 with dropout arch.py:186-187); the loss (train.py:32-53) and the optimiser (train.py:85-89)
 * stay with the caller, who passes d(loss)/d(logits) and receives fp32 parameter gradients.  Same kernels, order
 * and arithmetic as the op-by-op orchestration; ~75 launches come from two C calls.  The workspace holds every
 * activation the backward pass needs and must stay untouched between the two calls.  Weights as for
 * lpgnn_predict_basis (bf16 mode: bf16 copies of the hidden weights + the [hids,64] input-layer matrices).
 * Graph views as written by lpgnn_graph_build (CSR of A [m x n] and CSC).  dropout_p = 0 for eval mode.
 * ------------------------------------------------------------------------------------------- */
typedef struct lpgnn_gcn_fc_grads {   /* all fp32, shapes of the corresponding parameters */
  float *c1_l2r_wrel, *c1_l2r_b, *c1_l2r_wroot, *c1_r2l_wrel, *c1_r2l_b, *c1_r2l_wroot;
  float* l2r_wrel[LPGNN_MAX_HIDDEN_LAYERS];
  float* l2r_wroot[LPGNN_MAX_HIDDEN_LAYERS];
  float* l2r_b[LPGNN_MAX_HIDDEN_LAYERS];
  float* r2l_wrel[LPGNN_MAX_HIDDEN_LAYERS];
  float* r2l_wroot[LPGNN_MAX_HIDDEN_LAYERS];
  float* r2l_b[LPGNN_MAX_HIDDEN_LAYERS];
  float *head_left_w, *head_left_b, *head_right_w, *head_right_b;
} lpgnn_gcn_fc_grads;
LPGNN_API size_t lpgnn_train_workspace_bytes(int32_t m, int32_t n, int32_t p, int32_t q, int32_t hids,
                                             int32_t depth, int precision);
LPGNN_API int lpgnn_train_forward(const lpgnn_gcn_fc_weights* w, const int32_t* rowptr, const int32_t* col,
                        const float* val, const int32_t* colptr, const int32_t* row_csc, const float* val_csc,
                        int32_t m, int32_t n, const float* x_s, const float* x_t, float dropout_p, uint64_t seed,
                        float* logits_s, float* logits_t, void* workspace, size_t workspace_bytes,
                        lpgnn_stream_t stream);
LPGNN_API int lpgnn_train_backward(const lpgnn_gcn_fc_weights* w, const int32_t* rowptr, const int32_t* col,
                         const float* val, const int32_t* colptr, const int32_t* row_csc, const float* val_csc,
                         int32_t m, int32_t n, float dropout_p, const float* dlogits_s, const float* dlogits_t,
                         const lpgnn_gcn_fc_grads* grads, void* workspace, size_t workspace_bytes,
                         lpgnn_stream_t stream);
/* The same pass in two parts for data-parallel training: LPGNN_BWD_TAIL = head gradients + weight / bias gradients of
 * the LAST hidden layer (at depth 3: all [hids,hids] matrices, i.e. 99 % of the parameters), LPGNN_BWD_REST = the rest
 * (data gradients, earlier layers, input layer).  TAIL then REST enqueues exactly the kernels of lpgnn_train_backward in
 * the same order (bit-identical gradients); the caller starts the all-reduce of the tail gradients in between, so the
 * collective overlaps the remaining ~40 % of the backward pass (train.py, `training.set_gradient_sync`). */
#define LPGNN_BWD_TAIL 1
#define LPGNN_BWD_REST 2
LPGNN_API int lpgnn_train_backward_ex(const lpgnn_gcn_fc_weights* w, const int32_t* rowptr, const int32_t* col,
                         const float* val, const int32_t* colptr, const int32_t* row_csc, const float* val_csc,
                         int32_t m, int32_t n, float dropout_p, const float* dlogits_s, const float* dlogits_t,
                         const lpgnn_gcn_fc_grads* grads, int phases, void* workspace, size_t workspace_bytes,
                         lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (f-3) Balanced cross-entropy of the training loop, value and gradient in one pass.  Replaces balanced()
 * (reference train.py:39-46) = (m+n)/m * CE_w(logits_cons, y_s) + (m+n)/n * CE_w(logits_vars, y_t) with the
 * inverse-frequency class weights of labels_to_balanced_weights (utils.py:286-299; merge_lu averages the
 * lower / upper weights unless exactly two classes occur) and CrossEntropyLoss(weight=w) (weighted mean).
 * logits [rows,3] f32, labels int64 in {0,1,2} (other values are ignored, weight 0); loss_out: 1 float
 * (device); dlogits_* (optional, both or neither, f32 [rows,3]) = d loss / d logits.  Deterministic.
 * ------------------------------------------------------------------------------------------- */
LPGNN_API size_t lpgnn_balanced_ce_workspace_bytes(int32_t m, int32_t n);
LPGNN_API int lpgnn_balanced_ce(const float* logits_s, const int64_t* y_s, int32_t m, const float* logits_t,
                      const int64_t* y_t, int32_t n, int merge_lu, float* loss_out, float* dlogits_s,
                      float* dlogits_t, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* balanced() over a block-diagonal PACK of LPs (mini-batches of LP graphs; the reference trains one LP per step,
 * train.py:70): LP b owns constraints [cons_ptr[b], cons_ptr[b+1]) of logits_s / y_s and variables [vars_ptr[b],
 * vars_ptr[b+1]) of logits_t / y_t (device int32 [n_segments+1]), keeps its own class weights and (m_b+n_b)/m_b,
 * (m_b+n_b)/n_b factors (train.py:39-46 per graph); loss_out = mean over the LPs, dlogits_* = its gradient (the average
 * of the per-LP gradients).  One CTA per (LP, side); deterministic. */
LPGNN_API size_t lpgnn_balanced_ce_segmented_workspace_bytes(int32_t n_segments);
LPGNN_API int lpgnn_balanced_ce_segmented(const float* logits_s, const int64_t* y_s, const int32_t* cons_ptr,
                                const float* logits_t, const int64_t* y_t, const int32_t* vars_ptr, int32_t n_segments,
                                int merge_lu, float* loss_out, float* dlogits_s, float* dlogits_t, void* workspace,
                                size_t workspace_bytes, lpgnn_stream_t stream);

/* (f-3) The other two losses of the training loop, value and gradient in one kernel.  focal = 0: unbalanced()
 * (reference train.py:30-37) = F.cross_entropy over the concatenation of both sides (plain mean over m+n rows);
 * focal = 1: focal() (train.py:18-28, 49-53) = (1 - exp(-ce))^gamma * ce of that same mean CE (the reference applies
 * the focal factor to the batch mean; gamma = 2).  loss_out, grad_scale_out: one float each (device); dlogits_*
 * (optional, f32 [rows,3]) receive (softmax - onehot) / (m+n), and d loss / d logits = grad_scale_out * dlogits
 * (the caller folds the scalar into its upstream gradient, no second pass over the rows).  Deterministic. */
LPGNN_API size_t lpgnn_flat_ce_workspace_bytes(int32_t m, int32_t n);
LPGNN_API int lpgnn_flat_ce(const float* logits_s, const int64_t* y_s, int32_t m, const float* logits_t,
                  const int64_t* y_t, int32_t n, int focal, float gamma, float* loss_out, float* grad_scale_out,
                  float* dlogits_s, float* dlogits_t, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);

/* (f-3) Counters behind val.accuracy (reference val.py:199-237: per-side accuracy and sklearn precision / recall of
 * class 1): counts_out (device int32[8]) = per side {#pred == gt, #pred == 1 and gt == 1, #pred == 1, #gt == 1},
 * constraints then variables.  status [m+n] as written by lpgnn_basis_select (uint8 or int64), labels int64.
 * One kernel; the caller reads eight integers instead of moving both vectors to the host. */
LPGNN_API int lpgnn_basis_metrics(const void* status, int status_is_i64, const int64_t* y_s, int32_t m,
                        const int64_t* y_t, int32_t n, int32_t* counts_out, lpgnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (f-2) LP scaling + node features on the device.  Replaces dataset.scaling (reference dataset.py:23-76,
 * utils.py:323-332) and dataset.cvt_to_features (dataset.py:79-96, utils.py:335-383), float64 like the
 * reference: a raw LP  min c^T x, b_l <= A x <= b_u, l <= x <= u  becomes the scaled matrix values of both graph
 * orientations (fp32, written over val / val_csc of a structure built by lpgnn_graph_build) and the feature
 * rows x_s [m,8], x_t [n,8] (fp32, layout of SURVEY Appendix A; columns 5 and 7 are the +-inf tags the mask
 * kernel tests).  a_csr: raw float64 values in canonical CSR order.  The scaled LP is also returned in float64
 * (a_scaled [nnz] CSR order, c_out [n], bl_out/bu_out [m], l_out/u_out [n]; all required).  Per-row / per-column
 * sums run in CSR / CSC order with separately rounded multiplies and adds: scaled values and dot products are
 * bit-identical to the reference's scipy arithmetic; the five vector norms differ from numpy's pairwise sums in
 * the last bits only.
 * ------------------------------------------------------------------------------------------- */
LPGNN_API size_t lpgnn_lp_features_workspace_bytes(int64_t nnz, int32_t m, int32_t n);
LPGNN_API int lpgnn_lp_features(const int32_t* rowptr, const int32_t* col, const int32_t* colptr, const int32_t* row_csc,
                      const int32_t* csr2csc, const double* a_csr, const double* c, const double* b_l,
                      const double* b_u, const double* l, const double* u, int64_t nnz, int32_t m, int32_t n,
                      float* val, float* val_csc, float* x_s, float* x_t, double* a_scaled, double* c_out,
                      double* bl_out, double* bu_out, double* l_out, double* u_out, void* workspace,
                      size_t workspace_bytes, lpgnn_stream_t stream);

/* Tuning knob (process-wide): the wide bf16 node transform runs as 2-CTA clusters that share the W tiles through TMA
 * multicast (default on); 0 selects the one-CTA-per-tile form.  Returns the previous setting; results are identical. */
LPGNN_API int lpgnn_set_gemm_cluster(int enable);

/* ---------------------------------------------------------------------------------------------
 * (f-4) Sampled-subgraph path for LPs above edge_num_thresh.  Replaces torch_geometric NeighborLoader as driven at
 * reference train.py:107-116 (num_neighbors=[6]*depth, directed=False) and val.py:22-27 (num_neighbors=[-1]*depth)
 * plus the relabelling of MyToBipartite (dataset.py:275-332), on the resident CSR / CSC of the full LP.
 *
 * lpgnn_sample_mark: for every node f of `frontier` (ids in the row space of (ptr, idx)) choose min(deg, fanout) of
 * its neighbours without replacement (fanout < 0: all) and set marks_other[neighbour] = 1 (uint8, other side's id
 * space).  The choice is a pure function of (seed, f): reproducible; callers fold the hop number into `seed`.
 * lpgnn_induced_count / _fill: subgraph induced by the sampled nodes.  rows[i] = global id of local row i;
 * map_other[j] = local id of other-side node j or -1.  counts[i] = surviving entries of row i; with
 * offsets = exclusive scan of counts, _fill writes the COO (local row, local col, value) in local row order
 * (columns in the order of the full matrix's row, so only sorted if map_other is monotone).
 * ------------------------------------------------------------------------------------------- */
LPGNN_API int lpgnn_sample_mark(const int32_t* ptr, const int32_t* idx, const int32_t* frontier, int32_t n_frontier,
                      int32_t fanout, uint64_t seed, uint8_t* marks_other, lpgnn_stream_t stream);
LPGNN_API int lpgnn_induced_count(const int32_t* ptr, const int32_t* idx, const int32_t* rows, int32_t n_rows,
                        const int32_t* map_other, int32_t* counts, lpgnn_stream_t stream);
LPGNN_API int lpgnn_induced_fill(const int32_t* ptr, const int32_t* idx, const float* val, const int32_t* rows,
                       int32_t n_rows, const int32_t* map_other, const int64_t* offsets, int32_t* out_row,
                       int32_t* out_col, float* out_val, lpgnn_stream_t stream);

/* (f-4) One mini-batch with ONE host read.  lpgnn_sample_nodes builds the node sets of a sampled mini-batch entirely on
 * the device: seeds (global unipartite ids, int64: constraints 0..m-1, variables m..m+n-1; dataset.py:258-260) are split by
 * side in seed order, then `n_hops` hops of neighbour sampling (fanouts_host[h] neighbours per frontier node, < 0 = all;
 * NeighborLoader num_neighbors, train.py:110 / val.py:23) append the newly reached nodes of each side in ascending id
 * order -- the local numbering of MyToBipartite (dataset.py:288-294).  cons_nodes [m] / var_nodes [n]: local -> global id
 * (filled up to the counts below); map_cons [m] / map_vars [n]: global -> local id or -1.  sizes: device int32
 * [lpgnn_sample_sizes_len()]: [0] = sampled constraints, [1] = sampled variables, [2] / [3] = seed constraints / variables
 * (s_bs, t_bs), [4 + 2h + side] = members after hop h.  lpgnn_induced_offsets then writes, for the sampled constraints,
 * the start of every local row in the induced subgraph (offsets [rows_capacity + 1]) and its nnz into
 * sizes[lpgnn_sample_sizes_len() - 2]; the caller reads `sizes` ONCE, allocates the COO and calls
 * lpgnn_induced_fill_sorted, which emits the entries in canonical order (local row, ascending local column), i.e. ready
 * for lpgnn_graph_build(LPGNN_COO_SORTED).  All ordered compactions are stable scans (decoupled look-back): deterministic. */
LPGNN_API int32_t lpgnn_sample_sizes_len(void);
LPGNN_API size_t lpgnn_sample_nodes_workspace_bytes(int32_t m, int32_t n, int32_t n_seeds);
LPGNN_API int lpgnn_sample_nodes(const int32_t* rowptr, const int32_t* col, const int32_t* colptr, const int32_t* row_csc,
                       int32_t m, int32_t n, const int64_t* seeds, int32_t n_seeds, const int32_t* fanouts_host,
                       int32_t n_hops, uint64_t seed, int32_t* cons_nodes, int32_t* var_nodes, int32_t* map_cons,
                       int32_t* map_vars, int32_t* sizes, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream);
LPGNN_API int lpgnn_induced_offsets(const int32_t* ptr, const int32_t* idx, const int32_t* rows, int32_t rows_capacity,
                          const int32_t* map_other, int32_t* offsets, int32_t* sizes, void* workspace,
                          size_t workspace_bytes, lpgnn_stream_t stream);
LPGNN_API int lpgnn_induced_fill_sorted(const int32_t* ptr, const int32_t* idx, const float* val, const int32_t* rows,
                              int32_t n_rows, const int32_t* map_other, const int32_t* offsets, int32_t* out_row,
                              int32_t* out_col, float* out_val, lpgnn_stream_t stream);
/* Node data of the sampled mini-batch in ONE launch (what NeighborLoader's n_id slicing yields, train.py:117-123):
 * out_xs [mc,p] = x_s[cons_nodes], out_xt [nv,q] = x_t[var_nodes], out_ys / out_yt = int64 labels of those nodes
 * (y_s, y_t, out_ys, out_yt may all be null), ids_s / ids_t = copies of the node lists (cons_nodes / var_nodes live in
 * the sampler's reused buffers). */
LPGNN_API int lpgnn_sample_gather(const float* x_s, const float* x_t, const int64_t* y_s, const int64_t* y_t,
                        const int32_t* cons_nodes, int32_t mc, const int32_t* var_nodes, int32_t nv, int32_t p,
                        int32_t q, float* out_xs, float* out_xt, int64_t* out_ys, int64_t* out_yt, int32_t* ids_s,
                        int32_t* ids_t, lpgnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LPGNN_H_ */
