// Training step of GCN_FC enqueued from native code: lpgnn_train_forward / lpgnn_train_backward.
//
// Replaces model(batch) + loss.backward() of the reference's training loop (train.py:121-128) for the
// activation-sized work: the loss itself (train.py:32-53) stays with the caller, who passes d(loss)/d(logits).
// Same kernels and the same order as the Python orchestration in training.py; what changes is the host cost:
// ~75 kernel launches are issued by two C calls instead of ~75 Python round trips, which bounds small LPs.
// All activations that the backward pass needs live in ONE caller-owned workspace between the two calls.
//
//   forward   conv1 (+relu) -> [spmm x2 -> transform x2 (+relu) -> dropout x2]* -> head+mask (raw logits kept)
//   backward  head_mask_bwd -> head wgrad -> per hidden layer (last to first): weight grads (MN-major tensor-core
//             GEMM in bf16 mode; transposes + CUDA-core GEMM in fp32 mode), bias grads, data grads (transform with
//             transposed weights), aggregation backward (spmm on the other orientation), relu/dropout mask with the
//             sum of both gradient paths -> conv1 weight grads.
#include "common.cuh"

using namespace lpgnn;

namespace {

struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* p) : base(reinterpret_cast<char*>(p)) {}
  template <typename T> T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
};

struct TB {
  float *z_s, *z_t;
  void *zb_s, *zb_t;
  void *left[LPGNN_MAX_HIDDEN_LAYERS + 1], *right[LPGNN_MAX_HIDDEN_LAYERS + 1];
  void *agg_s[LPGNN_MAX_HIDDEN_LAYERS], *agg_t[LPGNN_MAX_HIDDEN_LAYERS];
  float *raw_s, *raw_t;
  // backward
  void *dpre_s, *dpre_t, *droot_s, *droot_t, *dlagg, *dragg;
  float *draw_s, *draw_t;
  void *drawb_s, *drawb_t;   // bf16 [rows,64] = [draw | 0]
  void *wS[LPGNN_MAX_HIDDEN_LAYERS][4], *wT[LPGNN_MAX_HIDDEN_LAYERS][4];  // compute-dtype weights, straight / transposed
  void* wcat[2];       // bf16 [H,64] input-layer matrices [W_rel | W_root | 0]
  float *g_tmp;        // [H, max(p+q, 2H)] fp32 scratch for concatenated / untransposed weight grads
  void *tr_d, *tr_x;   // fp32 mode: transposed operands of the weight-gradient GEMM
  void* scratch; size_t scratch_bytes;
};

size_t carve(Bump& b, TB& B, int32_t m, int32_t n, int32_t p, int32_t q, int32_t H, int32_t depth, int bf16) {
  const size_t es = bf16 ? 2 : 4;
  const int nh = depth - 2;
  const int kt = lpgnn_conv_in_zcat_width(p, q);
  B.z_s = bf16 ? nullptr : b.take<float>((size_t)m * kt);
  B.z_t = bf16 ? nullptr : b.take<float>((size_t)n * kt);
  B.zb_s = bf16 ? b.take<char>((size_t)m * 128) : nullptr;
  B.zb_t = bf16 ? b.take<char>((size_t)n * 128) : nullptr;
  for (int i = 0; i <= nh; ++i) { B.left[i] = b.take<char>((size_t)m * H * es); B.right[i] = b.take<char>((size_t)n * H * es); }
  for (int i = 0; i < nh; ++i) { B.agg_s[i] = b.take<char>((size_t)m * H * es); B.agg_t[i] = b.take<char>((size_t)n * H * es); }
  B.raw_s = b.take<float>((size_t)m * 3); B.raw_t = b.take<float>((size_t)n * 3);
  B.dpre_s = b.take<char>((size_t)m * H * es); B.dpre_t = b.take<char>((size_t)n * H * es);
  B.droot_s = b.take<char>((size_t)m * H * es); B.droot_t = b.take<char>((size_t)n * H * es);
  B.dlagg = b.take<char>((size_t)m * H * es); B.dragg = b.take<char>((size_t)n * H * es);
  B.draw_s = b.take<float>((size_t)m * 3); B.draw_t = b.take<float>((size_t)n * 3);
  B.drawb_s = bf16 ? b.take<char>((size_t)m * 128) : nullptr;
  B.drawb_t = bf16 ? b.take<char>((size_t)n * 128) : nullptr;
  for (int i = 0; i < nh; ++i)
    for (int k = 0; k < 4; ++k) {
      B.wS[i][k] = bf16 ? b.take<char>((size_t)H * H * es) : nullptr;
      B.wT[i][k] = b.take<char>((size_t)H * H * es);
    }
  for (int k = 0; k < 2; ++k) B.wcat[k] = bf16 ? b.take<char>((size_t)H * 128) : nullptr;
  const size_t gcols = (size_t)(2 * H > 64 ? 2 * H : 64);
  B.g_tmp = b.take<float>((size_t)H * gcols);
  if (!bf16 && nh > 0) {
    const size_t ld = align_up((size_t)(n > m ? n : m), 64);
    B.tr_d = b.take<char>((size_t)H * ld * es);
    B.tr_x = b.take<char>((size_t)2 * H * ld * es);
  } else {
    B.tr_d = B.tr_x = nullptr;
  }
  const int64_t big = n > m ? n : m;
  size_t sc = lpgnn_small_wgrad_workspace_bytes(big, H, p + q > 3 ? p + q : 3);
  const size_t c1 = lpgnn_colsum_workspace_bytes(big, H), c3 = lpgnn_colsum_workspace_bytes(big, 3);
  if (c1 > sc) sc = c1;
  if (c3 > sc) sc = c3;
  const size_t c5 = lpgnn_head_mask_bwd_colsum_workspace_bytes((int32_t)big, H);
  if (c5 > sc) sc = c5;
  if (bf16) {
    const size_t c2 = nh > 0 ? lpgnn_wgrad_workspace_bytes(big, H, H) : 0, c4 = lpgnn_wgrad_workspace_bytes(big, H, 64);
    if (c2 > sc) sc = c2;
    if (c4 > sc) sc = c4;
  }
  B.scratch_bytes = sc;
  B.scratch = b.take<char>(sc);
  return align_up(b.off, 256);
}

// Per-step weight preparation, one launch for all hidden matrices: compute-dtype copy (bf16 mode) and the
// transposed copy the data-gradient transforms read ([K,N] K-major for the TN kernel).
struct PrepArgs {
  const float* src[4 * LPGNN_MAX_HIDDEN_LAYERS];
  void* dst[4 * LPGNN_MAX_HIDDEN_LAYERS];
  void* dstT[4 * LPGNN_MAX_HIDDEN_LAYERS];
};

template <typename T>
__global__ void __launch_bounds__(256) prep_weights_kernel(const __grid_constant__ PrepArgs a, int H) {
  __shared__ float tile[32][33];
  const float* src = a.src[blockIdx.z];
  T* dst = reinterpret_cast<T*>(a.dst[blockIdx.z]);
  T* dstT = reinterpret_cast<T*>(a.dstT[blockIdx.z]);
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < H && c < H) {
      v = src[(size_t)r * H + c];
      if (dst) dst[(size_t)r * H + c] = T(v);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = c0 + j, c = r0 + threadIdx.x;   // transposed coordinates
    if (r < H && c < H) dstT[(size_t)r * H + c] = T(tile[threadIdx.x][j]);
  }
}

// [W_rel | W_root | 0] as bf16 [H,64] for both directions of the input layer (blockIdx.y).
__global__ void wcat_kernel(const float* rel0, const float* root0, __nv_bfloat16* out0, int krel0, int kroot0,
                            const float* rel1, const float* root1, __nv_bfloat16* out1, int krel1, int kroot1, int H) {
  const float* rel = blockIdx.y ? rel1 : rel0;
  const float* root = blockIdx.y ? root1 : root0;
  __nv_bfloat16* out = blockIdx.y ? out1 : out0;
  const int krel = blockIdx.y ? krel1 : krel0, kroot = blockIdx.y ? kroot1 : kroot0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * 64) return;
  const int r = i >> 6, c = i & 63;
  float v = 0.f;
  if (c < krel) v = rel[r * krel + c];
  else if (c < krel + kroot) v = root[r * kroot + (c - krel)];
  out[i] = __float2bfloat16(v);
}

// dst[c, h] = src[h, c] for c < ncols: the first columns of a [H, ld] gradient as a [ncols, H] matrix.
__global__ void take_cols_t_kernel(const float* __restrict__ src, int ld, int H, int ncols, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * ncols) { const int c = i / H, h = i % H; dst[i] = src[(size_t)h * ld + c]; }
}

// [H,64] input-layer gradient -> lin_rel.weight [H,ks], lin_root.weight [H,kd], lin_rel.bias [H] (column ks+kd).
__global__ void split_c1_grad_kernel(const float* __restrict__ src, int H, int ks, int kd, float* __restrict__ wrel,
                                     float* __restrict__ wroot, float* __restrict__ bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * 64) return;
  const int h = i >> 6, c = i & 63;
  const float v = src[i];
  if (c < ks) wrel[h * ks + c] = v;
  else if (c < ks + kd) wroot[h * kd + (c - ks)] = v;
  else if (c == ks + kd && bias) bias[h] = v;
}

#define LPGNN_TRY(expr) do { if (int _rc = (expr)) return _rc; } while (0)

int copy2d(float* dst, int dst_cols, const float* src, int src_cols, int width, int rows, cudaStream_t st) {
  LPGNN_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)dst_cols * 4, src, (size_t)src_cols * 4, (size_t)width * 4, rows,
                                  cudaMemcpyDeviceToDevice, st));
  return LPGNN_OK;
}

}  // namespace

extern "C" size_t lpgnn_train_workspace_bytes(int32_t m, int32_t n, int32_t p, int32_t q, int32_t hids, int32_t depth,
                                              int precision) {
  Bump b(nullptr);
  TB B;
  return carve(b, B, m, n, p, q, hids, depth, precision == LPGNN_BF16);
}

extern "C" int lpgnn_train_forward(const lpgnn_gcn_fc_weights* w, const int32_t* rowptr, const int32_t* col,
                                   const float* val, const int32_t* colptr, const int32_t* row_csc,
                                   const float* val_csc, int32_t m, int32_t n, const float* x_s, const float* x_t,
                                   float dropout_p, uint64_t seed, float* logits_s, float* logits_t, void* workspace,
                                   size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(w && rowptr && colptr && x_s && x_t && logits_s && logits_t && workspace && m > 0 && n > 0,
                "train_forward: bad arguments");
  const int H = w->hids, depth = w->depth, p = w->p, q = w->q, nh = depth - 2;
  const int bf16 = w->precision == LPGNN_BF16;
  LPGNN_REQUIRE(depth >= 2 && nh <= LPGNN_MAX_HIDDEN_LAYERS && (!bf16 || H % 64 == 0), "train_forward: unsupported shape");
  LPGNN_REQUIRE((uintptr_t)workspace % 256 == 0, "train_forward: workspace must be 256-byte aligned");
  if (workspace_bytes < lpgnn_train_workspace_bytes(m, n, p, q, H, depth, w->precision)) {
    set_error("train_forward: workspace too small");
    return LPGNN_EWORKSPACE;
  }
  Bump b(workspace);
  TB B;
  carve(b, B, m, n, p, q, H, depth, bf16);
  const int dt = bf16 ? LPGNN_BF16 : LPGNN_F32;
  cudaStream_t st = (cudaStream_t)stream;
  // weights: `w` carries the fp32 master parameters; compute-dtype and transposed copies are made here
  if (nh > 0) {
    PrepArgs a;
    for (int i = 0; i < nh; ++i) {
      const void* src[4] = {w->l2r_wrel[i], w->l2r_wroot[i], w->r2l_wrel[i], w->r2l_wroot[i]};
      for (int k = 0; k < 4; ++k) {
        LPGNN_REQUIRE(src[k], "train_forward: missing hidden-layer weight");
        a.src[4 * i + k] = (const float*)src[k];
        a.dst[4 * i + k] = B.wS[i][k];
        a.dstT[4 * i + k] = B.wT[i][k];
      }
    }
    dim3 grid((H + 31) / 32, (H + 31) / 32, 4 * nh), block(32, 8);
    if (bf16) prep_weights_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(a, H);
    else prep_weights_kernel<float><<<grid, block, 0, st>>>(a, H);
    LPGNN_LAUNCH_OK();
    count_launches(1);
  }
  const void *wl2r_rel[LPGNN_MAX_HIDDEN_LAYERS], *wl2r_root[LPGNN_MAX_HIDDEN_LAYERS], *wr2l_rel[LPGNN_MAX_HIDDEN_LAYERS],
      *wr2l_root[LPGNN_MAX_HIDDEN_LAYERS];
  for (int i = 0; i < nh; ++i) {
    wl2r_rel[i] = bf16 ? B.wS[i][0] : w->l2r_wrel[i];
    wl2r_root[i] = bf16 ? B.wS[i][1] : w->l2r_wroot[i];
    wr2l_rel[i] = bf16 ? B.wS[i][2] : w->r2l_wrel[i];
    wr2l_root[i] = bf16 ? B.wS[i][3] : w->r2l_wroot[i];
  }
  // conv1 (+relu)
  if (bf16 && p == 8 && q == 8 && H % 32 == 0 && H <= 4096) {
    // the reference's shape: aggregate + transform + ReLU + store of both directions in one kernel; zb = [z | 1 | 0] bf16 is the
    // operand of the layer's weight gradient (lpgnn_wgrad below)
    LPGNN_TRY(lpgnn_conv_in_16_pair(rowptr, col, val, colptr, row_csc, val_csc, m, n, x_s, x_t, w->c1_l2r_wrel, w->c1_l2r_b,
                                    w->c1_l2r_wroot, w->c1_r2l_wrel, w->c1_r2l_b, w->c1_r2l_wroot, H, B.left[0], B.right[0], dt,
                                    LPGNN_EPI_RELU, B.zb_s, B.zb_t, stream));
  } else if (bf16) {
    LPGNN_REQUIRE(p + q <= 64, "train_forward: bf16 input layer needs p + q <= 64");
    wcat_kernel<<<dim3((H * 64 + 255) / 256, 2), 256, 0, st>>>(
        w->c1_l2r_wrel, w->c1_l2r_wroot, (__nv_bfloat16*)B.wcat[0], p, q,
        w->c1_r2l_wrel, w->c1_r2l_wroot, (__nv_bfloat16*)B.wcat[1], q, p, H);
    LPGNN_LAUNCH_OK();
    count_launches(1);
    LPGNN_TRY(lpgnn_gather_cat(colptr, row_csc, val_csc, n, x_s, p, x_t, q, B.z_t, B.zb_t, stream));
    LPGNN_TRY(lpgnn_gather_cat(rowptr, col, val, m, x_t, q, x_s, p, B.z_s, B.zb_s, stream));
    LPGNN_TRY(lpgnn_node_transform(B.zb_t, 64, B.wcat[0], nullptr, 0, nullptr, w->c1_l2r_b, n, H, B.right[0], dt, dt,
                                   LPGNN_EPI_RELU, stream));
    LPGNN_TRY(lpgnn_node_transform(B.zb_s, 64, B.wcat[1], nullptr, 0, nullptr, w->c1_r2l_b, m, H, B.left[0], dt, dt,
                                   LPGNN_EPI_RELU, stream));
  } else {
    LPGNN_TRY(lpgnn_conv_in_fused(colptr, row_csc, val_csc, n, x_s, p, x_t, q, w->c1_l2r_wrel, w->c1_l2r_b, w->c1_l2r_wroot,
                                  H, B.right[0], dt, LPGNN_EPI_RELU, B.z_t, stream));
    LPGNN_TRY(lpgnn_conv_in_fused(rowptr, col, val, m, x_t, q, x_s, p, w->c1_r2l_wrel, w->c1_r2l_b, w->c1_r2l_wroot, H,
                                  B.left[0], dt, LPGNN_EPI_RELU, B.z_s, stream));
  }
  for (int li = 0; li < nh; ++li) {
    // A^T . left and A . right in one launch where both take the banded sweep (nnz is not known on the host here: -1)
    LPGNN_TRY(lpgnn_spmm_pair(rowptr, col, val, m, colptr, row_csc, val_csc, n, -1, B.left[li], B.right[li], B.agg_s[li],
                              B.agg_t[li], H, dt, stream));
    // reference order is dropout then relu_ (arch.py:186-188); the two commute and both sit in the epilogue
    lpgnn_epilogue_args ea;
    ea.epilogue = LPGNN_EPI_RELU; ea.dropout_p = dropout_p; ea.mask_act = nullptr; ea.mask_scale = 1.f;
    if (bf16 && li == nh - 1) {
      // last hidden layer: the basis-status head rides in the transform's epilogue (on the values that are stored, after
      // dropout), so the stored activation is not read back for it; the partial dot products live in the gradient
      // buffers of the backward pass (dpre_*: unused until then, >= 12 * nparts bytes per row)
      const int nparts = lpgnn_node_transform_head_parts(H);
      float *part_t = reinterpret_cast<float*>(B.dpre_t), *part_s = reinterpret_cast<float*>(B.dpre_s);
      ea.dropout_seed = seed + 2 * li;
      LPGNN_TRY(lpgnn_node_transform_head_train(B.agg_t[li], H, wl2r_rel[li], B.right[li], H, wl2r_root[li], w->l2r_b[li], n, H,
                                                B.right[li + 1], &ea, w->head_right_w, part_t, stream));
      ea.dropout_seed = seed + 2 * li + 1;
      LPGNN_TRY(lpgnn_node_transform_head_train(B.agg_s[li], H, wr2l_rel[li], B.left[li], H, wr2l_root[li], w->r2l_b[li], m, H,
                                                B.left[li + 1], &ea, w->head_left_w, part_s, stream));
      LPGNN_TRY(lpgnn_head_finish_ex(part_s, nparts, m, w->head_left_b, x_s, p, logits_s, B.raw_s, stream));
      LPGNN_TRY(lpgnn_head_finish_ex(part_t, nparts, n, w->head_right_b, x_t, q, logits_t, B.raw_t, stream));
      return LPGNN_OK;
    }
    ea.dropout_seed = seed + 2 * li;
    LPGNN_TRY(lpgnn_node_transform_ex(B.agg_t[li], H, wl2r_rel[li], B.right[li], H, wl2r_root[li], w->l2r_b[li], n, H,
                                      B.right[li + 1], dt, &ea, stream));
    ea.dropout_seed = seed + 2 * li + 1;
    LPGNN_TRY(lpgnn_node_transform_ex(B.agg_s[li], H, wr2l_rel[li], B.left[li], H, wr2l_root[li], w->r2l_b[li], m, H,
                                      B.left[li + 1], dt, &ea, stream));
  }
  LPGNN_TRY(lpgnn_head_mask(B.left[nh], dt, m, H, w->head_left_w, w->head_left_b, x_s, p, logits_s, B.raw_s, stream));
  LPGNN_TRY(lpgnn_head_mask(B.right[nh], dt, n, H, w->head_right_w, w->head_right_b, x_t, q, logits_t, B.raw_t, stream));
  return LPGNN_OK;
}

extern "C" int lpgnn_train_backward(const lpgnn_gcn_fc_weights* w, const int32_t* rowptr, const int32_t* col,
                                    const float* val, const int32_t* colptr, const int32_t* row_csc,
                                    const float* val_csc, int32_t m, int32_t n, float dropout_p,
                                    const float* dlogits_s, const float* dlogits_t, const lpgnn_gcn_fc_grads* g,
                                    void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  return lpgnn_train_backward_ex(w, rowptr, col, val, colptr, row_csc, val_csc, m, n, dropout_p, dlogits_s, dlogits_t, g,
                                 LPGNN_BWD_TAIL | LPGNN_BWD_REST, workspace, workspace_bytes, stream);
}

// phases: LPGNN_BWD_TAIL = head gradients + the weight / bias gradients of the LAST hidden layer (the bulk of the
// parameters at depth 3), LPGNN_BWD_REST = everything else (data gradients, earlier layers, input layer).  Calling TAIL
// then REST enqueues exactly the kernels of the one-call form in the same order; a data-parallel caller starts the
// all-reduce of the tail gradients between the two calls so that it overlaps the rest of the backward pass.
extern "C" int lpgnn_train_backward_ex(const lpgnn_gcn_fc_weights* w, const int32_t* rowptr, const int32_t* col,
                                       const float* val, const int32_t* colptr, const int32_t* row_csc,
                                       const float* val_csc, int32_t m, int32_t n, float dropout_p,
                                       const float* dlogits_s, const float* dlogits_t, const lpgnn_gcn_fc_grads* g,
                                       int phases, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(phases != 0 && (phases & ~(LPGNN_BWD_TAIL | LPGNN_BWD_REST)) == 0, "train_backward: bad phases %d", phases);
  LPGNN_REQUIRE(w && g && rowptr && colptr && dlogits_s && dlogits_t && workspace && m > 0 && n > 0,
                "train_backward: bad arguments");
  const int H = w->hids, depth = w->depth, p = w->p, q = w->q, nh = depth - 2;
  const int bf16 = w->precision == LPGNN_BF16;
  if (workspace_bytes < lpgnn_train_workspace_bytes(m, n, p, q, H, depth, w->precision)) {
    set_error("train_backward: workspace too small");
    return LPGNN_EWORKSPACE;
  }
  Bump b(workspace);
  TB B;
  carve(b, B, m, n, p, q, H, depth, bf16);
  cudaStream_t st = (cudaStream_t)stream;
  const int dt = bf16 ? LPGNN_BF16 : LPGNN_F32;
  const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  // ---- heads: dPre of the last activations (relu / dropout mask fused), head weight + bias grads
  const float last_scale = nh > 0 ? scale : 1.f;
  auto hidden_wgrads = [&](int li, const void* dps, const void* dpt) -> int {
    if (bf16) {  // MN-major tensor-core operands: dW = dPre^T X straight from the row-major activations
      LPGNN_TRY(lpgnn_wgrad(dpt, B.agg_t[li], n, H, H, g->l2r_wrel[li], B.scratch, B.scratch_bytes, stream));
      LPGNN_TRY(lpgnn_wgrad(dpt, B.right[li], n, H, H, g->l2r_wroot[li], B.scratch, B.scratch_bytes, stream));
      LPGNN_TRY(lpgnn_wgrad(dps, B.agg_s[li], m, H, H, g->r2l_wrel[li], B.scratch, B.scratch_bytes, stream));
      LPGNN_TRY(lpgnn_wgrad(dps, B.left[li], m, H, H, g->r2l_wroot[li], B.scratch, B.scratch_bytes, stream));
    } else {     // fp32: transposed copies + CUDA-core GEMM over the concatenated [rel | root] operand
      for (int side = 0; side < 2; ++side) {
        const int64_t rows = side ? m : n;
        const int64_t ld = (rows + 63) / 64 * 64;
        const void* dpre = side ? dps : dpt;
        const void* xrel = side ? B.agg_s[li] : B.agg_t[li];
        const void* xroot = side ? B.left[li] : B.right[li];
        LPGNN_TRY(lpgnn_transpose(dpre, dt, rows, H, B.tr_d, ld, stream));
        LPGNN_TRY(lpgnn_transpose(xrel, dt, rows, H, B.tr_x, ld, stream));
        LPGNN_TRY(lpgnn_transpose(xroot, dt, rows, H, (char*)B.tr_x + (size_t)H * ld * 4, ld, stream));
        LPGNN_TRY(lpgnn_node_transform(B.tr_d, (int32_t)ld, B.tr_x, nullptr, 0, nullptr, nullptr, H, 2 * H, B.g_tmp,
                                       LPGNN_F32, LPGNN_F32, LPGNN_EPI_NONE, stream));
        LPGNN_TRY(copy2d(side ? g->r2l_wrel[li] : g->l2r_wrel[li], H, B.g_tmp, 2 * H, H, H, st));
        LPGNN_TRY(copy2d(side ? g->r2l_wroot[li] : g->l2r_wroot[li], H, B.g_tmp + H, 2 * H, H, H, st));
      }
    }
    if (li < nh - 1) {   // the layer under the head got its bias gradients from head_mask_bwd
      LPGNN_TRY(lpgnn_colsum(dpt, dt, n, H, g->l2r_b[li], B.scratch, B.scratch_bytes, stream));
      LPGNN_TRY(lpgnn_colsum(dps, dt, m, H, g->r2l_b[li], B.scratch, B.scratch_bytes, stream));
    }
    return LPGNN_OK;
  };
  if (phases & LPGNN_BWD_TAIL) {
  // with a hidden layer under the head, dPre's column sums (that layer's bias gradients) come out of the same pass
  // (and the column sums of draw: the heads' own bias gradients)
  LPGNN_TRY(lpgnn_head_mask_bwd_colsum(dlogits_s, B.raw_s, B.left[nh], dt, m, H, w->head_left_w, last_scale, B.dpre_s, B.draw_s,
                                       B.drawb_s, nh > 0 ? g->r2l_b[nh - 1] : nullptr, g->head_left_b, B.scratch,
                                       B.scratch_bytes, stream));
  LPGNN_TRY(lpgnn_head_mask_bwd_colsum(dlogits_t, B.raw_t, B.right[nh], dt, n, H, w->head_right_w, last_scale, B.dpre_t, B.draw_t,
                                       B.drawb_t, nh > 0 ? g->l2r_b[nh - 1] : nullptr, g->head_right_b, B.scratch,
                                       B.scratch_bytes, stream));
  for (int side = 0; side < 2; ++side) {   // head weight [3,H] and bias [3] gradients
    const int64_t rows = side ? n : m;
    const void* act = side ? B.right[nh] : B.left[nh];
    float* gw = side ? g->head_right_w : g->head_left_w;
    const int ld = bf16 ? 64 : 3;
    if (bf16)    // tensor cores: [H,64] = act^T [draw | 0]
      LPGNN_TRY(lpgnn_wgrad(act, side ? B.drawb_t : B.drawb_s, rows, H, 64, B.g_tmp, B.scratch, B.scratch_bytes, stream));
    else
      LPGNN_TRY(lpgnn_small_wgrad(act, dt, side ? B.draw_t : B.draw_s, 3, 3, rows, H, B.g_tmp, nullptr, B.scratch,
                                  B.scratch_bytes, stream));
    take_cols_t_kernel<<<(3 * H + 255) / 256, 256, 0, st>>>(B.g_tmp, ld, H, 3, gw);
    LPGNN_LAUNCH_OK();
    count_launches(1);
  }
  if (nh > 0) LPGNN_TRY(hidden_wgrads(nh - 1, B.dpre_s, B.dpre_t));
  }   // LPGNN_BWD_TAIL
  if (!(phases & LPGNN_BWD_REST)) return LPGNN_OK;
  void *dps = B.dpre_s, *dpt = B.dpre_t;
  // ---- hidden layers, last to first
  for (int li = nh - 1; li >= 0; --li) {
    if (li < nh - 1) LPGNN_TRY(hidden_wgrads(li, dps, dpt));
    // data gradients.  dL = A (dPre_t W_rel^{l2r}) + dPre_s W_root^{r2l} = [A dPre_t | dPre_s] [W_rel^{l2r} ; W_root^{r2l}]:
    // aggregate first, then ONE two-operand transform per side (transposed weights prepared by the forward call)
    // whose epilogue applies the ReLU / dropout mask of the layer input.
    // A dPre_t [m,H] and A^T dPre_s [n,H]: the same pair launch on the gradients
    LPGNN_TRY(lpgnn_spmm_pair(rowptr, col, val, m, colptr, row_csc, val_csc, n, -1, dps, dpt, B.dlagg, B.dragg, H, dt, stream));
    lpgnn_epilogue_args ea;
    ea.epilogue = LPGNN_EPI_NONE; ea.dropout_p = 0.f; ea.dropout_seed = 0;
    ea.mask_scale = li > 0 ? scale : 1.f;   // conv1's output has no dropout
    ea.mask_act = B.left[li];
    LPGNN_TRY(lpgnn_node_transform_ex(B.dlagg, H, B.wT[li][0], dps, H, B.wT[li][3], nullptr, m, H, B.droot_s, dt, &ea, stream));
    ea.mask_act = B.right[li];
    LPGNN_TRY(lpgnn_node_transform_ex(B.dragg, H, B.wT[li][2], dpt, H, B.wT[li][1], nullptr, n, H, B.droot_t, dt, &ea, stream));
    // the masked sums are the next dPre; recycle the old dPre buffers as the next layer's outputs
    void* t;
    t = dps; dps = B.droot_s; B.droot_s = t;
    t = dpt; dpt = B.droot_t; B.droot_t = t;
  }
  // ---- conv1: weight gradients only (its inputs are data).  z_t = [A^T x_s | x_t], z_s = [A x_t | x_s]
  if (bf16) {   // tensor cores: [H,64] = dPre^T [agg | x_dst | 1 | 0]; the ones column yields the bias gradient
    for (int side = 0; side < 2; ++side) {
      const int64_t rows = side ? m : n;
      const void* dpre = side ? dps : dpt;
      const int ks = side ? q : p, kd = side ? p : q;
      float* gb = side ? g->c1_r2l_b : g->c1_l2r_b;
      LPGNN_TRY(lpgnn_wgrad(dpre, side ? B.zb_s : B.zb_t, rows, H, 64, B.g_tmp, B.scratch, B.scratch_bytes, stream));
      split_c1_grad_kernel<<<(H * 64 + 255) / 256, 256, 0, st>>>(B.g_tmp, H, ks, kd, side ? g->c1_r2l_wrel : g->c1_l2r_wrel,
                                                                side ? g->c1_r2l_wroot : g->c1_l2r_wroot, gb);
      LPGNN_LAUNCH_OK();
      count_launches(1);
      if (ks + kd == 64) LPGNN_TRY(lpgnn_colsum(dpre, dt, rows, H, gb, B.scratch, B.scratch_bytes, stream));
    }
    return LPGNN_OK;
  }
  const int kt = lpgnn_conv_in_zcat_width(p, q);
  LPGNN_TRY(lpgnn_small_wgrad(dpt, dt, B.z_t, kt, p + q, n, H, B.g_tmp, g->c1_l2r_b, B.scratch, B.scratch_bytes, stream));
  LPGNN_TRY(copy2d(g->c1_l2r_wrel, p, B.g_tmp, p + q, p, H, st));
  LPGNN_TRY(copy2d(g->c1_l2r_wroot, q, B.g_tmp + p, p + q, q, H, st));
  LPGNN_TRY(lpgnn_small_wgrad(dps, dt, B.z_s, kt, p + q, m, H, B.g_tmp, g->c1_r2l_b, B.scratch, B.scratch_bytes, stream));
  LPGNN_TRY(copy2d(g->c1_r2l_wrel, q, B.g_tmp, p + q, q, H, st));
  LPGNN_TRY(copy2d(g->c1_r2l_wroot, p, B.g_tmp + q, p + q, p, H, st));
  return LPGNN_OK;
}
