// One-call basis prediction: graph build -> GCN_FC forward -> basis selection, enqueued from C++.
//
// Replaces the body of the reference's per-LP prediction (scripts/pred_basis.py:113-118 `inference_only`:
// model(batch) + inference_gnn) including the graph construction that the reference does in the DataLoader
// (dataset.py:299-304).  The host cost of a step is ~25 kernel launches from native code instead of ~25 Python
// round trips, which is what bounds small LPs (BASELINE config C5: 10K small/medium LPs).
// All device memory comes from ONE caller-provided workspace, carved with a bump allocator.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace lpgnn {
namespace {

struct Bump {
  char* base;
  size_t off = 0, cap;
  Bump(void* p, size_t c) : base(reinterpret_cast<char*>(p)), cap(c) {}
  template <typename T> T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
};

// The two directions of a layer are independent (arch.py:183-184: `left, right = conv(left, right, edge_index)` computes
// both from the previous layer's features), so the constraint side's kernels can run on a side stream next to the
// variable side's.  Measured on B200: LPs whose grids do not fill the GPU gain (C5, one call per LP: 5.9K -> 7.2K LPs/s);
// at C2 size nothing is gained -- the transforms are power-bound, an idle tail lets the busy SMs clock higher -- and the
// three-LP pipeline loses ~3 % to the extra interleaving.  Hence mode 1 (default) forks only LPs of up to kForkMaxNodes
// nodes; 0 = never, 2 = always.  One side stream and event pair per (device, caller stream), created on first use and
// kept for the life of the process.
constexpr int64_t kForkMaxNodes = 65536;
struct Side { cudaStream_t aux; cudaEvent_t fork, join; };
int g_predict_fork = [] { const char* e = getenv("LPGNN_PREDICT_FORK"); const int v = e ? atoi(e) : 1; return v < 0 ? 0 : (v > 2 ? 2 : v); }();

int side_for(cudaStream_t st, Side* out) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, Side> sides;
  int dev = 0;
  LPGNN_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = sides.find({dev, st});
  if (it == sides.end()) {
    Side s;
    LPGNN_CUDA_OK(cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking));
    LPGNN_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    LPGNN_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    it = sides.emplace(std::make_pair(dev, st), s).first;
  }
  *out = it->second;
  return LPGNN_OK;
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

namespace {

struct Buffers {
  int32_t *rowptr, *colptr, *col, *row_csc, *csr2csc, *status;
  float *val, *val_csc;
  void* build_ws; size_t build_ws_bytes;
  float *z_s, *z_t;          // fp32 [m,KT], [n,KT] (fp32 mode)
  void *zb_s, *zb_t;         // bf16 [m,64], [n,64] (bf16 mode)
  void *act[2][2];           // ping-pong activations: act[i][0] = left [m,H], act[i][1] = right [n,H]
  void *agg_s, *agg_t;       // [m,H], [n,H]
  // fp32 tensor-core mode: x2 operands (half hi / lo) of the aggregate and of the node's own features, row scales
  void *xa_hi[2], *xa_lo[2], *xx_hi[2], *xx_lo[2];   // [0] = constraint side [m,H], [1] = variable side [n,H]
  float *xscale[2], *xscale_x[2];                     // row scales of the aggregate (or of both, split together) / of x
  float* wabs;
  float *part_s, *part_t;    // fused-head partials
  float *logit_s, *logit_t;  // [m,3], [n,3]
  void* sel_ws; size_t sel_ws_bytes;
};

size_t carve(Bump& b, Buffers& B, int64_t z, int32_t m, int32_t n, int32_t p, int32_t q, int32_t H, int32_t depth,
             int bf16, int x2) {
  const size_t es = bf16 ? 2 : 4;
  B.rowptr = b.take<int32_t>((size_t)m + 1);
  B.colptr = b.take<int32_t>((size_t)n + 1);
  B.col = b.take<int32_t>(z); B.row_csc = b.take<int32_t>(z); B.csr2csc = b.take<int32_t>(z);
  B.val = b.take<float>(z); B.val_csc = b.take<float>(z);
  B.status = b.take<int32_t>(4);
  B.build_ws_bytes = lpgnn_graph_build_workspace_bytes(z, m, n);
  B.build_ws = b.take<char>(B.build_ws_bytes);
  const int kt = lpgnn_conv_in_zcat_width(p, q);
  if (bf16) {
    B.zb_s = b.take<char>((size_t)m * 64 * 2); B.zb_t = b.take<char>((size_t)n * 64 * 2);
    B.z_s = B.z_t = nullptr;
  } else {
    B.z_s = b.take<float>((size_t)m * kt); B.z_t = b.take<float>((size_t)n * kt);
    B.zb_s = B.zb_t = nullptr;
  }
  const int n_act = depth > 3 ? 2 : 1;
  for (int i = 0; i < 2; ++i) {
    B.act[i][0] = i < n_act ? b.take<char>((size_t)m * H * es) : nullptr;
    B.act[i][1] = i < n_act ? b.take<char>((size_t)n * H * es) : nullptr;
  }
  if (!bf16 && !x2 && depth > 2 && n_act == 1) {  // CUDA-core fp32 mode has no fused head: the last layer's output is materialised
    B.act[1][0] = b.take<char>((size_t)m * H * es);
    B.act[1][1] = b.take<char>((size_t)n * H * es);
  }
  if (depth > 2) {
    B.agg_s = b.take<char>((size_t)m * H * es); B.agg_t = b.take<char>((size_t)n * H * es);
  } else {
    B.agg_s = B.agg_t = nullptr;
  }
  for (int i = 0; i < 2; ++i) {
    const size_t rows = i ? n : m;
    const bool on = x2 && depth > 2;
    B.xa_hi[i] = on ? b.take<char>(rows * H * 2) : nullptr; B.xa_lo[i] = on ? b.take<char>(rows * H * 2) : nullptr;
    B.xx_hi[i] = on ? b.take<char>(rows * H * 2) : nullptr; B.xx_lo[i] = on ? b.take<char>(rows * H * 2) : nullptr;
    B.xscale[i] = on ? b.take<float>(rows) : nullptr;
    B.xscale_x[i] = on ? b.take<float>(rows) : nullptr;
  }
  B.wabs = b.take<float>(2 * 80);
  const int nparts = lpgnn_node_transform_head_parts(H);
  B.part_s = b.take<float>((size_t)nparts * m * 3); B.part_t = b.take<float>((size_t)nparts * n * 3);
  B.logit_s = b.take<float>((size_t)m * 3); B.logit_t = b.take<float>((size_t)n * 3);
  B.sel_ws_bytes = lpgnn_basis_select_workspace_bytes((int64_t)m + n);
  B.sel_ws = b.take<char>(B.sel_ws_bytes);
  return align_up(b.off, 256);
}

}  // namespace

extern "C" size_t lpgnn_predict_workspace_bytes(int64_t nnz, int32_t m, int32_t n, int32_t p, int32_t q, int32_t hids,
                                                int32_t depth, int precision) {
  Bump b(nullptr, ~(size_t)0);
  Buffers B;
  return carve(b, B, nnz > 0 ? nnz : 1, m, n, p, q, hids, depth, is_16bit(precision & 15),
               (precision & LPGNN_WS_X2) != 0);
}

extern "C" int lpgnn_predict_basis(const lpgnn_gcn_fc_weights* w, const int32_t* coo_row, const int32_t* coo_col,
                                   const float* coo_val, int64_t nnz, int32_t m, int32_t n, int flags, const float* x_s,
                                   const float* x_t, uint8_t* status_out, float* logits_out, int32_t* graph_status,
                                   void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  return lpgnn_predict_basis_packed(w, coo_row, coo_col, coo_val, nnz, m, n, flags, x_s, x_t, nullptr, nullptr, 0,
                                    status_out, logits_out, graph_status, workspace, workspace_bytes, stream);
}

extern "C" int lpgnn_predict_basis_packed(const lpgnn_gcn_fc_weights* w, const int32_t* coo_row,
                                          const int32_t* coo_col, const float* coo_val, int64_t nnz, int32_t m,
                                          int32_t n, int flags, const float* x_s, const float* x_t,
                                          const int32_t* cons_ptr, const int32_t* vars_ptr, int32_t n_segments,
                                          uint8_t* status_out, float* logits_out, int32_t* graph_status,
                                          void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(w && m > 0 && n > 0 && nnz >= 0 && x_s && x_t && status_out && workspace, "predict_basis: bad arguments");
  const int H = w->hids, depth = w->depth, p = w->p, q = w->q;
  const int bf16 = is_16bit(w->precision);   // 16-bit storage (bf16 or IEEE half): the tensor-core path
  LPGNN_REQUIRE(dtype_ok(w->precision), "predict_basis: bad precision %d", w->precision);
  LPGNN_REQUIRE(depth >= 2 && depth - 2 <= LPGNN_MAX_HIDDEN_LAYERS, "predict_basis: depth %d unsupported", depth);
  LPGNN_REQUIRE(!bf16 || H % 64 == 0, "predict_basis: bf16 mode needs hids %% 64 == 0");
  const int x2 = !bf16 && depth > 2 && H % 64 == 0 && w->l2r_wrel_hi[0] != nullptr;
  const size_t need = lpgnn_predict_workspace_bytes(nnz, m, n, p, q, H, depth, w->precision | (x2 ? LPGNN_WS_X2 : 0));
  if (workspace_bytes < need) { set_error("predict_basis: workspace %zu < required %zu", workspace_bytes, need); return LPGNN_EWORKSPACE; }
  LPGNN_REQUIRE((uintptr_t)workspace % 256 == 0, "predict_basis: workspace must be 256-byte aligned");
  Bump b(workspace, workspace_bytes);
  Buffers B;
  carve(b, B, nnz > 0 ? nnz : 1, m, n, p, q, H, depth, bf16, x2);
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* gstat = graph_status ? graph_status : B.status;
  if (!graph_status) LPGNN_CUDA_OK(cudaMemsetAsync(B.status, 0, sizeof(int32_t), st));

#define LPGNN_TRY(expr) do { if (int _rc = (expr)) return _rc; } while (0)
  // ---- (a1) graph
  LPGNN_TRY(lpgnn_graph_build(coo_row, coo_col, 0, coo_val, nnz, m, n, flags, B.rowptr, B.col, B.val, B.colptr, B.row_csc,
                              B.val_csc, B.csr2csc, gstat, B.build_ws, B.build_ws_bytes, stream));
  const int dt = bf16 ? w->precision : LPGNN_F32;
  void *left = B.act[0][0], *right = B.act[0][1];
  // The constraint side's chain (conv1 -> aggregate -> transform -> head) runs on the side stream (see Side above), the
  // variable side's on `stream`; the chains meet where a layer reads the other side's features (LPGNN_CROSS).
  Side side{};
  const bool fork = depth > 2 && (bf16 || x2) && (g_predict_fork == 2 || (g_predict_fork == 1 && (int64_t)m + n <= kForkMaxNodes));
  if (fork) LPGNN_TRY(side_for(st, &side));
  lpgnn_stream_t st2 = fork ? (lpgnn_stream_t)side.aux : stream;
#define LPGNN_FORK() do { if (fork) { LPGNN_CUDA_OK(cudaEventRecord(side.fork, st)); LPGNN_CUDA_OK(cudaStreamWaitEvent(side.aux, side.fork, 0)); } } while (0)
#define LPGNN_JOIN() do { if (fork) { LPGNN_CUDA_OK(cudaEventRecord(side.join, side.aux)); LPGNN_CUDA_OK(cudaStreamWaitEvent(st, side.join, 0)); } } while (0)
#define LPGNN_CROSS() do { if (fork) { LPGNN_CUDA_OK(cudaEventRecord(side.fork, st)); LPGNN_CUDA_OK(cudaEventRecord(side.join, side.aux)); \
    LPGNN_CUDA_OK(cudaStreamWaitEvent(side.aux, side.fork, 0)); LPGNN_CUDA_OK(cudaStreamWaitEvent(st, side.join, 0)); } } while (0)
  LPGNN_FORK();
  // ---- conv1 (+relu).  CSC view: dst = variables, src = constraints; CSR view: dst = constraints, src = variables
  if (bf16 && p == 8 && q == 8 && H % 32 == 0 && H <= 4096) {   // the reference's shape: one kernel per direction
    if (fork) {   // small LPs: the sides run next to each other on two streams
      LPGNN_TRY(lpgnn_conv_in_16(B.colptr, B.row_csc, B.val_csc, n, x_s, x_t, w->c1_l2r_wrel, w->c1_l2r_b, w->c1_l2r_wroot, H,
                                 right, dt, LPGNN_EPI_RELU, nullptr, stream));
      LPGNN_TRY(lpgnn_conv_in_16(B.rowptr, B.col, B.val, m, x_t, x_s, w->c1_r2l_wrel, w->c1_r2l_b, w->c1_r2l_wroot, H, left, dt,
                                 LPGNN_EPI_RELU, nullptr, st2));
    } else {      // one launch for both directions
      LPGNN_TRY(lpgnn_conv_in_16_pair(B.rowptr, B.col, B.val, B.colptr, B.row_csc, B.val_csc, m, n, x_s, x_t, w->c1_l2r_wrel,
                                      w->c1_l2r_b, w->c1_l2r_wroot, w->c1_r2l_wrel, w->c1_r2l_b, w->c1_r2l_wroot, H, left, right,
                                      dt, LPGNN_EPI_RELU, nullptr, nullptr, stream));
    }
  } else if (bf16) {
    LPGNN_TRY(lpgnn_gather_cat_ex(B.colptr, B.row_csc, B.val_csc, n, x_s, p, x_t, q, nullptr, B.zb_t, dt, stream));
    LPGNN_TRY(lpgnn_gather_cat_ex(B.rowptr, B.col, B.val, m, x_t, q, x_s, p, nullptr, B.zb_s, dt, st2));
    LPGNN_TRY(lpgnn_node_transform(B.zb_t, 64, w->c1_l2r_wcat, nullptr, 0, nullptr, w->c1_l2r_b, n, H, right, dt, dt,
                                   LPGNN_EPI_RELU, stream));
    LPGNN_TRY(lpgnn_node_transform(B.zb_s, 64, w->c1_r2l_wcat, nullptr, 0, nullptr, w->c1_r2l_b, m, H, left, dt, dt,
                                   LPGNN_EPI_RELU, st2));
  } else if (x2) {   // fp32 input layer that also emits its output as x2 operands (lin_root side of the first hidden layer)
    LPGNN_TRY(lpgnn_conv_in_fused_x2(B.colptr, B.row_csc, B.val_csc, n, x_s, p, x_t, q, w->c1_l2r_wrel, w->c1_l2r_b,
                                     w->c1_l2r_wroot, H, (float*)right, LPGNN_EPI_RELU, B.z_t, B.xx_hi[1], B.xx_lo[1],
                                     B.xscale_x[1], B.wabs, stream));
    LPGNN_TRY(lpgnn_conv_in_fused_x2(B.rowptr, B.col, B.val, m, x_t, q, x_s, p, w->c1_r2l_wrel, w->c1_r2l_b,
                                     w->c1_r2l_wroot, H, (float*)left, LPGNN_EPI_RELU, B.z_s, B.xx_hi[0], B.xx_lo[0],
                                     B.xscale_x[0], B.wabs + 80, st2));
  } else {
    LPGNN_TRY(lpgnn_conv_in_fused(B.colptr, B.row_csc, B.val_csc, n, x_s, p, x_t, q, w->c1_l2r_wrel, w->c1_l2r_b,
                                  w->c1_l2r_wroot, H, right, dt, LPGNN_EPI_RELU, B.z_t, stream));
    LPGNN_TRY(lpgnn_conv_in_fused(B.rowptr, B.col, B.val, m, x_t, q, x_s, p, w->c1_r2l_wrel, w->c1_r2l_b,
                                  w->c1_r2l_wroot, H, left, dt, LPGNN_EPI_RELU, B.z_s, stream));
  }
  // ---- hidden layers
  const int n_hidden = depth - 2;
  bool head_done = false;
  int cur = 0;
  for (int li = 0; li < n_hidden; ++li) {
    const bool x2_direct = x2 && li == 0;    // first hidden layer: its inputs (conv1 outputs) already exist as x2 operands
    LPGNN_CROSS();                           // both sides' features of the previous layer are complete (and were consumed)
    if (x2_direct && !fork) {   // both aggregations in one launch, straight into x2 operands (see below)
      LPGNN_TRY(lpgnn_spmm_x2_pair(B.rowptr, B.col, B.val, m, B.colptr, B.row_csc, B.val_csc, n, nnz, (const float*)left,
                                   (const float*)right, H, B.xscale_x[0], B.xscale_x[1], B.xa_hi[0], B.xa_lo[0], B.xscale[0],
                                   B.xa_hi[1], B.xa_lo[1], B.xscale[1], (float*)B.agg_s, (float*)B.agg_t, stream));
    } else if (x2_direct) {   // aggregate straight into x2 operands (scale from the sources' scales), no fp32 aggregate, no split pass
      LPGNN_TRY(lpgnn_spmm_x2(B.colptr, B.row_csc, B.val_csc, n, (const float*)left, H, B.xscale_x[0], B.xa_hi[1], B.xa_lo[1],
                              B.xscale[1], (float*)B.agg_t, stream));                          // A^T . left
      LPGNN_TRY(lpgnn_spmm_x2(B.rowptr, B.col, B.val, m, (const float*)right, H, B.xscale_x[1], B.xa_hi[0], B.xa_lo[0],
                              B.xscale[0], (float*)B.agg_s, st2));                             // A   . right
    } else if (!fork) {       // A^T . left and A . right in one launch where both take the banded sweep
      LPGNN_TRY(lpgnn_spmm_pair(B.rowptr, B.col, B.val, m, B.colptr, B.row_csc, B.val_csc, n, nnz, left, right, B.agg_s, B.agg_t,
                                H, dt, stream));
    } else {
      LPGNN_TRY(lpgnn_spmm(B.colptr, B.row_csc, B.val_csc, n, left, B.agg_t, H, dt, stream));    // A^T . left
      LPGNN_TRY(lpgnn_spmm(B.rowptr, B.col, B.val, m, right, B.agg_s, H, dt, st2));              // A   . right
    }
    const bool last = li == n_hidden - 1;
    if (last && bf16) {  // head fused into the epilogue; the last activation never reaches HBM
      const int nparts = lpgnn_node_transform_head_parts(H);
      LPGNN_TRY(lpgnn_node_transform_head_ex(B.agg_t, H, w->l2r_wrel[li], right, H, w->l2r_wroot[li], w->l2r_b[li], n, H,
                                             nullptr, dt, LPGNN_EPI_RELU, w->head_right_w, B.part_t, stream));
      LPGNN_TRY(lpgnn_node_transform_head_ex(B.agg_s, H, w->r2l_wrel[li], left, H, w->r2l_wroot[li], w->r2l_b[li], m, H,
                                             nullptr, dt, LPGNN_EPI_RELU, w->head_left_w, B.part_s, st2));
      LPGNN_TRY(lpgnn_head_finish(B.part_s, nparts, m, w->head_left_b, x_s, p, B.logit_s, st2));
      LPGNN_TRY(lpgnn_head_finish(B.part_t, nparts, n, w->head_right_b, x_t, q, B.logit_t, stream));
      head_done = true;
    } else if (x2) {  // the reference's default precision on the tensor cores: three half x half passes over x2 operands
      if (!x2_direct) {   // deeper layers: inputs are fp32 outputs of the previous transform -> split them (shared row scale)
        LPGNN_TRY(lpgnn_split_x2((const float*)B.agg_s, H, (const float*)left, H, m, B.xa_hi[0], B.xa_lo[0], B.xx_hi[0],
                                 B.xx_lo[0], B.xscale[0], st2));
        LPGNN_TRY(lpgnn_split_x2((const float*)B.agg_t, H, (const float*)right, H, n, B.xa_hi[1], B.xa_lo[1], B.xx_hi[1],
                                 B.xx_lo[1], B.xscale[1], stream));
      }
      const float* rsx_t = x2_direct ? B.xscale_x[1] : nullptr;   // row scales of the lin_root operand (null: shared)
      const float* rsx_s = x2_direct ? B.xscale_x[0] : nullptr;
      void *nl = last ? nullptr : B.act[cur ^ 1][0], *nr = last ? nullptr : B.act[cur ^ 1][1];
      LPGNN_TRY(lpgnn_node_transform_x2(B.xa_hi[1], B.xa_lo[1], H, w->l2r_wrel_hi[li], w->l2r_wrel_lo[li], B.xx_hi[1],
                                        B.xx_lo[1], H, w->l2r_wroot_hi[li], w->l2r_wroot_lo[li], B.xscale[1], rsx_t,
                                        w->l2r_wscale[li], w->l2r_b[li], n, H, (float*)nr, LPGNN_EPI_RELU,
                                        last ? w->head_right_w : nullptr, last ? B.part_t : nullptr, stream));
      LPGNN_TRY(lpgnn_node_transform_x2(B.xa_hi[0], B.xa_lo[0], H, w->r2l_wrel_hi[li], w->r2l_wrel_lo[li], B.xx_hi[0],
                                        B.xx_lo[0], H, w->r2l_wroot_hi[li], w->r2l_wroot_lo[li], B.xscale[0], rsx_s,
                                        w->r2l_wscale[li], w->r2l_b[li], m, H, (float*)nl, LPGNN_EPI_RELU,
                                        last ? w->head_left_w : nullptr, last ? B.part_s : nullptr, st2));
      const int nparts = lpgnn_node_transform_head_parts(H);
      if (last) {
        LPGNN_TRY(lpgnn_head_finish(B.part_s, nparts, m, w->head_left_b, x_s, p, B.logit_s, st2));
        LPGNN_TRY(lpgnn_head_finish(B.part_t, nparts, n, w->head_right_b, x_t, q, B.logit_t, stream));
        head_done = true;
      } else {
        left = nl; right = nr; cur ^= 1;
      }
    } else {
      void *nl = B.act[cur ^ 1][0], *nr = B.act[cur ^ 1][1];
      LPGNN_TRY(lpgnn_node_transform(B.agg_t, H, w->l2r_wrel[li], right, H, w->l2r_wroot[li], w->l2r_b[li], n, H, nr, dt,
                                     dt, LPGNN_EPI_RELU, stream));
      LPGNN_TRY(lpgnn_node_transform(B.agg_s, H, w->r2l_wrel[li], left, H, w->r2l_wroot[li], w->r2l_b[li], m, H, nl, dt,
                                     dt, LPGNN_EPI_RELU, st2));
      left = nl; right = nr; cur ^= 1;
    }
  }
  if (!head_done) {
    LPGNN_TRY(lpgnn_head_mask(left, dt, m, H, w->head_left_w, w->head_left_b, x_s, p, B.logit_s, nullptr, st2));
    LPGNN_TRY(lpgnn_head_mask(right, dt, n, H, w->head_right_w, w->head_right_b, x_t, q, B.logit_t, nullptr, stream));
  }
  LPGNN_JOIN();
  if (logits_out) {
    LPGNN_CUDA_OK(cudaMemcpyAsync(logits_out, B.logit_s, sizeof(float) * 3 * (size_t)m, cudaMemcpyDeviceToDevice, st));
    LPGNN_CUDA_OK(cudaMemcpyAsync(logits_out + 3 * (size_t)m, B.logit_t, sizeof(float) * 3 * (size_t)n,
                                  cudaMemcpyDeviceToDevice, st));
  }
  // ---- (a6) basis decision
  if (n_segments > 0) {
    LPGNN_REQUIRE(cons_ptr && vars_ptr, "predict_basis_packed: null segment pointers");
    LPGNN_TRY(lpgnn_basis_select_segmented_ex(B.logit_s, B.logit_t, cons_ptr, vars_ptr, n_segments, m, n, status_out, 0,
                                              (flags & LPGNN_STATUS_LP_MAJOR) ? 1 : 0, B.sel_ws, B.sel_ws_bytes, stream));
  } else {
    LPGNN_TRY(lpgnn_basis_select(B.logit_s, m, B.logit_t, n, m, status_out, 0, nullptr, B.sel_ws, B.sel_ws_bytes, stream));
  }
#undef LPGNN_TRY
#undef LPGNN_FORK
#undef LPGNN_JOIN
#undef LPGNN_CROSS
  return LPGNN_OK;
}

extern "C" int lpgnn_set_predict_fork(int mode) {
  const int prev = lpgnn::g_predict_fork;
  lpgnn::g_predict_fork = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return prev;
}
