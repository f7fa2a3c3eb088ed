// Balanced cross-entropy of the reference's training loop, forward value and d(loss)/d(logits) in one pass.
//
// Replaces balanced() (reference train.py:39-46) with the class weights of labels_to_balanced_weights
// (utils.py:286-299) and torch.nn.CrossEntropyLoss(weight=...) (weighted mean):
//   w_side[c]  = total / cnt[c] for the classes that occur (0 otherwise); unless exactly two classes occur and with
//                merge_lu, w[0] = w[2] = (w[0] + w[2]) / 2
//   CE_side    = sum_i w[y_i] * (-log softmax(x_i)[y_i]) / sum_i w[y_i]
//   loss       = (m+n)/m * CE_cons + (m+n)/n * CE_vars
//   dlogits[i] = coef_side * w[y_i] / sum_j w[y_j] * (softmax(x_i) - onehot(y_i))
// ~30 framework launches (unique, where, log_softmax, nll_loss, ... and their backward) become 2 kernels + 1 memset.
// The other two losses of train.py (unbalanced, focal) and the counters behind val.accuracy follow below.
// Integer atomics only (class counts, block counter); floating-point sums run in a fixed order: bit-reproducible.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;

struct CeWs {
  int cnt[2][4];          // class counts per side (3 used)
  unsigned int done;      // blocks finished (last-block reduction)
  unsigned int pad[7];
};

__global__ void __launch_bounds__(kThreads)
ce_count_kernel(const int64_t* __restrict__ y_s, int m, const int64_t* __restrict__ y_t, int n, int blocks_s, CeWs* ws) {
  __shared__ int c[3];
  if (threadIdx.x < 3) c[threadIdx.x] = 0;
  __syncthreads();
  const int side = blockIdx.x >= blocks_s;
  const int64_t* y = side ? y_t : y_s;
  const int rows = side ? n : m;
  const int i = (blockIdx.x - (side ? blocks_s : 0)) * kThreads + threadIdx.x;
  if (i < rows) {
    const int64_t l = y[i];
    if (l >= 0 && l < 3) atomicAdd(&c[(int)l], 1);
  }
  __syncthreads();
  if (threadIdx.x < 3 && c[threadIdx.x]) atomicAdd(&ws->cnt[side][threadIdx.x], c[threadIdx.x]);
}

__device__ __forceinline__ void class_weights(const int* cnt, int merge_lu, float* w, float* denom) {
  const int total = cnt[0] + cnt[1] + cnt[2];
  int present = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    w[c] = cnt[c] > 0 ? (float)total / (float)cnt[c] : 0.f;
    present += cnt[c] > 0;
  }
  if (merge_lu && present != 2) w[0] = w[2] = (w[0] + w[2]) / 2.f;
  *denom = (float)cnt[0] * w[0] + (float)cnt[1] * w[1] + (float)cnt[2] * w[2];   // = sum_i w[y_i]
}

__global__ void __launch_bounds__(kThreads)
ce_loss_kernel(const float* __restrict__ x_s, const int64_t* __restrict__ y_s, int m, const float* __restrict__ x_t,
               const int64_t* __restrict__ y_t, int n, int blocks_s, int merge_lu, CeWs* ws, float* __restrict__ partials,
               float* __restrict__ d_s, float* __restrict__ d_t, float* __restrict__ loss_out) {
  const int side = blockIdx.x >= blocks_s;
  const float* x = side ? x_t : x_s;
  const int64_t* y = side ? y_t : y_s;
  float* d = side ? d_t : d_s;
  const int rows = side ? n : m;
  float w[3], denom;
  class_weights(ws->cnt[side], merge_lu, w, &denom);
  const float coef = (float)(m + n) / (float)rows;
  const int i = (blockIdx.x - (side ? blocks_s : 0)) * kThreads + threadIdx.x;
  float part = 0.f;
  if (i < rows) {
    const float a = x[3 * (int64_t)i], b = x[3 * (int64_t)i + 1], c = x[3 * (int64_t)i + 2];
    const float mx = fmaxf(a, fmaxf(b, c));
    const float ea = expf(a - mx), eb = expf(b - mx), ec = expf(c - mx);
    const float se = ea + eb + ec;
    const float lse = mx + logf(se);
    const int64_t l = y[i];
    const bool ok = l >= 0 && l < 3;
    const float wi = ok ? w[(int)l] : 0.f;
    const float xl = l == 0 ? a : (l == 1 ? b : c);
    part = ok ? wi * (lse - xl) : 0.f;
    if (d) {
      const float g = coef * wi / denom, inv = 1.f / se;
      d[3 * (int64_t)i] = g * (ea * inv - (l == 0 ? 1.f : 0.f));
      d[3 * (int64_t)i + 1] = g * (eb * inv - (l == 1 ? 1.f : 0.f));
      d[3 * (int64_t)i + 2] = g * (ec * inv - (l == 2 ? 1.f : 0.f));
    }
  }
  // block sum, fixed order
  __shared__ float red[kThreads / 32];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) s += red[k];
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // last block: both sides' partials summed in index order (double), then the weighted means
  __shared__ double acc[2][kThreads];
  for (int sd = 0; sd < 2; ++sd) {
    const int b0 = sd ? blocks_s : 0, b1 = sd ? (int)gridDim.x : blocks_s;
    const int per = (b1 - b0 + kThreads - 1) / kThreads;
    double s = 0.0;
    for (int k = 0; k < per; ++k) {
      const int b = b0 + threadIdx.x * per + k;
      if (b < b1) s += (double)__ldcg(partials + b);
    }
    acc[sd][threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double loss = 0.0;
    for (int sd = 0; sd < 2; ++sd) {
      double s = 0.0;
      for (int k = 0; k < kThreads; ++k) s += acc[sd][k];
      float ww[3], dn;
      class_weights(ws->cnt[sd], merge_lu, ww, &dn);
      loss += (double)((float)(m + n) / (float)(sd ? n : m)) * (s / (double)dn);
    }
    *loss_out = (float)loss;
  }
}

// ------------------------------------------------------------------------------------------------ flat CE / focal
// unbalanced() (reference train.py:30-37): F.cross_entropy over the concatenation of both sides, plain mean.
// focal() (train.py:18-28, 49-53): the SAME mean CE pushed through the focal factor as a scalar,
//   ce = mean_i(-log softmax(x_i)[y_i]),  pt = exp(-ce),  loss = (1 - pt)^gamma * ce
// (the reference applies the factor to the batch-mean CE, reduction='mean', not per sample).  One kernel: per-row CE
// and the un-scaled gradient (softmax - onehot) / (m+n); the last block sums the block partials in index order
// (double) and writes loss and d(loss)/d(ce), which the caller folds into the upstream gradient.
struct FlatWs {
  unsigned int done;
  unsigned int pad[3];
};

__global__ void __launch_bounds__(kThreads)
ce_flat_kernel(const float* __restrict__ x_s, const int64_t* __restrict__ y_s, int m, const float* __restrict__ x_t,
               const int64_t* __restrict__ y_t, int n, int focal, float gamma, FlatWs* ws, float* __restrict__ partials,
               float* __restrict__ d_s, float* __restrict__ d_t, float* __restrict__ loss_out,
               float* __restrict__ scale_out) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t total = (int64_t)m + n;
  float part = 0.f;
  if (i < total) {
    const bool side = i >= m;
    const int64_t r = side ? i - m : i;
    const float* x = (side ? x_t : x_s) + 3 * r;
    const int64_t l = (side ? y_t : y_s)[r];
    const float a = x[0], b = x[1], c = x[2];
    const float mx = fmaxf(a, fmaxf(b, c));
    const float ea = expf(a - mx), eb = expf(b - mx), ec = expf(c - mx);
    const float se = ea + eb + ec;
    const bool ok = l >= 0 && l < 3;
    part = ok ? (mx + logf(se)) - (l == 0 ? a : (l == 1 ? b : c)) : 0.f;
    float* d = side ? d_t : d_s;
    if (d) {
      const float g = ok ? 1.f / (float)total : 0.f, inv = 1.f / se;
      d[3 * r] = g * (ea * inv - (l == 0 ? 1.f : 0.f));
      d[3 * r + 1] = g * (eb * inv - (l == 1 ? 1.f : 0.f));
      d[3 * r + 2] = g * (ec * inv - (l == 2 ? 1.f : 0.f));
    }
  }
  __shared__ float red[kThreads / 32];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) s += red[k];
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ double acc[kThreads];
  const int per = ((int)gridDim.x + kThreads - 1) / kThreads;
  double s = 0.0;
  for (int k = 0; k < per; ++k) {
    const int b = threadIdx.x * per + k;
    if (b < (int)gridDim.x) s += (double)__ldcg(partials + b);
  }
  acc[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < kThreads; ++k) t += acc[k];
    const float ce = (float)(t / (double)total);
    float loss = ce, scale = 1.f;
    if (focal) {
      const float pt = expf(-ce), q = 1.f - pt;
      const float qg = gamma == 2.f ? q * q : powf(q, gamma);
      const float qg1 = gamma == 2.f ? q : powf(q, gamma - 1.f);
      loss = qg * ce;
      scale = gamma * qg1 * pt * ce + qg;          // d loss / d ce
    }
    *loss_out = loss;
    *scale_out = scale;
  }
}

// ------------------------------------------------------------------------------------------------ accuracy counters
// accuracy() (reference val.py:199-237) needs, per side, #correct and the class-1 confusion counts (sklearn
// precision / recall with labels=[1]): counts[side*4 + {0: pred == gt, 1: pred == 1 and gt == 1, 2: pred == 1, 3: gt == 1}].
template <typename StatusT>
__global__ void __launch_bounds__(kThreads)
basis_metrics_kernel(const StatusT* __restrict__ status, const int64_t* __restrict__ y_s, int m, const int64_t* __restrict__ y_t,
                     int n, int32_t* __restrict__ counts) {
  __shared__ int c[8];
  if (threadIdx.x < 8) c[threadIdx.x] = 0;
  __syncthreads();
  const int64_t total = (int64_t)m + n;
  int loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int side = i >= m;
    const int64_t gt = side ? y_t[i - m] : y_s[i];
    const int64_t pr = (int64_t)status[i];
    loc[side * 4 + 0] += pr == gt;
    loc[side * 4 + 1] += pr == 1 && gt == 1;
    loc[side * 4 + 2] += pr == 1;
    loc[side * 4 + 3] += gt == 1;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int v = loc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&c[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 8 && c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], c[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ packed mini-batches
// balanced() for a block-diagonal PACK of LPs (north star: "mini-batches of LP graphs"; the reference trains one LP per
// step, train.py:70): every LP b keeps its OWN class weights and (m_b+n_b)/m_b, (m_b+n_b)/n_b factors (train.py:39-46
// applied per graph), the pack's loss is the mean over its LPs -- the gradient is the average of the per-LP gradients,
// i.e. what one optimiser step over B graphs (or B data-parallel ranks) would apply.  One CTA per (LP, side): class
// counts, weights, the weighted CE (fixed-order block reduction) and the logit gradients; the last CTA to finish adds
// the per-(LP, side) terms in index order.  Deterministic.
struct SegWs {
  unsigned int done;
  unsigned int pad[3];
};

__global__ void __launch_bounds__(kThreads)
ce_segmented_kernel(const float* __restrict__ x_s, const int64_t* __restrict__ y_s, const int32_t* __restrict__ cptr,
                    const float* __restrict__ x_t, const int64_t* __restrict__ y_t, const int32_t* __restrict__ vptr,
                    int n_seg, int merge_lu, SegWs* ws, float* __restrict__ terms /*[n_seg][2]*/, float* __restrict__ d_s,
                    float* __restrict__ d_t, float* __restrict__ loss_out) {
  const int b = blockIdx.x >> 1, side = blockIdx.x & 1;
  const int32_t c0 = cptr[b], mb = cptr[b + 1] - c0, v0 = vptr[b], nb = vptr[b + 1] - v0;
  const float* x = side ? x_t + 3 * (int64_t)v0 : x_s + 3 * (int64_t)c0;
  const int64_t* y = side ? y_t + v0 : y_s + c0;
  float* d = side ? (d_t ? d_t + 3 * (int64_t)v0 : nullptr) : (d_s ? d_s + 3 * (int64_t)c0 : nullptr);
  const int rows = side ? nb : mb;
  __shared__ int cnt[4];
  __shared__ float red[kThreads / 32];
  __shared__ bool last;
  if (threadIdx.x < 4) cnt[threadIdx.x] = 0;
  __syncthreads();
  int c[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < rows; i += kThreads) { const int64_t l = y[i]; if (l >= 0 && l < 3) ++c[(int)l]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[k], v);      // integer: order-free
  }
  __syncthreads();
  float w[3], denom;
  class_weights(cnt, merge_lu, w, &denom);
  const float coef = rows > 0 ? (float)(mb + nb) / (float)rows : 0.f;
  const float inv_b = 1.f / (float)n_seg;
  float part = 0.f;
  for (int i = threadIdx.x; i < rows; i += kThreads) {
    const float a = x[3 * (int64_t)i], bb = x[3 * (int64_t)i + 1], cc = x[3 * (int64_t)i + 2];
    const float mx = fmaxf(a, fmaxf(bb, cc));
    const float ea = expf(a - mx), eb = expf(bb - mx), ec = expf(cc - mx);
    const float se = ea + eb + ec;
    const int64_t l = y[i];
    const bool ok = l >= 0 && l < 3;
    const float wi = ok ? w[(int)l] : 0.f;
    part += ok ? wi * ((mx + logf(se)) - (l == 0 ? a : (l == 1 ? bb : cc))) : 0.f;
    if (d) {
      const float g = denom > 0.f ? inv_b * coef * wi / denom : 0.f, inv = 1.f / se;
      d[3 * (int64_t)i] = g * (ea * inv - (l == 0 ? 1.f : 0.f));
      d[3 * (int64_t)i + 1] = g * (eb * inv - (l == 1 ? 1.f : 0.f));
      d[3 * (int64_t)i + 2] = g * (ec * inv - (l == 2 ? 1.f : 0.f));
    }
  }
  // block sum in a fixed order (per-thread strided partials, then lanes, then warps)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) sum += red[k];
    terms[blockIdx.x] = denom > 0.f ? coef * sum / denom : 0.f;
    __threadfence();
    last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double total = 0.0;
  for (int k = 0; k < 2 * n_seg; ++k) total += (double)__ldcg(terms + k);
  *loss_out = (float)(total / (double)n_seg);
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" size_t lpgnn_balanced_ce_workspace_bytes(int32_t m, int32_t n) {
  return sizeof(CeWs) + (size_t)(ceil_div(m, kThreads) + ceil_div(n, kThreads)) * sizeof(float);
}

extern "C" int lpgnn_balanced_ce(const float* logits_s, const int64_t* y_s, int32_t m, const float* logits_t,
                                 const int64_t* y_t, int32_t n, int merge_lu, float* loss_out, float* dlogits_s,
                                 float* dlogits_t, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m > 0 && n > 0, "balanced_ce: m=%d, n=%d must be positive", m, n);
  LPGNN_REQUIRE(logits_s && y_s && logits_t && y_t && loss_out && workspace, "balanced_ce: null pointer");
  LPGNN_REQUIRE((dlogits_s == nullptr) == (dlogits_t == nullptr), "balanced_ce: pass both gradient outputs or neither");
  LPGNN_REQUIRE((uintptr_t)workspace % 16 == 0, "balanced_ce: workspace must be 16-byte aligned");
  if (workspace_bytes < lpgnn_balanced_ce_workspace_bytes(m, n)) {
    set_error("balanced_ce: workspace too small");
    return LPGNN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  CeWs* ws = reinterpret_cast<CeWs*>(workspace);
  float* partials = reinterpret_cast<float*>(ws + 1);
  const int bs = ceil_div(m, kThreads), bt = ceil_div(n, kThreads);
  LPGNN_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(CeWs), st));
  ce_count_kernel<<<bs + bt, kThreads, 0, st>>>(y_s, m, y_t, n, bs, ws);
  ce_loss_kernel<<<bs + bt, kThreads, 0, st>>>(logits_s, y_s, m, logits_t, y_t, n, bs, merge_lu, ws, partials, dlogits_s,
                                              dlogits_t, loss_out);
  LPGNN_LAUNCH_OK();
  count_launches(2);
  return LPGNN_OK;
}

extern "C" size_t lpgnn_flat_ce_workspace_bytes(int32_t m, int32_t n) {
  return sizeof(FlatWs) + (size_t)ceil_div((int64_t)m + n, kThreads) * sizeof(float);
}

extern "C" int lpgnn_flat_ce(const float* logits_s, const int64_t* y_s, int32_t m, const float* logits_t, const int64_t* y_t,
                             int32_t n, int focal, float gamma, float* loss_out, float* grad_scale_out, float* dlogits_s,
                             float* dlogits_t, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m >= 0 && n >= 0 && (int64_t)m + n > 0, "flat_ce: m=%d, n=%d (need m + n > 0)", m, n);
  LPGNN_REQUIRE((m == 0 || (logits_s && y_s)) && (n == 0 || (logits_t && y_t)) && loss_out && grad_scale_out && workspace,
                "flat_ce: null pointer");
  LPGNN_REQUIRE((dlogits_s == nullptr || m == 0) == (dlogits_t == nullptr || n == 0) || m == 0 || n == 0,
                "flat_ce: pass both gradient outputs or neither");
  LPGNN_REQUIRE((uintptr_t)workspace % 16 == 0, "flat_ce: workspace must be 16-byte aligned");
  LPGNN_REQUIRE(!focal || gamma >= 1.f, "flat_ce: focal gamma=%f must be >= 1", gamma);
  if (workspace_bytes < lpgnn_flat_ce_workspace_bytes(m, n)) { set_error("flat_ce: workspace too small"); return LPGNN_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  FlatWs* ws = reinterpret_cast<FlatWs*>(workspace);
  float* partials = reinterpret_cast<float*>(ws + 1);
  LPGNN_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(FlatWs), st));
  ce_flat_kernel<<<ceil_div((int64_t)m + n, kThreads), kThreads, 0, st>>>(logits_s, y_s, m, logits_t, y_t, n, focal ? 1 : 0, gamma,
                                                                       ws, partials, dlogits_s, dlogits_t, loss_out,
                                                                       grad_scale_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_basis_metrics(const void* status, int status_is_i64, const int64_t* y_s, int32_t m, const int64_t* y_t,
                                   int32_t n, int32_t* counts_out, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m >= 0 && n >= 0 && counts_out, "basis_metrics: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  LPGNN_CUDA_OK(cudaMemsetAsync(counts_out, 0, 8 * sizeof(int32_t), st));
  if ((int64_t)m + n == 0) return LPGNN_OK;
  LPGNN_REQUIRE(status && (m == 0 || y_s) && (n == 0 || y_t), "basis_metrics: null pointer");
  const int grid = min(ceil_div((int64_t)m + n, kThreads), sm_count() * 4);
  if (status_is_i64)
    basis_metrics_kernel<int64_t><<<grid, kThreads, 0, st>>>(reinterpret_cast<const int64_t*>(status), y_s, m, y_t, n, counts_out);
  else
    basis_metrics_kernel<uint8_t><<<grid, kThreads, 0, st>>>(reinterpret_cast<const uint8_t*>(status), y_s, m, y_t, n, counts_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" size_t lpgnn_balanced_ce_segmented_workspace_bytes(int32_t n_segments) {
  return sizeof(SegWs) + (size_t)2 * (n_segments > 0 ? n_segments : 1) * sizeof(float);
}

extern "C" int lpgnn_balanced_ce_segmented(const float* logits_s, const int64_t* y_s, const int32_t* cons_ptr,
                                           const float* logits_t, const int64_t* y_t, const int32_t* vars_ptr,
                                           int32_t n_segments, int merge_lu, float* loss_out, float* dlogits_s, float* dlogits_t,
                                           void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(n_segments > 0, "balanced_ce_segmented: n_segments=%d must be positive", n_segments);
  LPGNN_REQUIRE(logits_s && y_s && cons_ptr && logits_t && y_t && vars_ptr && loss_out && workspace, "balanced_ce_segmented: null pointer");
  LPGNN_REQUIRE((dlogits_s == nullptr) == (dlogits_t == nullptr), "balanced_ce_segmented: pass both gradient outputs or neither");
  LPGNN_REQUIRE((uintptr_t)workspace % 16 == 0, "balanced_ce_segmented: workspace must be 16-byte aligned");
  if (workspace_bytes < lpgnn_balanced_ce_segmented_workspace_bytes(n_segments)) {
    set_error("balanced_ce_segmented: workspace too small");
    return LPGNN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  SegWs* ws = reinterpret_cast<SegWs*>(workspace);
  LPGNN_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(SegWs), st));
  ce_segmented_kernel<<<2 * n_segments, kThreads, 0, st>>>(logits_s, y_s, cons_ptr, logits_t, y_t, vars_ptr, n_segments, merge_lu, ws,
                                                          reinterpret_cast<float*>(ws + 1), dlogits_s, dlogits_t, loss_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}
