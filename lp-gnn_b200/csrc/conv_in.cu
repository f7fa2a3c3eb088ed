// (a2+a3, input layer) Fused aggregation + node transform for narrow inputs: conv1 of GCN_FC.
//
// Replaces PyG GraphConv.forward for the (p,q)->hids layer (reference arch.py:170 built,
// arch.py:75-80 + 181-182 executed): spmm_sum over 8-wide features, two Linear layers, bias,
// add and relu_ -- 6 library kernels and two [rows,hids] round trips in the reference -- with one
// kernel that reads the 8-wide inputs and writes the hids-wide activation exactly once.
//
// Per block: kRows destination rows.
//   phase 1  z[r] = [ sum_e val[e]*Xsrc[idx[e],:]  |  Xdst[r,:] ]   (fp32, CSR order) -> smem
//   phase 2  out[r, c] = epi(b[c] + sum_k z[r][k] * Wcat[c][k]), one output column per thread,
//            weights of the column held in registers, rows looped; stores are coalesced.
// HBM-bound on the output write: (rows*N*sizeof(out)) bytes; see DESIGN.md.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 32;

template <int KT, typename OutT>
__global__ void __launch_bounds__(kThreads)
conv_in_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
               int32_t rows, const float* __restrict__ Xsrc, int k_src, const float* __restrict__ Xdst, int k_dst,
               const float* __restrict__ W_rel, const float* __restrict__ b_rel, const float* __restrict__ W_root,
               int N, OutT* __restrict__ out, int relu, float* __restrict__ agg_out) {
  __shared__ __align__(16) float z[kRows][KT];
  const int K = k_src + k_dst;
  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  const int nrows = (int)min((int64_t)kRows, rows - row0);

  // ---- phase 1: aggregate + stage x_dst.  8 adjacent lanes share a row (they read the same (idx,val) pair
  // through one broadcast transaction and 32 contiguous bytes of the source row); four neighbours in flight.
  for (int p = threadIdx.x; p < kRows * 8; p += kThreads) {
    const int r = p >> 3, f0 = p & 7;
    if (r < nrows) {
      const int64_t row = row0 + r;
      const int32_t beg = ptr[row], end = ptr[row + 1];
      for (int k = f0; k < k_src; k += 8) {
        float v = 0.f;
        int32_t e = beg;
        for (; e + 4 <= end; e += 4) {
          const int32_t i0 = __ldg(idx + e), i1 = __ldg(idx + e + 1), i2 = __ldg(idx + e + 2), i3 = __ldg(idx + e + 3);
          const float w0 = __ldg(val + e), w1 = __ldg(val + e + 1), w2 = __ldg(val + e + 2), w3 = __ldg(val + e + 3);
          const float x0 = __ldg(Xsrc + (int64_t)i0 * k_src + k), x1 = __ldg(Xsrc + (int64_t)i1 * k_src + k);
          const float x2 = __ldg(Xsrc + (int64_t)i2 * k_src + k), x3 = __ldg(Xsrc + (int64_t)i3 * k_src + k);
          v = fmaf(w0, x0, v); v = fmaf(w1, x1, v); v = fmaf(w2, x2, v); v = fmaf(w3, x3, v);   // CSR order
        }
        for (; e < end; ++e) v = fmaf(__ldg(val + e), __ldg(Xsrc + (int64_t)__ldg(idx + e) * k_src + k), v);
        if (agg_out) agg_out[row * k_src + k] = v;
        z[r][k] = v;
      }
      for (int k = f0; k < k_dst; k += 8) z[r][k_src + k] = __ldg(Xdst + row * k_dst + k);
      for (int k = K + f0; k < KT; k += 8) z[r][k] = 0.f;
    } else {
      for (int k = f0; k < KT; k += 8) z[r][k] = 0.f;
    }
  }
  __syncthreads();

  // ---- phase 2: one output column per thread, all rows of the tile
  for (int c = threadIdx.x; c < N; c += kThreads) {
    float w[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      float wk = 0.f;
      if (k < k_src) wk = __ldg(W_rel + (int64_t)c * k_src + k);
      else if (k < K) wk = __ldg(W_root + (int64_t)c * k_dst + (k - k_src));
      w[k] = wk;
    }
    const float bias = b_rel ? __ldg(b_rel + c) : 0.f;
    for (int r = 0; r < nrows; ++r) {
      float acc = bias;
#pragma unroll
      for (int k4 = 0; k4 < KT; k4 += 4) {
        const float4 zz = *reinterpret_cast<const float4*>(&z[r][k4]);
        acc = fmaf(zz.x, w[k4 + 0], acc);
        acc = fmaf(zz.y, w[k4 + 1], acc);
        acc = fmaf(zz.z, w[k4 + 2], acc);
        acc = fmaf(zz.w, w[k4 + 3], acc);
      }
      if (relu) acc = fmaxf(acc, 0.f);
      out[(row0 + r) * N + c] = static_cast<OutT>(acc);
    }
  }
}

template <int KT>
int launch(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const float* Xsrc, int k_src,
           const float* Xdst, int k_dst, const float* W_rel, const float* b_rel, const float* W_root, int N,
           void* out, int out_dtype, int relu, float* agg_out, cudaStream_t st) {
  const int grid = ceil_div(rows, kRows);
  if (out_dtype == LPGNN_F32)
    conv_in_kernel<KT, float><<<grid, kThreads, 0, st>>>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel,
                                                         W_root, N, reinterpret_cast<float*>(out), relu, agg_out);
  else
    conv_in_kernel<KT, __nv_bfloat16><<<grid, kThreads, 0, st>>>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel,
                                                                 b_rel, W_root, N,
                                                                 reinterpret_cast<__nv_bfloat16*>(out), relu, agg_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" int lpgnn_conv_in_fused(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                                   const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst,
                                   const float* W_rel, const float* b_rel, const float* W_root, int32_t N, void* out,
                                   int out_dtype, int epilogue, float* agg_out, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && N > 0, "conv_in_fused: bad shape rows=%d N=%d", rows, N);
  LPGNN_REQUIRE(k_src >= 1 && k_dst >= 0 && k_src + k_dst <= 64, "conv_in_fused: k_src+k_dst=%d must be in [1,64]",
                k_src + k_dst);
  LPGNN_REQUIRE(out_dtype == LPGNN_F32 || out_dtype == LPGNN_BF16, "conv_in_fused: bad out dtype %d", out_dtype);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && Xsrc && W_rel && out && (k_dst == 0 || (Xdst && W_root)), "conv_in_fused: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int relu = (epilogue & LPGNN_EPI_RELU) ? 1 : 0;
  const int K = k_src + k_dst;
  if (K <= 16)
    return launch<16>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, out_dtype, relu,
                      agg_out, st);
  if (K <= 32)
    return launch<32>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, out_dtype, relu,
                      agg_out, st);
  return launch<64>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, out_dtype, relu,
                    agg_out, st);
}
