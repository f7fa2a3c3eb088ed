// (f-2) LP scaling + node features on the device: a raw LP (c, b_l, A, b_u, l, u) becomes the scaled matrix
// values of both graph orientations and the 8-per-node feature rows the model reads.
//
// Replaces dataset.scaling (reference dataset.py:23-76, helpers utils.py:323-332) and dataset.cvt_to_features
// (dataset.py:79-96, helpers utils.py:335-383), which the reference runs offline in float64 on the host:
//   rows     s_row = max(unit|b_l|, unit|b_u|)           A[i,:] /= s_row,  b_l /= s_row,  b_u /= s_row
//   columns  s_col = max(unit(colmax|A|), 1/unit|l|, 1/unit|u|)   A[:,j] /= s_col,  l *= s_col,  u *= s_col,  c /= s_col
//   cost     c /= unit(max|c|)                                     (unit(x) = 1 where x is 0 or inf)
//   x_t[j] = [c_j, nnz(A[:,j])/m, cos(b_l, A[:,j]), cos(b_u, A[:,j]), l_j|0, tag(l_j), u_j|0, tag(u_j)]
//   x_s[i] = [cos(c, A[i,:]), nnz(A[i,:])/n, cos(l, A[i,:]), cos(u, A[i,:]), b_l,i|0, tag, b_u,i|0, tag]
//   cos(v, a) = <clip(v, +-1e8), a> / (max'(|clip(v)|) * max'(|a|)),  max'(0) = 1e-6
// All arithmetic is float64 with separately rounded multiplies and adds (no FMA contraction) and the per-row /
// per-column sums run in CSR / CSC order, which is the order scipy uses -- so the scaled values and the dot
// products are bit-identical to the reference; only the vector norms (numpy's pairwise summation) differ in the
// last bits.  One thread per row / column: LP rows are short (~10 entries).  Deterministic: the only atomic is an
// integer max.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;
constexpr int kRedBlocks = 128;
constexpr double kBig = 1e308, kClip = 1e8, kTiny = 1e-6;

__device__ __forceinline__ double unit_deg(double v) {
  const double s = fabs(v);
  return (isinf(s) || s == 0.0) ? 1.0 : s;
}
__device__ __forceinline__ double clipv(double v) { return fmin(fmax(v, -kClip), kClip); }

struct Scal {                 // device scalars
  unsigned long long cmax_bits;   // max |c / s_col| as the bit pattern of a non-negative double (integer max is exact)
  double part[5][kRedBlocks];     // per-block partial sums of squares: b_l, b_u, l, u, c (after scaling, clipped)
  double norm[5];                 // sqrt of the totals (0 -> 1e-6)
  double s_c;
};

// rows: row scale, scaled bounds, row-scaled matrix values (CSR order)
__global__ void __launch_bounds__(kThreads)
rows_scale_kernel(const int32_t* __restrict__ rowptr, const double* __restrict__ a, const double* __restrict__ b_l,
                  const double* __restrict__ b_u, int m, double* __restrict__ a1, double* __restrict__ bl_o,
                  double* __restrict__ bu_o) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= m) return;
  double bl = b_l[i], bu = b_u[i];
  if (bu > kBig) bu = INFINITY;
  if (bl < -kBig) bl = -INFINITY;
  const double s = fmax(unit_deg(bl), unit_deg(bu));
  bl_o[i] = bl / s;
  bu_o[i] = bu / s;
  for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) a1[e] = a[e] / s;
}

// columns: column scale, scaled l / u / c, max |c|
__global__ void __launch_bounds__(kThreads)
cols_scale_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ csr2csc, const double* __restrict__ a1,
                  const double* __restrict__ c, const double* __restrict__ l, const double* __restrict__ u, int n,
                  double* __restrict__ s_col, double* __restrict__ c1, double* __restrict__ l_o, double* __restrict__ u_o,
                  Scal* sc) {
  const int j = blockIdx.x * kThreads + threadIdx.x;
  double cabs = 0.0;
  if (j < n) {
    double lj = l[j], uj = u[j];
    if (uj > kBig) uj = INFINITY;
    if (lj < -kBig) lj = -INFINITY;
    double mx = 0.0;
    for (int k = colptr[j]; k < colptr[j + 1]; ++k) mx = fmax(mx, fabs(a1[csr2csc[k]]));
    if (isinf(mx) || mx == 0.0) mx = 1.0;
    const double s = fmax(mx, fmax(1.0 / unit_deg(lj), 1.0 / unit_deg(uj)));
    s_col[j] = s;
    l_o[j] = lj * s;
    u_o[j] = uj * s;
    const double cj = c[j] / s;
    c1[j] = cj;
    cabs = fabs(cj);
  }
  // block max -> integer atomic max on the bit pattern (non-negative doubles order like their bits); NaN ignored
  unsigned long long bits = (cabs == cabs) ? (unsigned long long)__double_as_longlong(cabs) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
    bits = other > bits ? other : bits;
  }
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(&sc->cmax_bits, bits);
}

// entries: fully scaled values, float64 (CSR order) and float32 for both orientations of the graph
__global__ void __launch_bounds__(kThreads)
entries_scale_kernel(const int32_t* __restrict__ col, const double* __restrict__ a1, const double* __restrict__ s_col,
                     int64_t z, double* __restrict__ a2, float* __restrict__ val) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= z) return;
  const double v = a1[e] / s_col[col[e]];
  a2[e] = v;
  val[e] = (float)v;
}
__global__ void __launch_bounds__(kThreads)
entries_csc_kernel(const int32_t* __restrict__ csr2csc, const double* __restrict__ a2, int64_t z, float* __restrict__ val_csc) {
  const int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (k < z) val_csc[k] = (float)a2[csr2csc[k]];
}

// c /= s_c in place, and the five sums of squares (fixed-order partials)
__global__ void __launch_bounds__(kThreads)
norms_partial_kernel(double* __restrict__ c1, const double* __restrict__ bl, const double* __restrict__ bu,
                     const double* __restrict__ l, const double* __restrict__ u, int m, int n, Scal* sc) {
  double s_c = __longlong_as_double((long long)sc->cmax_bits);
  if (s_c == 0.0) s_c = 1.0;
  double acc[5] = {0, 0, 0, 0, 0};
  const int per_m = (m + kRedBlocks * kThreads - 1) / (kRedBlocks * kThreads);
  const int per_n = (n + kRedBlocks * kThreads - 1) / (kRedBlocks * kThreads);
  const int t = blockIdx.x * kThreads + threadIdx.x;
  for (int k = 0; k < per_m; ++k) {                  // contiguous slice per thread: a fixed summation order
    const int i = t * per_m + k;
    if (i < m) {
      const double x = clipv(bl[i]), y = clipv(bu[i]);
      acc[0] = __dadd_rn(acc[0], __dmul_rn(x, x));
      acc[1] = __dadd_rn(acc[1], __dmul_rn(y, y));
    }
  }
  for (int k = 0; k < per_n; ++k) {
    const int j = t * per_n + k;
    if (j < n) {
      const double x = clipv(l[j]), y = clipv(u[j]);
      const double cj = c1[j] / s_c;
      c1[j] = cj;
      const double w = clipv(cj);
      acc[2] = __dadd_rn(acc[2], __dmul_rn(x, x));
      acc[3] = __dadd_rn(acc[3], __dmul_rn(y, y));
      acc[4] = __dadd_rn(acc[4], __dmul_rn(w, w));
    }
  }
  __shared__ double red[5][kThreads];
#pragma unroll
  for (int q = 0; q < 5; ++q) red[q][threadIdx.x] = acc[q];
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    for (int k = 0; k < kThreads; ++k) s += red[threadIdx.x][k];
    sc->part[threadIdx.x][blockIdx.x] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) sc->s_c = s_c;
}

__global__ void norms_final_kernel(Scal* sc) {
  const int q = threadIdx.x;
  if (q >= 5) return;
  double s = 0.0;
  for (int k = 0; k < kRedBlocks; ++k) s += sc->part[q][k];
  s = sqrt(s);
  sc->norm[q] = s == 0.0 ? kTiny : s;
}

__device__ __forceinline__ void value_tag(double v, float* out) {
  out[0] = isinf(v) ? 0.f : (float)v;
  out[1] = v == INFINITY ? 1.f : (v == -INFINITY ? -1.f : 0.f);
}

// variables: x_t rows.  Column sums in CSC order (ascending row), products and sums rounded separately.
__global__ void __launch_bounds__(kThreads)
var_features_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row_csc,
                    const int32_t* __restrict__ csr2csc, const double* __restrict__ a2, const double* __restrict__ c,
                    const double* __restrict__ bl, const double* __restrict__ bu, const double* __restrict__ l,
                    const double* __restrict__ u, int m, int n, const Scal* __restrict__ sc, float* __restrict__ x_t) {
  const int j = blockIdx.x * kThreads + threadIdx.x;
  if (j >= n) return;
  double sq = 0.0, dl = 0.0, du = 0.0;
  int cnt = 0;
  for (int k = colptr[j]; k < colptr[j + 1]; ++k) {
    const double a = a2[csr2csc[k]];
    const int i = row_csc[k];
    sq = __dadd_rn(sq, __dmul_rn(a, a));
    dl = __dadd_rn(dl, __dmul_rn(clipv(bl[i]), a));
    du = __dadd_rn(du, __dmul_rn(clipv(bu[i]), a));
    cnt += a != 0.0;
  }
  double nc = sqrt(sq);
  if (nc == 0.0) nc = kTiny;
  float* o = x_t + (int64_t)j * 8;
  o[0] = (float)c[j];
  o[1] = (float)((double)cnt / (double)m);
  o[2] = (float)(dl / __dmul_rn(sc->norm[0], nc));
  o[3] = (float)(du / __dmul_rn(sc->norm[1], nc));
  value_tag(l[j], o + 4);
  value_tag(u[j], o + 6);
}

// constraints: x_s rows.  Row sums in CSR order (ascending column).
__global__ void __launch_bounds__(kThreads)
con_features_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ a2,
                    const double* __restrict__ c, const double* __restrict__ bl, const double* __restrict__ bu,
                    const double* __restrict__ l, const double* __restrict__ u, int m, int n,
                    const Scal* __restrict__ sc, float* __restrict__ x_s) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= m) return;
  double sq = 0.0, dc = 0.0, dl = 0.0, du = 0.0;
  int cnt = 0;
  for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
    const double a = a2[e];
    const int j = col[e];
    sq = __dadd_rn(sq, __dmul_rn(a, a));
    dc = __dadd_rn(dc, __dmul_rn(clipv(c[j]), a));
    dl = __dadd_rn(dl, __dmul_rn(clipv(l[j]), a));
    du = __dadd_rn(du, __dmul_rn(clipv(u[j]), a));
    cnt += a != 0.0;
  }
  double nr = sqrt(sq);
  if (nr == 0.0) nr = kTiny;
  float* o = x_s + (int64_t)i * 8;
  o[0] = (float)(dc / __dmul_rn(sc->norm[4], nr));
  o[1] = (float)((double)cnt / (double)n);
  o[2] = (float)(dl / __dmul_rn(sc->norm[2], nr));
  o[3] = (float)(du / __dmul_rn(sc->norm[3], nr));
  value_tag(bl[i], o + 4);
  value_tag(bu[i], o + 6);
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" size_t lpgnn_lp_features_workspace_bytes(int64_t nnz, int32_t m, int32_t n) {
  size_t b = align_up(sizeof(Scal), 256);
  b += align_up((size_t)nnz * 8, 256);           // row-scaled values
  b += align_up((size_t)n * 8, 256);             // s_col
  return b;
}

extern "C" int lpgnn_lp_features(const int32_t* rowptr, const int32_t* col, const int32_t* colptr, const int32_t* row_csc,
                                 const int32_t* csr2csc, const double* a_csr, const double* c, const double* b_l,
                                 const double* b_u, const double* l, const double* u, int64_t nnz, int32_t m, int32_t n,
                                 float* val, float* val_csc, float* x_s, float* x_t, double* a_scaled, double* c_out,
                                 double* bl_out, double* bu_out, double* l_out, double* u_out, void* workspace,
                                 size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m > 0 && n > 0 && nnz >= 0, "lp_features: bad shape m=%d n=%d nnz=%lld", m, n, (long long)nnz);
  LPGNN_REQUIRE(rowptr && colptr && c && b_l && b_u && l && u && val && val_csc && x_s && x_t && a_scaled && c_out && bl_out &&
                    bu_out && l_out && u_out && workspace && (nnz == 0 || (col && row_csc && csr2csc && a_csr)),
                "lp_features: null pointer");
  LPGNN_REQUIRE((uintptr_t)workspace % 256 == 0, "lp_features: workspace must be 256-byte aligned");
  if (workspace_bytes < lpgnn_lp_features_workspace_bytes(nnz, m, n)) {
    set_error("lp_features: workspace too small");
    return LPGNN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* p = reinterpret_cast<char*>(workspace);
  Scal* sc = reinterpret_cast<Scal*>(p);            p += align_up(sizeof(Scal), 256);
  double* a1 = reinterpret_cast<double*>(p);        p += align_up((size_t)nnz * 8, 256);
  double* s_col = reinterpret_cast<double*>(p);
  LPGNN_CUDA_OK(cudaMemsetAsync(sc, 0, sizeof(unsigned long long), st));
  rows_scale_kernel<<<ceil_div(m, kThreads), kThreads, 0, st>>>(rowptr, a_csr, b_l, b_u, m, a1, bl_out, bu_out);
  cols_scale_kernel<<<ceil_div(n, kThreads), kThreads, 0, st>>>(colptr, csr2csc, a1, c, l, u, n, s_col, c_out, l_out, u_out, sc);
  int launches = 2;
  if (nnz > 0) {
    entries_scale_kernel<<<ceil_div(nnz, kThreads), kThreads, 0, st>>>(col, a1, s_col, nnz, a_scaled, val);
    entries_csc_kernel<<<ceil_div(nnz, kThreads), kThreads, 0, st>>>(csr2csc, a_scaled, nnz, val_csc);
    launches += 2;
  }
  norms_partial_kernel<<<kRedBlocks, kThreads, 0, st>>>(c_out, bl_out, bu_out, l_out, u_out, m, n, sc);
  norms_final_kernel<<<1, 32, 0, st>>>(sc);
  var_features_kernel<<<ceil_div(n, kThreads), kThreads, 0, st>>>(colptr, row_csc, csr2csc, a_scaled, c_out, bl_out, bu_out,
                                                                 l_out, u_out, m, n, sc, x_t);
  con_features_kernel<<<ceil_div(m, kThreads), kThreads, 0, st>>>(rowptr, col, a_scaled, c_out, bl_out, bu_out, l_out, u_out,
                                                                 m, n, sc, x_s);
  LPGNN_LAUNCH_OK();
  count_launches(launches + 4);
  return LPGNN_OK;
}
