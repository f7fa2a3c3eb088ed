// (a2+a3, input layer) Fused aggregation + node transform for narrow inputs: conv1 of GCN_FC.
//
// Replaces PyG GraphConv.forward for the (p,q)->hids layer (reference arch.py:170 built,
// arch.py:75-80 + 181-182 executed): spmm_sum over 8-wide features, two Linear layers, bias,
// add and relu_ -- 6 library kernels and two [rows,hids] round trips in the reference -- with one
// kernel that reads the 8-wide inputs and writes the hids-wide activation exactly once.
//
// Two launches behind one entry point (a single block doing both phases is latency-bound on the
// dependent gather chain):
//   gather_cat          z[r] = [ sum_e val[e]*Xsrc[idx[e],:] | Xdst[r,:] ]  (fp32, CSR order), 64 B per row;
//                       z is also what the weight gradient of the layer needs, so it is an OUTPUT
//   small_k_transform   out[r, c] = epi(b[c] + sum_k z[r][k] * Wcat[c][k]), streaming, packed-FMA
// HBM-bound on the output write: rows*N*sizeof(out) bytes; see DESIGN.md.
#include "common.cuh"

#include <type_traits>

namespace lpgnn {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 64;            // rows per transform block
constexpr int kColsPerBlock = 2 * kThreads;

// ---- kernel 1: z[row] = [ sum_e val[e]*Xsrc[idx[e],:] | Xdst[row,:] | 0-pad ]   (fp32, CSR order)
// 8 adjacent lanes share a row: they read the same (idx,val) pair through one broadcast transaction and
// 32 contiguous bytes of the source row; four neighbours are in flight per lane.
template <int KT, typename ZT>
__global__ void __launch_bounds__(kThreads)
gather_cat_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                  int32_t rows, const float* __restrict__ Xsrc, int k_src, const float* __restrict__ Xdst, int k_dst,
                  float* __restrict__ z, ZT* __restrict__ zb /*[rows,64] bf16 / half, or null*/) {
  const int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t row = p >> 3;
  const int f0 = (int)(p & 7);
  if (row >= rows) return;
  const int K = k_src + k_dst;
  const int32_t beg = ptr[row], end = ptr[row + 1];
  float* zr = z ? z + row * KT : nullptr;
  ZT* zbr = zb ? zb + row * 64 : nullptr;
  auto pk = [](float a, float b) { return Half16<ZT>::pack(a, b); };
  auto cv = [](float a) -> ZT { if constexpr (sizeof(ZT) == 2 && std::is_same<ZT, __half>::value) return __float2half_rn(a); else return __float2bfloat16_rn(a); };
  if (k_src == 8 && k_dst == 8) {
    // The reference's shape (GCN_FC(8, 8, ...)): lane f0 aggregates feature f0, the 8 lanes of the row exchange their
    // sums, and every lane then writes ONE 16-byte chunk of the row -- a single coalesced store per output instead of
    // 8 scalar ones per lane.
    float v = 0.f;
    int32_t e = beg;
    for (; e + 4 <= end; e += 4) {
      const int32_t i0 = __ldg(idx + e), i1 = __ldg(idx + e + 1), i2 = __ldg(idx + e + 2), i3 = __ldg(idx + e + 3);
      const float w0 = __ldg(val + e), w1 = __ldg(val + e + 1), w2 = __ldg(val + e + 2), w3 = __ldg(val + e + 3);
      const float x0 = __ldg(Xsrc + (int64_t)i0 * 8 + f0), x1 = __ldg(Xsrc + (int64_t)i1 * 8 + f0);
      const float x2 = __ldg(Xsrc + (int64_t)i2 * 8 + f0), x3 = __ldg(Xsrc + (int64_t)i3 * 8 + f0);
      v = fmaf(w0, x0, v); v = fmaf(w1, x1, v); v = fmaf(w2, x2, v); v = fmaf(w3, x3, v);   // CSR order
    }
    for (; e < end; ++e) v = fmaf(__ldg(val + e), __ldg(Xsrc + (int64_t)__ldg(idx + e) * 8 + f0), v);
    const int gbase = (threadIdx.x & 31) & ~7;
    const uint32_t gmask = 0xffu << gbase;
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __shfl_sync(gmask, v, gbase + k);
    float4 xa, xb;                        // the node's own 8 features (vector loads only when the caller's array allows)
    if ((reinterpret_cast<uintptr_t>(Xdst) & 15) == 0) {
      const float4* xd = reinterpret_cast<const float4*>(Xdst + row * 8);
      xa = __ldg(xd); xb = __ldg(xd + 1);
    } else {
      const float* xd = Xdst + row * 8;
      xa = make_float4(__ldg(xd), __ldg(xd + 1), __ldg(xd + 2), __ldg(xd + 3));
      xb = make_float4(__ldg(xd + 4), __ldg(xd + 5), __ldg(xd + 6), __ldg(xd + 7));
    }
    if (zr && f0 < KT / 4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f0 == 0) o = make_float4(a[0], a[1], a[2], a[3]);
      else if (f0 == 1) o = make_float4(a[4], a[5], a[6], a[7]);
      else if (f0 == 2) o = xa;
      else if (f0 == 3) o = xb;
      reinterpret_cast<float4*>(zr)[f0] = o;
    }
    if (KT / 4 > 8 && zr) for (int c = 8 + f0; c < KT / 4; c += 8) reinterpret_cast<float4*>(zr)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (zbr) {
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (f0 == 0) o = make_uint4(pk(a[0], a[1]), pk(a[2], a[3]), pk(a[4], a[5]), pk(a[6], a[7]));
      else if (f0 == 1) o = make_uint4(pk(xa.x, xa.y), pk(xa.z, xa.w), pk(xb.x, xb.y), pk(xb.z, xb.w));
      else if (f0 == 2) o.x = pk(1.f, 0.f);   // column K = 16 carries 1.0 (bias gradient through lpgnn_wgrad)
      reinterpret_cast<uint4*>(zbr)[f0] = o;
    }
    return;
  }
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
    if (zr) zr[k] = v;
    if (zbr) zbr[k] = cv(v);
  }
  for (int k = f0; k < k_dst; k += 8) {
    const float v = __ldg(Xdst + row * k_dst + k);
    if (zr) zr[k_src + k] = v;
    if (zbr) zbr[k_src + k] = cv(v);
  }
  if (zr) for (int k = K + f0; k < KT; k += 8) zr[k] = 0.f;
  // column K carries 1.0: the forward weight column there is zero, and dPre^T z_bf16 (lpgnn_wgrad) gets the bias gradient
  if (zbr) for (int k = K + f0; k < 64; k += 8) zbr[k] = cv(k == K ? 1.f : 0.f);
}

// ---- kernel 2: out[r, c] = epi(b[c] + sum_k z[r][k] * Wcat[c][k]).  Block = 64 rows x 512 columns; a thread
// owns two adjacent columns (weights in registers as float2, packed FFMA2) and walks the rows; the z tile is a
// broadcast shared-memory read.  Stores are coalesced (4 / 8 bytes per lane, 128 / 256 bytes per warp).
template <int KT, typename OutT>
__global__ void __launch_bounds__(kThreads, KT <= 32 ? 2 : 1)   // KT = 64: 128 weight registers per thread
small_k_transform_kernel(const float* __restrict__ z, int32_t rows, const float* __restrict__ W_rel, int k_src,
                         const float* __restrict__ W_root, int k_dst, const float* __restrict__ b_rel, int N,
                         OutT* __restrict__ out, int relu) {
  __shared__ __align__(16) float zs[kRows][KT];
  const int K = k_src + k_dst;
  const int64_t row0 = (int64_t)blockIdx.x * kRows;      // row blocks on grid.x (2^31 - 1 blocks), column blocks on grid.y
  const int nrows = (int)min((int64_t)kRows, rows - row0);
  {
    const float4* src = reinterpret_cast<const float4*>(z + row0 * KT);
    float4* dst = reinterpret_cast<float4*>(&zs[0][0]);
    for (int i = threadIdx.x; i < nrows * KT / 4; i += kThreads) dst[i] = __ldg(src + i);
  }
  const int c = blockIdx.y * kColsPerBlock + 2 * threadIdx.x;
  float2 w[KT];
  float2 bias = make_float2(0.f, 0.f);
  if (c < N) {
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      float a = 0.f, b = 0.f;
      if (k < k_src) { a = __ldg(W_rel + (int64_t)c * k_src + k); b = __ldg(W_rel + (int64_t)(c + 1) * k_src + k); }
      else if (k < K) { a = __ldg(W_root + (int64_t)c * k_dst + (k - k_src)); b = __ldg(W_root + (int64_t)(c + 1) * k_dst + (k - k_src)); }
      w[k] = make_float2(a, b);
    }
    if (b_rel) bias = make_float2(__ldg(b_rel + c), __ldg(b_rel + c + 1));
  }
  __syncthreads();
  if (c >= N) return;
#pragma unroll 2
  for (int r = 0; r < nrows; ++r) {
    float2 acc = bias;
#pragma unroll
    for (int k4 = 0; k4 < KT; k4 += 4) {
      const float4 zz = *reinterpret_cast<const float4*>(&zs[r][k4]);
      acc = __ffma2_rn(make_float2(zz.x, zz.x), w[k4 + 0], acc);
      acc = __ffma2_rn(make_float2(zz.y, zz.y), w[k4 + 1], acc);
      acc = __ffma2_rn(make_float2(zz.z, zz.z), w[k4 + 2], acc);
      acc = __ffma2_rn(make_float2(zz.w, zz.w), w[k4 + 3], acc);
    }
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); }
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float2*>(out + (row0 + r) * N + c) = acc;
    } else {
      *reinterpret_cast<uint32_t*>(out + (row0 + r) * N + c) = pack_bf16(acc.x, acc.y);
    }
  }
}

// ---- fp32 input layer that ALSO emits its output in x2 form (operand of the fp32 tensor-core transform, gemm_x2.cu).
// The row scale cannot wait for the row maximum (a row's columns are spread over blocks), and it does not have to: any
// power of two s with |x[r,:]| <= s * 2^13 keeps the 22 significant bits (IEEE half has 2^-24 of absolute resolution
// below that, i.e. 2^-37 of the bound), so an A-PRIORI bound serves:  |x[r,c]| <= sum_k |z[r,k]| * max_c |W[c,k]| + max_c |b|.
// wabs_kernel: wabs[k] = max over output features of |Wcat[c][k]| (k < K), wabs[KT] = max |b|.
template <int KT>
__global__ void __launch_bounds__(kThreads)
wabs_kernel(const float* __restrict__ W_rel, int k_src, const float* __restrict__ W_root, int k_dst, const float* __restrict__ b_rel,
            int N, float* __restrict__ wabs) {
  __shared__ float red[kThreads / 32][KT + 1];
  float mx[KT + 1];
#pragma unroll
  for (int k = 0; k <= KT; ++k) mx[k] = 0.f;
  for (int c = threadIdx.x; c < N; c += kThreads) {
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      float w = 0.f;
      if (k < k_src) w = __ldg(W_rel + (int64_t)c * k_src + k);
      else if (k < k_src + k_dst) w = __ldg(W_root + (int64_t)c * k_dst + (k - k_src));
      mx[k] = fmaxf(mx[k], fabsf(w));
    }
    if (b_rel) mx[KT] = fmaxf(mx[KT], fabsf(__ldg(b_rel + c)));
  }
#pragma unroll
  for (int k = 0; k <= KT; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = mx[k];
  }
  __syncthreads();
  if (threadIdx.x <= KT) {
    float v = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) v = fmaxf(v, red[w][threadIdx.x]);
    wabs[threadIdx.x] = v;
  }
}

// power-of-two scale s >= bound / 2^12 (one binade of margin over |x| <= s * 2^13 for the fp32 rounding of the bound)
__device__ __forceinline__ float x2_scale_for_bound(float bound) {
  int e = 12;
  if (bound > 0.f && bound < __int_as_float(0x7f800000)) e = (int)((__float_as_uint(bound) >> 23) & 0xffu) - 127 + 1;
  e = max(-100, min(e, 112));
  return __int_as_float((uint32_t)(127 + e - 12) << 23);      // 2^(e-12): bound < 2^e  =>  |x| / s < 2^12
}

template <int KT>
__global__ void __launch_bounds__(kThreads, KT <= 32 ? 2 : 1)
small_k_transform_x2_kernel(const float* __restrict__ z, int32_t rows, const float* __restrict__ W_rel, int k_src,
                            const float* __restrict__ W_root, int k_dst, const float* __restrict__ b_rel, int N,
                            float* __restrict__ out, int relu, const float* __restrict__ wabs, __half* __restrict__ hi,
                            __half* __restrict__ lo, float* __restrict__ scale) {
  __shared__ __align__(16) float zs[kRows][KT];
  __shared__ float down_s[kRows];
  const int K = k_src + k_dst;
  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  const int nrows = (int)min((int64_t)kRows, rows - row0);
  {
    const float4* src = reinterpret_cast<const float4*>(z + row0 * KT);
    float4* dst = reinterpret_cast<float4*>(&zs[0][0]);
    for (int i = threadIdx.x; i < nrows * KT / 4; i += kThreads) dst[i] = __ldg(src + i);
  }
  const int c = blockIdx.y * kColsPerBlock + 2 * threadIdx.x;
  float2 w[KT];
  float2 bias = make_float2(0.f, 0.f);
  if (c < N) {
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      float a = 0.f, b = 0.f;
      if (k < k_src) { a = __ldg(W_rel + (int64_t)c * k_src + k); b = __ldg(W_rel + (int64_t)(c + 1) * k_src + k); }
      else if (k < K) { a = __ldg(W_root + (int64_t)c * k_dst + (k - k_src)); b = __ldg(W_root + (int64_t)(c + 1) * k_dst + (k - k_src)); }
      w[k] = make_float2(a, b);
    }
    if (b_rel) bias = make_float2(__ldg(b_rel + c), __ldg(b_rel + c + 1));
  }
  __syncthreads();
  if (threadIdx.x < nrows) {                 // the row's scale from the a-priori bound (every column block derives the same)
    float bound = __ldg(wabs + KT);
#pragma unroll
    for (int k = 0; k < KT; ++k) bound = fmaf(fabsf(zs[threadIdx.x][k]), __ldg(wabs + k), bound);
    const float s = x2_scale_for_bound(bound);
    down_s[threadIdx.x] = 1.f / s;             // exact: a power of two
    if (blockIdx.y == 0) scale[row0 + threadIdx.x] = s;
  }
  __syncthreads();
  if (c >= N) return;
#pragma unroll 2
  for (int r = 0; r < nrows; ++r) {
    float2 acc = bias;
#pragma unroll
    for (int k4 = 0; k4 < KT; k4 += 4) {
      const float4 zz = *reinterpret_cast<const float4*>(&zs[r][k4]);
      acc = __ffma2_rn(make_float2(zz.x, zz.x), w[k4 + 0], acc);
      acc = __ffma2_rn(make_float2(zz.y, zz.y), w[k4 + 1], acc);
      acc = __ffma2_rn(make_float2(zz.z, zz.z), w[k4 + 2], acc);
      acc = __ffma2_rn(make_float2(zz.w, zz.w), w[k4 + 3], acc);
    }
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); }
    const int64_t o = (row0 + r) * N + c;
    *reinterpret_cast<float2*>(out + o) = acc;
    const float d = down_s[r];
    const float sx = acc.x * d, sy = acc.y * d;
    const __half hx = __float2half_rn(sx), hy = __float2half_rn(sy);
    *reinterpret_cast<__half2*>(hi + o) = __halves2half2(hx, hy);
    *reinterpret_cast<__half2*>(lo + o) = __halves2half2(__float2half_rn((sx - __half2float(hx)) * 2048.f),
                                                        __float2half_rn((sy - __half2float(hy)) * 2048.f));
  }
}

template <int KT>
int launch_x2out(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const float* Xsrc, int k_src,
                 const float* Xdst, int k_dst, const float* W_rel, const float* b_rel, const float* W_root, int N, float* out,
                 int relu, float* z, void* hi, void* lo, float* scale, float* wabs, cudaStream_t st) {
  LPGNN_REQUIRE(ceil_div(N, kColsPerBlock) <= 65535, "conv_in_fused_x2: N=%d exceeds the %d columns one launch covers", N,
                65535 * kColsPerBlock);
  wabs_kernel<KT><<<1, kThreads, 0, st>>>(W_rel, k_src, W_root, k_dst, b_rel, N, wabs);
  gather_cat_kernel<KT, __nv_bfloat16><<<ceil_div((int64_t)rows * 8, kThreads), kThreads, 0, st>>>(
      ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, z, nullptr);
  dim3 grid(ceil_div(rows, kRows), ceil_div(N, kColsPerBlock));
  small_k_transform_x2_kernel<KT><<<grid, kThreads, 0, st>>>(z, rows, W_rel, k_src, W_root, k_dst, b_rel, N, out, relu, wabs,
                                                           reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), scale);
  LPGNN_LAUNCH_OK();
  count_launches(3);
  return LPGNN_OK;
}

template <int KT>
int launch(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const float* Xsrc, int k_src,
           const float* Xdst, int k_dst, const float* W_rel, const float* b_rel, const float* W_root, int N,
           void* out, int out_dtype, int relu, float* z, cudaStream_t st) {
  LPGNN_REQUIRE(ceil_div(N, kColsPerBlock) <= 65535, "conv_in_fused: N=%d exceeds the %d columns one launch covers", N,
                65535 * kColsPerBlock);
  gather_cat_kernel<KT, __nv_bfloat16><<<ceil_div((int64_t)rows * 8, kThreads), kThreads, 0, st>>>(
      ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, z, nullptr);
  dim3 grid(ceil_div(rows, kRows), ceil_div(N, kColsPerBlock));
  if (out_dtype == LPGNN_F32)
    small_k_transform_kernel<KT, float><<<grid, kThreads, 0, st>>>(z, rows, W_rel, k_src, W_root, k_dst, b_rel, N,
                                                                   reinterpret_cast<float*>(out), relu);
  else
    small_k_transform_kernel<KT, __nv_bfloat16><<<grid, kThreads, 0, st>>>(
        z, rows, W_rel, k_src, W_root, k_dst, b_rel, N, reinterpret_cast<__nv_bfloat16*>(out), relu);
  LPGNN_LAUNCH_OK();
  count_launches(2);
  return LPGNN_OK;
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" int32_t lpgnn_conv_in_zcat_width(int32_t k_src, int32_t k_dst) {
  const int K = k_src + k_dst;
  return K <= 16 ? 16 : (K <= 32 ? 32 : 64);
}

extern "C" int lpgnn_gather_cat_ex(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                                   const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst, float* z_cat,
                                   void* z16, int z16_dtype, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && k_src >= 1 && k_dst >= 0 && k_src + k_dst <= 64, "gather_cat: bad shape");
  LPGNN_REQUIRE(!z16 || is_16bit(z16_dtype), "gather_cat: z16_dtype %d is not a 16-bit type", z16_dtype);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && Xsrc && (z_cat || z16) && (k_dst == 0 || Xdst), "gather_cat: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int KT = lpgnn_conv_in_zcat_width(k_src, k_dst);
  const int grid = ceil_div((int64_t)rows * 8, kThreads);
#define LPGNN_GC(KTV, ZT) gather_cat_kernel<KTV, ZT><<<grid, kThreads, 0, st>>>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, z_cat, reinterpret_cast<ZT*>(z16))
  if (z16 && z16_dtype == LPGNN_F16) {
    if (KT == 16) LPGNN_GC(16, __half); else if (KT == 32) LPGNN_GC(32, __half); else LPGNN_GC(64, __half);
  } else {
    if (KT == 16) LPGNN_GC(16, __nv_bfloat16); else if (KT == 32) LPGNN_GC(32, __nv_bfloat16); else LPGNN_GC(64, __nv_bfloat16);
  }
#undef LPGNN_GC
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_gather_cat(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                                const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst, float* z_cat,
                                void* z_bf16, lpgnn_stream_t stream) {
  return lpgnn_gather_cat_ex(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, z_cat, z_bf16, LPGNN_BF16, stream);
}

extern "C" int lpgnn_conv_in_fused(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                                   const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst,
                                   const float* W_rel, const float* b_rel, const float* W_root, int32_t N, void* out,
                                   int out_dtype, int epilogue, float* z_cat, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && N > 0 && N % 2 == 0, "conv_in_fused: bad shape rows=%d N=%d (N must be even)", rows, N);
  LPGNN_REQUIRE(k_src >= 1 && k_dst >= 0 && k_src + k_dst <= 64, "conv_in_fused: k_src+k_dst=%d must be in [1,64]",
                k_src + k_dst);
  LPGNN_REQUIRE(out_dtype == LPGNN_F32 || out_dtype == LPGNN_BF16, "conv_in_fused: bad out dtype %d", out_dtype);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && Xsrc && W_rel && out && z_cat && (k_dst == 0 || (Xdst && W_root)), "conv_in_fused: null pointer");
  LPGNN_REQUIRE((uintptr_t)z_cat % 16 == 0, "conv_in_fused: z_cat must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int relu = (epilogue & LPGNN_EPI_RELU) ? 1 : 0;
  const int KT = lpgnn_conv_in_zcat_width(k_src, k_dst);
  if (KT == 16)
    return launch<16>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, out_dtype, relu,
                      z_cat, st);
  if (KT == 32)
    return launch<32>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, out_dtype, relu,
                      z_cat, st);
  return launch<64>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, out_dtype, relu,
                    z_cat, st);
}

// conv_in_fused (fp32) that also writes its output as x2 operands: hi / lo IEEE-half [rows,N] and a power-of-two scale per row
// with out[r,c] = scale[r] * (hi + 2^-11 lo) to 22 bits; scale comes from an a-priori bound (|out[r,:]| <= scale[r] * 2^12),
// which is also what lpgnn_spmm_x2 needs to bound ITS rows.  wabs: float scratch [65].
extern "C" int lpgnn_conv_in_fused_x2(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows,
                                      const float* Xsrc, int32_t k_src, const float* Xdst, int32_t k_dst, const float* W_rel,
                                      const float* b_rel, const float* W_root, int32_t N, float* out, int epilogue, float* z_cat,
                                      void* hi, void* lo, float* scale, float* wabs, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && N > 0 && N % 2 == 0, "conv_in_fused_x2: bad shape rows=%d N=%d (N must be even)", rows, N);
  LPGNN_REQUIRE(k_src >= 1 && k_dst >= 0 && k_src + k_dst <= 64, "conv_in_fused_x2: k_src+k_dst=%d must be in [1,64]", k_src + k_dst);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && Xsrc && W_rel && out && z_cat && hi && lo && scale && wabs && (k_dst == 0 || (Xdst && W_root)),
                "conv_in_fused_x2: null pointer");
  LPGNN_REQUIRE((uintptr_t)z_cat % 16 == 0 && (uintptr_t)out % 8 == 0 && (uintptr_t)hi % 4 == 0 && (uintptr_t)lo % 4 == 0,
                "conv_in_fused_x2: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int relu = (epilogue & LPGNN_EPI_RELU) ? 1 : 0;
  const int KT = lpgnn_conv_in_zcat_width(k_src, k_dst);
  if (KT == 16) return launch_x2out<16>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, relu, z_cat, hi, lo, scale, wabs, st);
  if (KT == 32) return launch_x2out<32>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, relu, z_cat, hi, lo, scale, wabs, st);
  return launch_x2out<64>(ptr, idx, val, rows, Xsrc, k_src, Xdst, k_dst, W_rel, b_rel, W_root, N, out, relu, z_cat, hi, lo, scale, wabs, st);
}
