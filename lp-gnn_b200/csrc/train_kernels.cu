// Backward-pass kernels of the GCN_FC hot path (training step, reference train.py:121-129 runs
// loss.backward() through PyG/torch_sparse/ATen autograd; here every activation-sized operation of the
// backward pass is a kernel of this library):
//
//   lpgnn_head_mask_bwd  d(add_knowledge o Linear(H,3)) wrt the hidden activation, fused with the ReLU
//                        (and inverted-dropout) mask of that activation        (arch.py:186-191, 129-141)
//   lpgnn_relu_bwd       dPre = (a [+ b]) * scale * (act > 0)                    (arch.py:182, 186-188)
//   lpgnn_dropout        in-place inverted dropout with a counter-based hash   (arch.py:186-187)
//   lpgnn_transpose      [M,N] -> [N,M]; feeds the tensor-core weight-gradient GEMM (reduction over nodes)
//   lpgnn_colsum         bias gradients: column sums, two-stage, fixed order (deterministic)
//   lpgnn_small_wgrad    dW[N,K] = dY^T Z for narrow Z (input layer K=16, head K=3), two-stage, deterministic
//
// The data-gradient GEMMs and the aggregation backward reuse lpgnn_node_transform (with transposed
// weights) and lpgnn_spmm (other orientation).  No atomics on floating-point data anywhere.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;

template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int E = 4;
  __device__ static void unpack(const uint4& v, float* x) {
    x[0] = __uint_as_float(v.x); x[1] = __uint_as_float(v.y); x[2] = __uint_as_float(v.z); x[3] = __uint_as_float(v.w);
  }
  __device__ static uint4 pack(const float* x) {
    return make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int E = 8;
  __device__ static void unpack(const uint4& v, float* x) {
    x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
    x[4] = bf16_lo(v.z); x[5] = bf16_hi(v.z); x[6] = bf16_lo(v.w); x[7] = bf16_hi(v.w);
  }
  __device__ static uint4 pack(const float* x) {
    return make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
  }
};

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------- relu_bwd
template <typename T>
__global__ void __launch_bounds__(kThreads)
relu_bwd_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ act, int64_t chunks,
                float scale, uint4* __restrict__ out) {
  constexpr int E = Vec16<T>::E;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * kThreads) {
    float xa[E], xb[E], xc[E], y[E];
    Vec16<T>::unpack(__ldg(a + i), xa);
    Vec16<T>::unpack(__ldg(act + i), xc);
    if (b) {
      Vec16<T>::unpack(__ldg(b + i), xb);
#pragma unroll
      for (int k = 0; k < E; ++k) xa[k] += xb[k];
    }
#pragma unroll
    for (int k = 0; k < E; ++k) y[k] = (xc[k] > 0.f) ? xa[k] * scale : 0.f;
    out[i] = Vec16<T>::pack(y);
  }
}

// ------------------------------------------------------------------------------------------- dropout
template <typename T>
__global__ void __launch_bounds__(kThreads)
dropout_kernel(uint4* __restrict__ x, int64_t chunks, uint32_t threshold, float scale, uint64_t seed) {
  constexpr int E = Vec16<T>::E;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * kThreads) {
    float v[E];
    Vec16<T>::unpack(x[i], v);
#pragma unroll
    for (int q = 0; q < E / 4; ++q) {           // one hash per four elements (P(drop) = (threshold >> 16) / 2^16 = p)
      const uint32_t keep = dropout_keep4(seed, (uint64_t)i * (E / 4) + q, threshold);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[4 * q + k] = ((keep >> k) & 1u) ? v[4 * q + k] * scale : 0.f;
    }
    x[i] = Vec16<T>::pack(v);
  }
}

// ------------------------------------------------------------------------------------------- transpose
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ in, int64_t M, int64_t N, T* __restrict__ out, int64_t ld) {
  __shared__ T tile[64][65];
  const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
#pragma unroll
  for (int r = ty; r < 64; r += 4) {
    const int64_t m = m0 + r, n = n0 + tx;
    tile[r][tx] = (m < M && n < N) ? in[m * N + n] : T(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 64; r += 4) {
    const int64_t n = n0 + r, m = m0 + tx;
    if (m < ld && n < N) out[n * ld + m] = tile[tx][r];   // columns [M, ld) are zero padding
  }
}

// ------------------------------------------------------------------------------------------- colsum
constexpr int kChunkRows = 512;

template <typename T>
__global__ void __launch_bounds__(kThreads)
colsum_partial_kernel(const T* __restrict__ X, int64_t M, int N, float* __restrict__ partial) {
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * kChunkRows, r1 = min(r0 + kChunkRows, M);
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += to_float(X[r * N + c]);
  partial[(int64_t)blockIdx.y * N + c] = acc;
}

__global__ void __launch_bounds__(kThreads)
reduce_partials_kernel(const float* __restrict__ partial, int nchunks, int64_t width, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (i >= width) return;
  float acc = 0.f;
  for (int ch = 0; ch < nchunks; ++ch) acc += partial[(int64_t)ch * width + i];  // fixed order
  out[i] = acc;
}

// ------------------------------------------------------------------------------------------- small_wgrad
// dW[c][k] = sum_r dY[r][c] * Z[r][k]  (k < K <= KMAX), optionally dB[c] = sum_r dY[r][c].
// Block = 512 columns x kChunkRows rows; a thread owns TWO adjacent columns and keeps their K partial sums in
// registers as float2 (packed FFMA2); Z rows are staged in shared memory 64 at a time and read as broadcast
// 128-bit loads shared by both columns.
__device__ __forceinline__ float2 load2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load2(const __nv_bfloat16* p) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(bf16_lo(w), bf16_hi(w));
}

template <typename T, int KMAX>
__global__ void __launch_bounds__(kThreads)
small_wgrad_partial_kernel(const T* __restrict__ dY, const float* __restrict__ Z, int ldz, int K, int64_t M, int N,
                           float* __restrict__ partial /*[nchunks][N][KMAX+1]*/, int chunk_rows) {
  __shared__ __align__(16) float zs[64][KMAX];
  const int c = (blockIdx.x * kThreads + threadIdx.x) * 2;      // N is even (checked on the host)
  const int64_t r0 = (int64_t)blockIdx.y * chunk_rows, r1 = min(r0 + chunk_rows, M);
  float2 acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);
  for (int64_t rb = r0; rb < r1; rb += 64) {
    const int nr = (int)min((int64_t)64, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * KMAX; i += kThreads) {
      const int r = i / KMAX, k = i % KMAX;
      zs[r][k] = (r < nr && k < K) ? __ldg(Z + (rb + r) * ldz + k) : 0.f;
    }
    __syncthreads();
    if (c < N) {
#pragma unroll 2
      for (int r = 0; r < nr; ++r) {
        const float2 g = load2(dY + (rb + r) * N + c);
        accb.x += g.x; accb.y += g.y;
#pragma unroll
        for (int k4 = 0; k4 < KMAX; k4 += 4) {
          const float4 zz = *reinterpret_cast<const float4*>(&zs[r][k4]);
          acc[k4 + 0] = __ffma2_rn(g, make_float2(zz.x, zz.x), acc[k4 + 0]);
          acc[k4 + 1] = __ffma2_rn(g, make_float2(zz.y, zz.y), acc[k4 + 1]);
          acc[k4 + 2] = __ffma2_rn(g, make_float2(zz.z, zz.z), acc[k4 + 2]);
          acc[k4 + 3] = __ffma2_rn(g, make_float2(zz.w, zz.w), acc[k4 + 3]);
        }
      }
    }
  }
  if (c < N) {
    float* p0 = partial + ((int64_t)blockIdx.y * N + c) * (KMAX + 1);
    float* p1 = p0 + (KMAX + 1);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { p0[k] = acc[k].x; p1[k] = acc[k].y; }
    p0[KMAX] = accb.x; p1[KMAX] = accb.y;
  }
}

template <int KMAX>
__global__ void __launch_bounds__(kThreads)
small_wgrad_reduce_kernel(const float* __restrict__ partial, int nchunks, int N, int K, float* __restrict__ dW,
                          float* __restrict__ dB) {
  const int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x;  // over N * (KMAX+1)
  if (i >= (int64_t)N * (KMAX + 1)) return;
  const int c = (int)(i / (KMAX + 1)), k = (int)(i % (KMAX + 1));
  if (k < KMAX && k >= K) return;
  if (k == KMAX && !dB) return;
  float acc = 0.f;
  for (int ch = 0; ch < nchunks; ++ch) acc += partial[(int64_t)ch * N * (KMAX + 1) + i];
  if (k == KMAX) dB[c] = acc; else dW[(int64_t)c * K + k] = acc;
}

// ------------------------------------------------------------------------------------------- head_mask_bwd
// draw = d(10 * raw / max(|raw|, eps)) ; dH[row, :] = (draw . W) * scale * (Hact > 0).
// A block covers the whole feature row: lane l of warp w owns the 16-byte chunk 32*w + l of EVERY row the block
// visits, so the three head-weight values of its columns live in 3*E registers (not 3*E*chunks-per-lane as with a warp
// per row) and the R independent row loads of a group are all in flight before the first is used -- the kernel is a
// pure stream (read Hact once, write dH once) and needs that memory parallelism to approach the HBM roofline.
// CS: per-block partial column sums of dH (fp32, before the output rounding) -- the bias gradient of the layer under
// the head -- without a pass that reads dH back; a column belongs to one thread of the block, rows are visited in
// a fixed order: deterministic (cs_partials [gridDim.x][Hdim + 4], combined by colsum_finish_kernel); columns
// Hdim .. Hdim+2 of a partial row carry the block's column sums of draw (the head's bias gradient).
template <typename T, int R, bool CS>
__global__ void __launch_bounds__(256)
head_mask_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ raw, const T* __restrict__ Hact,
                     int32_t rows, int32_t Hdim, const float* __restrict__ W, float scale, T* __restrict__ dH,
                     float* __restrict__ draw_out, __nv_bfloat16* __restrict__ draw_bf16, float* __restrict__ cs_partials) {
  constexpr int E = Vec16<T>::E;
  static_assert(R <= 32 && (R % 4) == 0, "a group's rows are handled by the first R lanes");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = Hdim / E;
  const int ch = warp * 32 + lane;
  const bool live = ch < chunks;
  float w[3][E], cs[E];
  float ds0 = 0.f, ds1 = 0.f, ds2 = 0.f;      // lane i < R: sums of draw over rows r0 + i of this block's groups
#pragma unroll
  for (int k = 0; k < E; ++k) {
    cs[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j][k] = live ? __ldg(W + (int64_t)j * Hdim + ch * E + k) : 0.f;
  }
  for (int64_t r0 = (int64_t)blockIdx.x * R; r0 < rows; r0 += (int64_t)gridDim.x * R) {
    uint4 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i)
      v[i] = (live && r0 + i < rows) ? __ldg(reinterpret_cast<const uint4*>(Hact + (r0 + i) * Hdim) + ch) : make_uint4(0, 0, 0, 0);
    // lane i < R: the normalise-Jacobian of row r0 + i (every warp redoes these 24 bytes per row: L1 hits)
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    const int64_t myrow = r0 + lane;
    if (lane < R && myrow < rows) {
      const float x0 = __ldg(raw + myrow * 3), x1 = __ldg(raw + myrow * 3 + 1), x2 = __ldg(raw + myrow * 3 + 2);
      const float g0 = __ldg(dlogits + myrow * 3), g1 = __ldg(dlogits + myrow * 3 + 1), g2 = __ldg(dlogits + myrow * 3 + 2);
      const float nrm = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
      if (nrm > 1e-12f) {
        const float inv = 1.f / nrm;
        const float u0 = x0 * inv, u1 = x1 * inv, u2 = x2 * inv;
        const float dot = u0 * g0 + u1 * g1 + u2 * g2;
        const float s = 10.f * inv;
        d0 = s * (g0 - u0 * dot); d1 = s * (g1 - u1 * dot); d2 = s * (g2 - u2 * dot);
      } else {  // clamp branch of F.normalize: the denominator is the constant eps
        d0 = 1e13f * g0; d1 = 1e13f * g1; d2 = 1e13f * g2;
      }
      if (warp == 0 && draw_out) { draw_out[myrow * 3] = d0; draw_out[myrow * 3 + 1] = d1; draw_out[myrow * 3 + 2] = d2; }
      if constexpr (CS) { ds0 += d0; ds1 += d1; ds2 += d2; }
    }
    if (warp == 0 && draw_bf16) {   // [d0 d1 d2 0 ... 0] as one 128-byte row: the MN-major operand of the head's weight gradient
#pragma unroll
      for (int i = 0; i < R / 4; ++i) {          // lane -> (row 4*i + lane / 8, 16-byte piece lane % 8)
        const int rr = 4 * i + (lane >> 3);
        const float e0 = __shfl_sync(0xffffffffu, d0, rr), e1 = __shfl_sync(0xffffffffu, d1, rr), e2 = __shfl_sync(0xffffffffu, d2, rr);
        uint4 o = make_uint4(0, 0, 0, 0);
        if ((lane & 7) == 0) {
          o.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e0)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e1)) << 16);
          o.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e2));
        }
        if (r0 + rr < rows) reinterpret_cast<uint4*>(draw_bf16 + (r0 + rr) * 64)[lane & 7] = o;
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float e0 = __shfl_sync(0xffffffffu, d0, i), e1 = __shfl_sync(0xffffffffu, d1, i), e2 = __shfl_sync(0xffffffffu, d2, i);
      if (live && r0 + i < rows) {
        float x[E], y[E];
        Vec16<T>::unpack(v[i], x);
#pragma unroll
        for (int k = 0; k < E; ++k) {
          const float g = fmaf(e0, w[0][k], fmaf(e1, w[1][k], e2 * w[2][k]));
          y[k] = (x[k] > 0.f) ? g * scale : 0.f;
          if constexpr (CS) cs[k] += y[k];
        }
        reinterpret_cast<uint4*>(dH + (r0 + i) * Hdim)[ch] = Vec16<T>::pack(y);
      }
    }
  }
  if constexpr (CS) {
    float* prow = cs_partials + (int64_t)blockIdx.x * (Hdim + 4);
    if (live) {
#pragma unroll
      for (int k = 0; k < E; ++k) prow[ch * E + k] = cs[k];
    }
    if (warp == 0) {                          // lanes >= R hold zeros: a fixed-shape butterfly, deterministic
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ds0 += __shfl_xor_sync(0xffffffffu, ds0, o); ds1 += __shfl_xor_sync(0xffffffffu, ds1, o); ds2 += __shfl_xor_sync(0xffffffffu, ds2, o);
      }
      if (lane == 0) { prow[Hdim] = ds0; prow[Hdim + 1] = ds1; prow[Hdim + 2] = ds2; prow[Hdim + 3] = 0.f; }
    }
  }
}

// Column sums of the [nrows][ld] partials in a fixed order: block = 32 columns x 32 row strides, then 32 -> 1 in shared
// memory.  Columns < N0 go to out0 (if given), columns N0 .. N0+2 to out1 (if given).
__global__ void __launch_bounds__(1024)
colsum_finish_kernel(const float* __restrict__ partials, int nrows, int ld, int N0, float* __restrict__ out0,
                     float* __restrict__ out1) {
  __shared__ float red[32][33];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  float acc = 0.f;
  if (j < N0 + 3)
    for (int i = r; i < nrows; i += 32) acc += partials[(int64_t)i * ld + j];
  red[r][c] = acc;
  __syncthreads();
  if (r == 0 && j < N0 + 3) {
    float s = red[0][c];
#pragma unroll
    for (int i = 1; i < 32; ++i) s += red[i][c];
    if (j < N0) { if (out0) out0[j] = s; }
    else if (out1) out1[j - N0] = s;
  }
}

int grid_stride(int64_t chunks) {
  const int64_t want = (chunks + kThreads - 1) / kThreads, cap = (int64_t)sm_count() * 16;
  return (int)(want < cap ? want : cap);
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

#define LPGNN_DT_OK(dt, name) LPGNN_REQUIRE(dt == LPGNN_F32 || dt == LPGNN_BF16, name ": bad dtype %d", dt)

extern "C" int lpgnn_relu_bwd(const void* a, const void* b, const void* act, int64_t count, int dtype, float scale,
                              void* out, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_DT_OK(dtype, "relu_bwd");
  const int E = dtype == LPGNN_F32 ? 4 : 8;
  LPGNN_REQUIRE(count >= 0 && count % E == 0, "relu_bwd: count=%lld must be a multiple of %d", (long long)count, E);
  if (count == 0) return LPGNN_OK;
  LPGNN_REQUIRE(a && act && out, "relu_bwd: null pointer");
  const int64_t chunks = count / E;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LPGNN_F32)
    relu_bwd_kernel<float><<<grid_stride(chunks), kThreads, 0, st>>>((const uint4*)a, (const uint4*)b, (const uint4*)act,
                                                                      chunks, scale, (uint4*)out);
  else
    relu_bwd_kernel<__nv_bfloat16><<<grid_stride(chunks), kThreads, 0, st>>>((const uint4*)a, (const uint4*)b,
                                                                             (const uint4*)act, chunks, scale, (uint4*)out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_dropout(void* x, int64_t count, int dtype, float p, uint64_t seed, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_DT_OK(dtype, "dropout");
  LPGNN_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f outside [0,1)", p);
  const int E = dtype == LPGNN_F32 ? 4 : 8;
  LPGNN_REQUIRE(count >= 0 && count % E == 0, "dropout: count must be a multiple of %d", E);
  if (count == 0 || p == 0.f) return LPGNN_OK;
  LPGNN_REQUIRE(x, "dropout: null pointer");
  const int64_t chunks = count / E;
  const uint32_t threshold = (uint32_t)((double)p * 4294967296.0);
  const float scale = 1.f / (1.f - p);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LPGNN_F32)
    dropout_kernel<float><<<grid_stride(chunks), kThreads, 0, st>>>((uint4*)x, chunks, threshold, scale, seed);
  else
    dropout_kernel<__nv_bfloat16><<<grid_stride(chunks), kThreads, 0, st>>>((uint4*)x, chunks, threshold, scale, seed);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_transpose(const void* X, int dtype, int64_t M, int64_t N, void* out, int64_t ld_out,
                               lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_DT_OK(dtype, "transpose");
  LPGNN_REQUIRE(M >= 0 && N >= 0 && ld_out >= M, "transpose: bad shape M=%lld N=%lld ld=%lld", (long long)M,
                (long long)N, (long long)ld_out);
  if (ld_out == 0 || N == 0) return LPGNN_OK;
  LPGNN_REQUIRE(X && out, "transpose: null pointer");
  LPGNN_REQUIRE(ceil_div(ld_out, 64) <= 65535, "transpose: ld_out=%lld exceeds the %d rows one launch covers",
                (long long)ld_out, 65535 * 64);
  dim3 grid(ceil_div(N, 64), ceil_div(ld_out, 64));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LPGNN_F32) transpose_kernel<float><<<grid, 256, 0, st>>>((const float*)X, M, N, (float*)out, ld_out);
  else transpose_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)X, M, N, (__nv_bfloat16*)out, ld_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" size_t lpgnn_colsum_workspace_bytes(int64_t M, int32_t N) {
  return (size_t)ceil_div(M > 0 ? M : 1, kChunkRows) * (size_t)N * sizeof(float);
}

extern "C" int lpgnn_colsum(const void* X, int dtype, int64_t M, int32_t N, float* out, void* workspace,
                            size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_DT_OK(dtype, "colsum");
  LPGNN_REQUIRE(M >= 0 && N > 0 && out, "colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) { LPGNN_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * N, st)); return LPGNN_OK; }
  if (workspace_bytes < lpgnn_colsum_workspace_bytes(M, N)) { set_error("colsum: workspace too small"); return LPGNN_EWORKSPACE; }
  const int nchunks = ceil_div(M, kChunkRows);
  dim3 grid(ceil_div(N, kThreads), nchunks);
  float* partial = reinterpret_cast<float*>(workspace);
  if (dtype == LPGNN_F32) colsum_partial_kernel<float><<<grid, kThreads, 0, st>>>((const float*)X, M, N, partial);
  else colsum_partial_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)X, M, N, partial);
  reduce_partials_kernel<<<ceil_div(N, kThreads), kThreads, 0, st>>>(partial, nchunks, N, out);
  LPGNN_LAUNCH_OK();
  count_launches(2);
  return LPGNN_OK;
}

static int wgrad_kmax(int K) { return K <= 4 ? 4 : (K <= 16 ? 16 : (K <= 32 ? 32 : 64)); }

// Rows per partial-sum chunk: 512 for big inputs, down to 64 when that is what it takes to put ~4 blocks on every SM
// (small LPs would otherwise run a handful of long serial loops).
static int small_wgrad_chunk_rows(int64_t M, int N) {
  const int64_t colblocks = ceil_div(N, 2 * kThreads);
  int64_t rows = (M * colblocks + 4 * sm_count() - 1) / (4 * (int64_t)sm_count());
  rows = (rows + 63) / 64 * 64;
  return (int)(rows < 64 ? 64 : (rows > kChunkRows ? kChunkRows : rows));
}

extern "C" size_t lpgnn_small_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K) {
  const int64_t m = M > 0 ? M : 1;
  (void)check_device();   // the chunking depends on the SM count: make sure it is the device's
  return (size_t)ceil_div(m, small_wgrad_chunk_rows(m, N)) * (size_t)N * (wgrad_kmax(K) + 1) * sizeof(float);
}

template <typename T, int KMAX>
static int small_wgrad_launch(const void* dY, const float* Z, int ldz, int K, int64_t M, int N, float* dW, float* dB,
                              float* partial, cudaStream_t st) {
  const int chunk = small_wgrad_chunk_rows(M, N);
  const int nchunks = ceil_div(M, chunk);
  dim3 grid(ceil_div(N, 2 * kThreads), nchunks);
  small_wgrad_partial_kernel<T, KMAX><<<grid, kThreads, 0, st>>>((const T*)dY, Z, ldz, K, M, N, partial, chunk);
  small_wgrad_reduce_kernel<KMAX><<<ceil_div((int64_t)N * (KMAX + 1), kThreads), kThreads, 0, st>>>(partial, nchunks, N,
                                                                                                    K, dW, dB);
  LPGNN_LAUNCH_OK();
  count_launches(2);
  return LPGNN_OK;
}

extern "C" int lpgnn_small_wgrad(const void* dY, int dtype, const float* Z, int32_t ldz, int32_t K, int64_t M,
                                 int32_t N, float* dW, float* dB, void* workspace, size_t workspace_bytes,
                                 lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_DT_OK(dtype, "small_wgrad");
  LPGNN_REQUIRE(M >= 0 && N > 0 && N % 2 == 0 && K >= 1 && K <= 64 && ldz >= K && dW,
                "small_wgrad: bad arguments (N=%d must be even, K=%d ldz=%d)", N, K, ldz);
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) {
    LPGNN_CUDA_OK(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, st));
    if (dB) LPGNN_CUDA_OK(cudaMemsetAsync(dB, 0, sizeof(float) * N, st));
    return LPGNN_OK;
  }
  LPGNN_REQUIRE(dY && Z && workspace, "small_wgrad: null pointer");
  if (workspace_bytes < lpgnn_small_wgrad_workspace_bytes(M, N, K)) { set_error("small_wgrad: workspace too small"); return LPGNN_EWORKSPACE; }
  float* partial = reinterpret_cast<float*>(workspace);
  const int kmax = wgrad_kmax(K);
  const bool f32 = dtype == LPGNN_F32;
#define LPGNN_WG(KM) (f32 ? small_wgrad_launch<float, KM>(dY, Z, ldz, K, M, N, dW, dB, partial, st) \
                          : small_wgrad_launch<__nv_bfloat16, KM>(dY, Z, ldz, K, M, N, dW, dB, partial, st))
  if (kmax == 4) return LPGNN_WG(4);
  if (kmax == 16) return LPGNN_WG(16);
  if (kmax == 32) return LPGNN_WG(32);
  return LPGNN_WG(64);
#undef LPGNN_WG
}

constexpr int kHeadBwdRows = 8;     // rows of one group: independent 16-byte loads in flight per lane

static int head_bwd_grid(int32_t rows) {
  // also the number of partial rows the column-sum finishing kernel reads
  return min(ceil_div(rows, kHeadBwdRows), sm_count() * 8);
}

template <typename T>
static int head_bwd_dispatch(const float* dlogits, const float* raw, const void* Hact, int32_t rows, int32_t Hdim,
                             const float* W, float scale, void* dH, float* draw, void* draw_bf16, int warps, float* cs_partials,
                             cudaStream_t st) {
  const int grid = head_bwd_grid(rows);
  if (cs_partials)
    head_mask_bwd_kernel<T, kHeadBwdRows, true><<<grid, 32 * warps, 0, st>>>(dlogits, raw, (const T*)Hact, rows, Hdim, W, scale, (T*)dH,
                                                                            draw, (__nv_bfloat16*)draw_bf16, cs_partials);
  else
    head_mask_bwd_kernel<T, kHeadBwdRows, false><<<grid, 32 * warps, 0, st>>>(dlogits, raw, (const T*)Hact, rows, Hdim, W, scale, (T*)dH,
                                                                             draw, (__nv_bfloat16*)draw_bf16, nullptr);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" size_t lpgnn_head_mask_bwd_colsum_workspace_bytes(int32_t rows, int32_t Hdim) {
  return (size_t)head_bwd_grid(rows > 0 ? rows : 1) * (size_t)(Hdim + 4) * sizeof(float) + 256;
}

extern "C" int lpgnn_head_mask_bwd_colsum(const float* dlogits, const float* raw, const void* Hact, int h_dtype, int32_t rows,
                                          int32_t Hdim, const float* W, float scale, void* dH, float* draw,
                                          void* draw_bf16, float* colsum_out, float* draw_colsum_out, void* workspace,
                                          size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_DT_OK(h_dtype, "head_mask_bwd");
  LPGNN_REQUIRE(rows >= 0 && Hdim > 0, "head_mask_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (colsum_out) LPGNN_CUDA_OK(cudaMemsetAsync(colsum_out, 0, sizeof(float) * (size_t)Hdim, st));
    if (draw_colsum_out) LPGNN_CUDA_OK(cudaMemsetAsync(draw_colsum_out, 0, sizeof(float) * 3, st));
    return LPGNN_OK;
  }
  LPGNN_REQUIRE(dlogits && raw && Hact && W && dH, "head_mask_bwd: null pointer");
  const int esz = h_dtype == LPGNN_F32 ? 4 : 2;
  LPGNN_REQUIRE((Hdim * esz) % 16 == 0, "head_mask_bwd: row bytes must be a multiple of 16");
  const int ch = (Hdim * esz / 16 + 31) / 32;    // warps per block: one 16-byte chunk of the row per lane
  LPGNN_REQUIRE(ch <= 8, "head_mask_bwd: Hdim=%d too wide (max 4096 bytes per row)", Hdim);
  LPGNN_REQUIRE((uintptr_t)Hact % 16 == 0 && (uintptr_t)dH % 16 == 0, "head_mask_bwd: Hact / dH must be 16-byte aligned");
  float* partials = nullptr;
  if (colsum_out || draw_colsum_out) {
    if (!workspace || workspace_bytes < lpgnn_head_mask_bwd_colsum_workspace_bytes(rows, Hdim)) {
      set_error("head_mask_bwd_colsum: workspace too small");
      return LPGNN_EWORKSPACE;
    }
    partials = reinterpret_cast<float*>(workspace);
  }
  int rc = h_dtype == LPGNN_F32
               ? head_bwd_dispatch<float>(dlogits, raw, Hact, rows, Hdim, W, scale, dH, draw, draw_bf16, ch, partials, st)
               : head_bwd_dispatch<__nv_bfloat16>(dlogits, raw, Hact, rows, Hdim, W, scale, dH, draw, draw_bf16, ch, partials, st);
  if (rc || !partials) return rc;
  const int first = colsum_out ? 0 : Hdim / 32;     // only the draw columns when the dH sums are not wanted
  colsum_finish_kernel<<<ceil_div(Hdim + 3, 32) - first, 1024, 0, st>>>(partials + first * 32, head_bwd_grid(rows), Hdim + 4,
                                                                         Hdim - first * 32, colsum_out ? colsum_out : nullptr,
                                                                         draw_colsum_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_head_mask_bwd(const float* dlogits, const float* raw, const void* Hact, int h_dtype, int32_t rows,
                                   int32_t Hdim, const float* W, float scale, void* dH, float* draw,
                                   void* draw_bf16, lpgnn_stream_t stream) {
  return lpgnn_head_mask_bwd_colsum(dlogits, raw, Hact, h_dtype, rows, Hdim, W, scale, dH, draw, draw_bf16, nullptr, nullptr,
                                    nullptr, 0, stream);
}
