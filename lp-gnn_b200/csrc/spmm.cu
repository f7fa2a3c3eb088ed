// (a2) Weighted-sum aggregation  Y[i,:] = sum_e val[e] * X[idx[e],:]   (CSR or CSC orientation).
//
// Replaces torch_sparse spmm_sum reached from PyG GraphConv (reference arch.py:75-80), forward
// and backward (the backward wrt the dense operand is this kernel on the other orientation, so it
// is atomics-free and deterministic).
//
// Mapping: a group of G lanes (G = 32 for wide features, a sub-warp for narrow ones) owns one
// output row; each lane owns CH 16-byte chunks of the feature row (chunk c of lane l is chunk
// l + c*G), accumulates in fp32 registers over the row's nonzeros IN CSR ORDER (same order as the
// reference CPU kernel), and writes the row once.  (idx,val) pairs are fetched G at a time with
// one coalesced load and broadcast with shuffles; feature rows are gathered with 128-bit loads,
// UNROLL neighbours in flight per lane.  HBM/L2-bound: see DESIGN.md for the byte model.
#include "common.cuh"

namespace lpgnn {
namespace {

template <typename T> struct Chunk;  // 16-byte chunk of features
template <> struct Chunk<float> {
  static constexpr int kElems = 4;
  __device__ static void fma(float (&acc)[4], float w, const uint4& v) {
    acc[0] = fmaf(w, __uint_as_float(v.x), acc[0]);
    acc[1] = fmaf(w, __uint_as_float(v.y), acc[1]);
    acc[2] = fmaf(w, __uint_as_float(v.z), acc[2]);
    acc[3] = fmaf(w, __uint_as_float(v.w), acc[3]);
  }
  __device__ static uint4 pack(const float (&acc)[4]) {
    return make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]),
                      __float_as_uint(acc[3]));
  }
};
template <> struct Chunk<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ static void fma(float (&acc)[8], float w, const uint4& v) {
    // plain FFMA: the packed FFMA2 form measured slower here (register-pair moves outweigh the saved issue slots)
    acc[0] = fmaf(w, bf16_lo(v.x), acc[0]); acc[1] = fmaf(w, bf16_hi(v.x), acc[1]);
    acc[2] = fmaf(w, bf16_lo(v.y), acc[2]); acc[3] = fmaf(w, bf16_hi(v.y), acc[3]);
    acc[4] = fmaf(w, bf16_lo(v.z), acc[4]); acc[5] = fmaf(w, bf16_hi(v.z), acc[5]);
    acc[6] = fmaf(w, bf16_lo(v.w), acc[6]); acc[7] = fmaf(w, bf16_hi(v.w), acc[7]);
  }
  __device__ static uint4 pack(const float (&acc)[8]) {
    return make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                      pack_bf16(acc[6], acc[7]));
  }
};

constexpr int kThreads = 256;

// chunks_per_row = F*sizeof(T)/16.  G lanes per row, CH chunks per lane (G*CH >= chunks_per_row).
template <typename T, int G, int CH>
__global__ void __launch_bounds__(kThreads)
spmm_rows_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                 int32_t rows, const uint4* __restrict__ X, uint4* __restrict__ Y, int32_t chunks_per_row) {
  constexpr int E = Chunk<T>::kElems;
  constexpr int kRowsPerBlock = kThreads / G;
  const int lane_in_warp = threadIdx.x & 31;
  const int g_lane = threadIdx.x % G;                  // lane inside the row group
  const int group_base = lane_in_warp - g_lane;        // first lane of the group inside the warp
  constexpr uint32_t kGroupBits = (G >= 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  const uint32_t gmask = kGroupBits << group_base;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + threadIdx.x / G;
  if (row >= rows) return;  // whole group exits together (row is per group)

  float acc[CH][E];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int k = 0; k < E; ++k) acc[c][k] = 0.f;

  const int32_t beg = ptr[row], end = ptr[row + 1];
  for (int32_t e0 = beg; e0 < end; e0 += G) {
    // one coalesced fetch of up to G (idx,val) pairs, then broadcast in order
    int32_t my_idx = 0; float my_val = 0.f;
    if (e0 + g_lane < end) { my_idx = __ldg(idx + e0 + g_lane); my_val = __ldg(val + e0 + g_lane); }
    const int cnt = min(G, end - e0);
    int j = 0;
    for (; j + 2 <= cnt; j += 2) {  // two neighbours in flight
      const int32_t i0 = __shfl_sync(gmask, my_idx, group_base + j);
      const int32_t i1 = __shfl_sync(gmask, my_idx, group_base + j + 1);
      const float w0 = __shfl_sync(gmask, my_val, group_base + j);
      const float w1 = __shfl_sync(gmask, my_val, group_base + j + 1);
      uint4 v0[CH], v1[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int ch = g_lane + c * G;
        if (ch < chunks_per_row) {
          v0[c] = __ldg(X + (int64_t)i0 * chunks_per_row + ch);
          v1[c] = __ldg(X + (int64_t)i1 * chunks_per_row + ch);
        } else {
          v0[c] = make_uint4(0, 0, 0, 0); v1[c] = make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) Chunk<T>::fma(acc[c], w0, v0[c]);
#pragma unroll
      for (int c = 0; c < CH; ++c) Chunk<T>::fma(acc[c], w1, v1[c]);
    }
    if (j < cnt) {
      const int32_t i0 = __shfl_sync(gmask, my_idx, group_base + j);
      const float w0 = __shfl_sync(gmask, my_val, group_base + j);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int ch = g_lane + c * G;
        if (ch < chunks_per_row) Chunk<T>::fma(acc[c], w0, __ldg(X + (int64_t)i0 * chunks_per_row + ch));
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int ch = g_lane + c * G;
    if (ch < chunks_per_row) Y[row * chunks_per_row + ch] = Chunk<T>::pack(acc[c]);
  }
}

template <typename T, int G, int CH>
int launch(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const void* X, void* Y,
           int32_t chunks, cudaStream_t st) {
  constexpr int kRowsPerBlock = kThreads / G;
  const int grid = ceil_div(rows, kRowsPerBlock);
  spmm_rows_kernel<T, G, CH><<<grid, kThreads, 0, st>>>(ptr, idx, val, rows, reinterpret_cast<const uint4*>(X),
                                                        reinterpret_cast<uint4*>(Y), chunks);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

template <typename T>
int dispatch(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const void* X, void* Y,
             int32_t chunks, cudaStream_t st) {
  if (chunks <= 1) return launch<T, 1, 1>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 2) return launch<T, 2, 1>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 4) return launch<T, 4, 1>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 8) return launch<T, 8, 1>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 16) return launch<T, 16, 1>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 32) return launch<T, 32, 1>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 64) return launch<T, 32, 2>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 128) return launch<T, 32, 4>(ptr, idx, val, rows, X, Y, chunks, st);
  if (chunks <= 256) return launch<T, 32, 8>(ptr, idx, val, rows, X, Y, chunks, st);
  set_error("spmm: feature row of %d bytes exceeds the 4096-byte limit of one pass", chunks * 16);
  return LPGNN_EINVAL;
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" int lpgnn_spmm(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const void* X,
                          void* Y, int32_t F, int dtype, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && F > 0, "spmm: bad shape rows=%d F=%d", rows, F);
  LPGNN_REQUIRE(dtype == LPGNN_F32 || dtype == LPGNN_BF16, "spmm: bad dtype %d", dtype);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && X && Y, "spmm: null pointer");
  const int esz = dtype == LPGNN_F32 ? 4 : 2;
  LPGNN_REQUIRE((F * esz) % 16 == 0, "spmm: F*sizeof(elem)=%d must be a multiple of 16", F * esz);
  LPGNN_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0), "spmm: X/Y must be 16-byte aligned");
  const int32_t chunks = F * esz / 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LPGNN_F32) return dispatch<float>(ptr, idx, val, rows, X, Y, chunks, st);
  return dispatch<__nv_bfloat16>(ptr, idx, val, rows, X, Y, chunks, st);
}
