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
#include <stdlib.h>

#include "common.cuh"

#include <type_traits>

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

template <> struct Chunk<__half> {
  static constexpr int kElems = 8;
  __device__ static void fma(float (&acc)[8], float w, const uint4& v) {
    const float2 a = f16x2_unpack(v.x), b = f16x2_unpack(v.y), c = f16x2_unpack(v.z), d = f16x2_unpack(v.w);
    acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
    acc[2] = fmaf(w, b.x, acc[2]); acc[3] = fmaf(w, b.y, acc[3]);
    acc[4] = fmaf(w, c.x, acc[4]); acc[5] = fmaf(w, c.y, acc[5]);
    acc[6] = fmaf(w, d.x, acc[6]); acc[7] = fmaf(w, d.y, acc[7]);
  }
  __device__ static uint4 pack(const float (&acc)[8]) {
    return make_uint4(pack_f16(acc[0], acc[1]), pack_f16(acc[2], acc[3]), pack_f16(acc[4], acc[5]),
                      pack_f16(acc[6], acc[7]));
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


// ---------------------------------------------------------------------------------------------------------------
// Wide features: banded sweep.
//
// The row-per-warp kernel above moves the gather-model bytes (z*F*s) from L2 to the SMs; on C2 that is 1 GB per
// direction at ~11.7 TB/s, the measured L2 throughput cap (ncu: L1 hit rate 8 %), i.e. 0.44 of the HBM roofline.
// LP matrices are banded in their natural ordering (staircase / time-indexed structure), so the source rows gathered
// by neighbouring output rows overlap.  Here ONE 1024-thread CTA per SM owns one SLAB of the feature row (32 lanes x
// CH x 16 bytes: 512 B or 1 KB) and a contiguous range of output rows, and sweeps that range top to bottom with its 32
// warps on 32 adjacent rows.  With the slab narrow enough, the active band (band rows x slab bytes) stays resident
// in the SM's L1, so every source-row slab travels L2 -> SM about once per sweep instead of once per nonzero (ncu on
// C2: L1 hit rate 60-72 %, L2->L1 sectors / 2.6-6, DRAM traffic = compulsory).  Matrices without locality lose
// nothing (measured: uniform-random C2 6 % faster than the row kernel).
//   * per gathered 16-byte chunk: 2 SHFL + 1 IMAD.WIDE + 1 LDG.128 + 8 unpack (4 IMAD.SHL on the FMA pipe, 4 LOP3 on the
//     ALU pipe) + 4 FFMA2 (packed two-lane FMA, per lane the same rounding as fmaf);
//   * (ptr) two rows ahead and the row's first (idx,val) fetch one row ahead are software-pipelined, the tail of a
//     row is gathered in batches of U, U/2, .., 1 so no load is issued alone behind another's latency.
// Same arithmetic as the row kernel: fp32 accumulation in CSR order, one owner per output element, no atomics --
// results are bit-identical between the two kernels and for every tuning.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct PairAcc;   // 16-byte chunk -> fp32 pair accumulators
template <> struct PairAcc<__nv_bfloat16> {
  static constexpr int kPairs = 4;
  __device__ static float2 unpack(uint32_t x) {
    return make_float2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
  }
  __device__ static void fma(float2 (&acc)[4], float2 ww, const uint4& v) {
    acc[0] = __ffma2_rn(ww, unpack(v.x), acc[0]);
    acc[1] = __ffma2_rn(ww, unpack(v.y), acc[1]);
    acc[2] = __ffma2_rn(ww, unpack(v.z), acc[2]);
    acc[3] = __ffma2_rn(ww, unpack(v.w), acc[3]);
  }
  __device__ static uint4 pack(const float2 (&acc)[4]) {
    return make_uint4(pack_bf16(acc[0].x, acc[0].y), pack_bf16(acc[1].x, acc[1].y), pack_bf16(acc[2].x, acc[2].y),
                      pack_bf16(acc[3].x, acc[3].y));
  }
};
template <> struct PairAcc<__half> {
  static constexpr int kPairs = 4;
  __device__ static void fma(float2 (&acc)[4], float2 ww, const uint4& v) {
    acc[0] = __ffma2_rn(ww, f16x2_unpack(v.x), acc[0]);
    acc[1] = __ffma2_rn(ww, f16x2_unpack(v.y), acc[1]);
    acc[2] = __ffma2_rn(ww, f16x2_unpack(v.z), acc[2]);
    acc[3] = __ffma2_rn(ww, f16x2_unpack(v.w), acc[3]);
  }
  __device__ static uint4 pack(const float2 (&acc)[4]) {
    return make_uint4(pack_f16(acc[0].x, acc[0].y), pack_f16(acc[1].x, acc[1].y), pack_f16(acc[2].x, acc[2].y),
                      pack_f16(acc[3].x, acc[3].y));
  }
};
template <> struct PairAcc<float> {
  static constexpr int kPairs = 2;
  __device__ static void fma(float2 (&acc)[2], float2 ww, const uint4& v) {
    acc[0] = __ffma2_rn(ww, make_float2(__uint_as_float(v.x), __uint_as_float(v.y)), acc[0]);
    acc[1] = __ffma2_rn(ww, make_float2(__uint_as_float(v.z), __uint_as_float(v.w)), acc[1]);
  }
  __device__ static uint4 pack(const float2 (&acc)[2]) {
    return make_uint4(__float_as_uint(acc[0].x), __float_as_uint(acc[0].y), __float_as_uint(acc[1].x),
                      __float_as_uint(acc[1].y));
  }
};

constexpr int kSweepThreads = 1024;

// Each lane owns CH 16-byte chunks (chunk c at lane*16 + c*512 bytes of the slab); a warp owns one row at a time.
// kX2 (fp32 features only): the aggregate is written as x2 operands of the fp32 tensor-core transform (gemm_x2.cu) instead
// of fp32 -- IEEE-half hi / lo rows and one power-of-two scale per row -- so no separate split pass reads it back.  The
// scale comes from a bound that is known when the row is: |Y[i,:]| <= sum_e |val[e]| * s_src[idx[e]] * 2^12, where
// |X[j,:]| <= s_src[j] * 2^12 is the guarantee the producer of X gave (lpgnn_conv_in_fused_x2).
struct X2Out {
  const float* src_scale;   // [n_src]
  char* hi;                 // [rows, F] half
  char* lo;
  float* scale;             // [rows]
};

__device__ __forceinline__ float x2_scale_for_bound(float bound) {   // as in conv_in.cu: bound < 2^e -> scale 2^(e-12)
  int e = 12;
  if (bound > 0.f && bound < __int_as_float(0x7f800000)) e = (int)((__float_as_uint(bound) >> 23) & 0xffu) - 127 + 1;
  e = max(-100, min(e, 112));
  return __int_as_float((uint32_t)(127 + e - 12) << 23);
}

// One aggregation: Y[rows, F] = A_view X.  (Both directions of a layer are independent and equally long at LP shapes, so
// a launch can carry two of them: half of the SMs each, one ramp and one tail instead of two.)
struct SweepSide {
  const int32_t* ptr; const int32_t* idx; const float* val; int32_t rows;
  const char* X; char* Y; int32_t rows_per_block;
  X2Out x2;
};

template <typename T, int CH, int U, bool kX2>
__device__ __forceinline__ void sweep_side(const SweepSide& S, const int side_block, const uint32_t row_bytes, const int32_t nslabs) {
  const int32_t* __restrict__ ptr = S.ptr;
  const int32_t* __restrict__ idx = S.idx;
  const float* __restrict__ val = S.val;
  const int32_t rows = S.rows;
  const char* __restrict__ X = S.X;
  char* __restrict__ Y = S.Y;
  const int32_t rows_per_block = S.rows_per_block;
  const X2Out& x2 = S.x2;
  constexpr int P = PairAcc<T>::kPairs;
  constexpr int kWarps = kSweepThreads / 32;
  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slab = side_block % nslabs;
  const int32_t r_begin = (side_block / nslabs) * rows_per_block;
  const int32_t r_end = min(rows, r_begin + rows_per_block);
  const uint32_t col0 = (uint32_t)(slab * 32 * CH + lane) * 16u;    // byte column of this lane's first chunk
  const uint64_t xbase = reinterpret_cast<uint64_t>(X) + col0;      // (row_bytes is a multiple of the slab: no bounds tests)

  int32_t row = r_begin + warp;
  int32_t beg = 0, end = 0, nbeg = 0, nend = 0;
  int32_t my_idx = 0; float my_val = 0.f;
  if (row < r_end) { beg = __ldg(ptr + row); end = __ldg(ptr + row + 1); }
  if (row + kWarps < r_end) { nbeg = __ldg(ptr + row + kWarps); nend = __ldg(ptr + row + kWarps + 1); }
  if (beg + lane < end) { my_idx = __ldg(idx + beg + lane); my_val = __ldg(val + beg + lane); }
  float my_sb = 0.f;            // kX2: this lane's share of the row bound, sum |val| * s_src[idx]
  if (kX2 && beg + lane < end) my_sb = fabsf(my_val) * __ldg(x2.src_scale + my_idx);

  for (; row < r_end; row += kWarps) {
    int32_t nnbeg = 0, nnend = 0, n_idx = 0; float n_val = 0.f, n_sb = 0.f;
    if (row + 2 * kWarps < r_end) { nnbeg = __ldg(ptr + row + 2 * kWarps); nnend = __ldg(ptr + row + 2 * kWarps + 1); }
    if (nbeg + lane < nend) {
      n_idx = __ldg(idx + nbeg + lane); n_val = __ldg(val + nbeg + lane);
      if (kX2) n_sb = fabsf(n_val) * __ldg(x2.src_scale + n_idx);
    }

    float2 acc[CH][P];
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int k = 0; k < P; ++k) acc[c][k] = make_float2(0.f, 0.f);

    for (int32_t e0 = beg; e0 < end; e0 += 32) {
      if (e0 != beg) {
        my_idx = 0; my_val = 0.f;
        if (e0 + lane < end) {
          my_idx = __ldg(idx + e0 + lane); my_val = __ldg(val + e0 + lane);
          if (kX2) my_sb += fabsf(my_val) * __ldg(x2.src_scale + my_idx);
        }
      }
      const int cnt = min(32, end - e0);
      int j = 0;
      auto batch = [&](auto nb) {                       // nb neighbours in flight: shuffles, then loads, then math
        constexpr int N = decltype(nb)::value;
        uint4 v[N][CH]; float w[N]; uint32_t i[N];
#pragma unroll
        for (int u = 0; u < N; ++u) {
          i[u] = (uint32_t)__shfl_sync(kFull, my_idx, j + u);
          w[u] = __shfl_sync(kFull, my_val, j + u);
        }
#pragma unroll
        for (int u = 0; u < N; ++u) {
          const uint4* src = reinterpret_cast<const uint4*>(xbase + (uint64_t)i[u] * row_bytes);   // one IMAD.WIDE
#pragma unroll
          for (int c = 0; c < CH; ++c) v[u][c] = __ldg(src + c * 32);
        }
#pragma unroll
        for (int u = 0; u < N; ++u)
#pragma unroll
          for (int c = 0; c < CH; ++c) PairAcc<T>::fma(acc[c], make_float2(w[u], w[u]), v[u][c]);
        j += N;
      };
      while (j + U <= cnt) batch(std::integral_constant<int, U>{});
      if (U > 2 && j + 2 <= cnt) batch(std::integral_constant<int, 2>{});
      if (j < cnt) batch(std::integral_constant<int, 1>{});
    }
    if constexpr (kX2) {
      static_assert(!kX2 || sizeof(T) == 4, "x2 output is the fp32 path");
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) my_sb += __shfl_xor_sync(kFull, my_sb, o);
      const float sc = x2_scale_for_bound(my_sb * 4096.f);
      const float down = 1.f / sc;                            // exact: a power of two
      if (slab == 0 && lane == 0) x2.scale[row] = sc;
      const uint64_t off = ((uint64_t)(uint32_t)row * row_bytes + col0) >> 1;     // same elements, 2 bytes each
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float v[4] = {acc[c][0].x, acc[c][0].y, acc[c][1].x, acc[c][1].y};
        __half h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float sx = v[k] * down;
          h[k] = __float2half_rn(sx);
          l[k] = __float2half_rn((sx - __half2float(h[k])) * 2048.f);
        }
        const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
        const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
        *reinterpret_cast<uint2*>(x2.hi + off + c * 256) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(x2.lo + off + c * 256) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
    } else {
      uint4* dst = reinterpret_cast<uint4*>(Y + (uint64_t)(uint32_t)row * row_bytes + col0);
#pragma unroll
      for (int c = 0; c < CH; ++c) __stcs(dst + c * 32, PairAcc<T>::pack(acc[c]));
    }
    beg = nbeg; end = nend; nbeg = nnbeg; nend = nnend;
    my_idx = n_idx; my_val = n_val; my_sb = n_sb;
  }
}

template <typename T, int CH, int U, bool kX2 = false>
__global__ void __launch_bounds__(kSweepThreads, 1)
spmm_sweep_kernel(const __grid_constant__ SweepSide sa, const __grid_constant__ SweepSide sb, int blocks_a, uint32_t row_bytes,
                  int32_t nslabs) {
  if ((int)blockIdx.x < blocks_a) sweep_side<T, CH, U, kX2>(sa, (int)blockIdx.x, row_bytes, nslabs);
  else sweep_side<T, CH, U, kX2>(sb, (int)blockIdx.x - blocks_a, row_bytes, nslabs);
}

// `share` of the co-resident CTAs (one per SM) this side may use: row ranges x slabs
inline int plan_side(SweepSide& s, int nslabs, int share) {
  constexpr int kWarps = kSweepThreads / 32;
  if (s.rows <= 0) { s.rows_per_block = kWarps; return 0; }
  const int want_blocks = max(1, share / nslabs);                // row ranges of this side
  s.rows_per_block = max(ceil_div(s.rows, want_blocks), kWarps);  // even split: every SM gets the same share
  return ceil_div(s.rows, s.rows_per_block) * nslabs;
}

template <typename T, int CH, int U, bool kX2 = false>
int launch_sweep_sides(SweepSide a, SweepSide b, int32_t chunks, int64_t nnz, cudaStream_t st) {
  const int nslabs = chunks / (32 * CH);
  static bool carved = false;      // no shared memory is used: ask for the whole 228 KB as L1 (it holds the band)
  if (!carved) {
    cudaFuncSetAttribute(spmm_sweep_kernel<T, CH, U, kX2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    carved = true;
  }
  // SMs in proportion to the sides' work: every nonzero gathers one slab row, every output row writes one
  int share_a = sm_count();
  if (b.rows > 0) {
    const double wa = (double)nnz + a.rows, wb = (double)nnz + b.rows;
    share_a = (int)(sm_count() * wa / (wa + wb) + 0.5);
    share_a = max(nslabs, min(sm_count() - nslabs, share_a / nslabs * nslabs));
  }
  const int blocks_a = plan_side(a, nslabs, share_a);
  const int blocks_b = plan_side(b, nslabs, sm_count() - share_a);
  spmm_sweep_kernel<T, CH, U, kX2><<<blocks_a + blocks_b, kSweepThreads, 0, st>>>(a, b, blocks_a, (uint32_t)chunks * 16u, nslabs);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

template <typename T, int CH, int U, bool kX2 = false>
int launch_sweep(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const void* X, void* Y,
                 int32_t chunks, cudaStream_t st, const X2Out x2 = X2Out()) {
  SweepSide a{ptr, idx, val, rows, reinterpret_cast<const char*>(X), reinterpret_cast<char*>(Y), 0, x2};
  SweepSide none{};
  return launch_sweep_sides<T, CH, U, kX2>(a, none, chunks, 0, st);
}

template <typename T>
int dispatch_sweep(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const void* X, void* Y,
                   int32_t chunks, int slab_bytes, int unroll, cudaStream_t st) {
  if (slab_bytes == 512 && unroll == 2) return launch_sweep<T, 1, 2>(ptr, idx, val, rows, X, Y, chunks, st);
  if (slab_bytes == 512 && unroll == 4) return launch_sweep<T, 1, 4>(ptr, idx, val, rows, X, Y, chunks, st);
  if (slab_bytes == 1024 && unroll == 2) return launch_sweep<T, 2, 2>(ptr, idx, val, rows, X, Y, chunks, st);
  if (slab_bytes == 1024 && unroll == 4) return launch_sweep<T, 2, 4>(ptr, idx, val, rows, X, Y, chunks, st);
  set_error("spmm: unsupported tuning slab_bytes=%d unroll=%d", slab_bytes, unroll);
  return LPGNN_EINVAL;
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

extern "C" int lpgnn_spmm_ex(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const void* X,
                             void* Y, int32_t F, int dtype, int slab_bytes, int unroll, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && F > 0, "spmm: bad shape rows=%d F=%d", rows, F);
  LPGNN_REQUIRE(dtype_ok(dtype), "spmm: bad dtype %d", dtype);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && X && Y, "spmm: null pointer");
  const int esz = dtype == LPGNN_F32 ? 4 : 2;
  LPGNN_REQUIRE((F * esz) % 16 == 0, "spmm: F*sizeof(elem)=%d must be a multiple of 16", F * esz);
  LPGNN_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0), "spmm: X/Y must be 16-byte aligned");
  const int32_t chunks = F * esz / 16;
  const int row_bytes = chunks * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (slab_bytes == 0) {            // automatic: 1 KB slabs when the row is made of them, else 512 B, else the row kernel
    slab_bytes = row_bytes % 1024 == 0 ? 1024 : (row_bytes % 512 == 0 ? 512 : -1);
    // a sweep has one CTA per (SM, slab): small matrices (few rows per CTA) keep the row kernel
    if (slab_bytes > 0 && (int64_t)rows * (row_bytes / slab_bytes) < (int64_t)sm_count() * 64) slab_bytes = -1;
  }
  if (slab_bytes < 0) {             // row-per-warp kernel
    if (dtype == LPGNN_F32) return dispatch<float>(ptr, idx, val, rows, X, Y, chunks, st);
    if (dtype == LPGNN_F16) return dispatch<__half>(ptr, idx, val, rows, X, Y, chunks, st);
    return dispatch<__nv_bfloat16>(ptr, idx, val, rows, X, Y, chunks, st);
  }
  LPGNN_REQUIRE(row_bytes % slab_bytes == 0, "spmm: row of %d bytes is not a multiple of the %d-byte slab", row_bytes, slab_bytes);
  if (unroll <= 0) unroll = dtype == LPGNN_F32 ? 4 : 2;
  if (dtype == LPGNN_F32) return dispatch_sweep<float>(ptr, idx, val, rows, X, Y, chunks, slab_bytes, unroll, st);
  if (dtype == LPGNN_F16) return dispatch_sweep<__half>(ptr, idx, val, rows, X, Y, chunks, slab_bytes, unroll, st);
  return dispatch_sweep<__nv_bfloat16>(ptr, idx, val, rows, X, Y, chunks, slab_bytes, unroll, st);
}

extern "C" int lpgnn_spmm(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const void* X,
                          void* Y, int32_t F, int dtype, lpgnn_stream_t stream) {
  return lpgnn_spmm_ex(ptr, idx, val, rows, X, Y, F, dtype, 0, 0, stream);
}

// 0: lpgnn_spmm_pair / lpgnn_spmm_x2_pair always run as two launches (environment LPGNN_SPMM_PAIR=0, lpgnn_set_spmm_pair: A/B)
static int g_spmm_pair = [] { const char* e = getenv("LPGNN_SPMM_PAIR"); return e ? atoi(e) != 0 : 1; }();

extern "C" int lpgnn_set_spmm_pair(int enable) {
  const int prev = g_spmm_pair;
  g_spmm_pair = enable ? 1 : 0;
  return prev;
}

// Both aggregations of a layer in ONE launch where both take the banded sweep (agg_t [n,F] = A^T L over the CSC view,
// agg_s [m,F] = A R over the CSR view): each side gets the SMs its work asks for, the launch has one ramp and one tail.
// Anything else (small graphs, rows that are no multiple of the slab) is two lpgnn_spmm calls.  Same bits either way.
extern "C" int lpgnn_spmm_pair(const int32_t* rowptr, const int32_t* col, const float* val, int32_t m, const int32_t* colptr,
                               const int32_t* row_csc, const float* val_csc, int32_t n, int64_t nnz, const void* L,
                               const void* R, void* agg_s, void* agg_t, int32_t F, int dtype, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m >= 0 && n >= 0 && F > 0, "spmm_pair: bad shape m=%d n=%d F=%d", m, n, F);
  if (nnz < 0) nnz = 4 * (int64_t)(m > n ? m : n);     // unknown on the host: LP-typical density, only steers the SM split
  LPGNN_REQUIRE(dtype_ok(dtype), "spmm_pair: bad dtype %d", dtype);
  const int esz = dtype == LPGNN_F32 ? 4 : 2;
  const int row_bytes = F * esz;
  const int slab = row_bytes % 1024 == 0 ? 1024 : (row_bytes % 512 == 0 ? 512 : -1);
  const bool sweep = g_spmm_pair && slab > 0 && (int64_t)m * (row_bytes / slab) >= (int64_t)sm_count() * 64 &&
                     (int64_t)n * (row_bytes / slab) >= (int64_t)sm_count() * 64;
  if (!sweep) {
    if (int rc = lpgnn_spmm(colptr, row_csc, val_csc, n, L, agg_t, F, dtype, stream)) return rc;
    return lpgnn_spmm(rowptr, col, val, m, R, agg_s, F, dtype, stream);
  }
  LPGNN_REQUIRE(rowptr && colptr && L && R && agg_s && agg_t, "spmm_pair: null pointer");
  LPGNN_REQUIRE(((uintptr_t)L | (uintptr_t)R | (uintptr_t)agg_s | (uintptr_t)agg_t) % 16 == 0, "spmm_pair: X/Y must be 16-byte aligned");
  SweepSide t{colptr, row_csc, val_csc, n, reinterpret_cast<const char*>(L), reinterpret_cast<char*>(agg_t), 0, X2Out()};
  SweepSide sd{rowptr, col, val, m, reinterpret_cast<const char*>(R), reinterpret_cast<char*>(agg_s), 0, X2Out()};
  const int32_t chunks = row_bytes / 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LPGNN_F32)
    return slab == 1024 ? launch_sweep_sides<float, 2, 4>(t, sd, chunks, nnz, st) : launch_sweep_sides<float, 1, 4>(t, sd, chunks, nnz, st);
  if (dtype == LPGNN_F16)
    return slab == 1024 ? launch_sweep_sides<__half, 2, 2>(t, sd, chunks, nnz, st) : launch_sweep_sides<__half, 1, 2>(t, sd, chunks, nnz, st);
  return slab == 1024 ? launch_sweep_sides<__nv_bfloat16, 2, 2>(t, sd, chunks, nnz, st)
                      : launch_sweep_sides<__nv_bfloat16, 1, 2>(t, sd, chunks, nnz, st);
}

// (a2, fp32 tensor-core mode) Aggregation of fp32 features straight into x2 operands: hi / lo IEEE-half [rows,F] + a
// power-of-two scale per row (lpgnn_split_x2's format; Y = scale * (hi + 2^-11 lo) to 22 bits) -- the fp32 aggregate never
// reaches HBM and no split pass reads it back.  src_scale [n_src]: |X[j,:]| <= src_scale[j] * 2^12 (as written by
// lpgnn_conv_in_fused_x2).  Graphs too small for the banded sweep take lpgnn_spmm into `scratch` (fp32 [rows,F]) followed by
// lpgnn_split_x2: the same outputs either way (the scale then comes from the row maximum instead of the bound).
extern "C" int lpgnn_spmm_x2(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const float* X, int32_t F,
                             const float* src_scale, void* hi, void* lo, float* scale, float* scratch, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && F > 0 && F % 4 == 0, "spmm_x2: bad shape rows=%d F=%d", rows, F);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && X && src_scale && hi && lo && scale, "spmm_x2: null pointer");
  LPGNN_REQUIRE((uintptr_t)X % 16 == 0 && (uintptr_t)hi % 8 == 0 && (uintptr_t)lo % 8 == 0, "spmm_x2: misaligned pointer");
  const int row_bytes = F * 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (row_bytes % 1024 == 0 && (int64_t)rows * (row_bytes / 1024) >= (int64_t)sm_count() * 64) {
    X2Out x2;
    x2.src_scale = src_scale; x2.hi = reinterpret_cast<char*>(hi); x2.lo = reinterpret_cast<char*>(lo); x2.scale = scale;
    return launch_sweep<float, 2, 4, true>(ptr, idx, val, rows, X, nullptr, row_bytes / 16, st, x2);
  }
  LPGNN_REQUIRE(scratch, "spmm_x2: this shape takes the two-step path and needs the fp32 scratch [rows,F]");
  if (int rc = lpgnn_spmm(ptr, idx, val, rows, X, scratch, F, LPGNN_F32, stream)) return rc;
  return lpgnn_split_x2(scratch, F, nullptr, 0, rows, hi, lo, nullptr, nullptr, scale, stream);
}

// The two aggregations of the first hidden layer in the fp32 tensor-core mode, one launch (see lpgnn_spmm_pair /
// lpgnn_spmm_x2): side t = A^T L -> (hi_t, lo_t, scale_t), side s = A R -> (hi_s, lo_s, scale_s).
extern "C" int lpgnn_spmm_x2_pair(const int32_t* rowptr, const int32_t* col, const float* val, int32_t m, const int32_t* colptr,
                                  const int32_t* row_csc, const float* val_csc, int32_t n, int64_t nnz, const float* L,
                                  const float* R, int32_t F, const float* scale_L, const float* scale_R, void* hi_s, void* lo_s,
                                  float* scale_s, void* hi_t, void* lo_t, float* scale_t, float* scratch_s, float* scratch_t,
                                  lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m >= 0 && n >= 0 && F > 0 && F % 4 == 0, "spmm_x2_pair: bad shape m=%d n=%d F=%d", m, n, F);
  if (nnz < 0) nnz = 4 * (int64_t)(m > n ? m : n);
  const int row_bytes = F * 4;
  const bool sweep = g_spmm_pair && row_bytes % 1024 == 0 && (int64_t)m * (row_bytes / 1024) >= (int64_t)sm_count() * 64 &&
                     (int64_t)n * (row_bytes / 1024) >= (int64_t)sm_count() * 64;
  if (!sweep) {
    if (int rc = lpgnn_spmm_x2(colptr, row_csc, val_csc, n, L, F, scale_L, hi_t, lo_t, scale_t, scratch_t, stream)) return rc;
    return lpgnn_spmm_x2(rowptr, col, val, m, R, F, scale_R, hi_s, lo_s, scale_s, scratch_s, stream);
  }
  LPGNN_REQUIRE(rowptr && colptr && L && R && scale_L && scale_R && hi_s && lo_s && scale_s && hi_t && lo_t && scale_t,
                "spmm_x2_pair: null pointer");
  LPGNN_REQUIRE(((uintptr_t)L | (uintptr_t)R) % 16 == 0 && ((uintptr_t)hi_s | (uintptr_t)lo_s | (uintptr_t)hi_t | (uintptr_t)lo_t) % 8 == 0,
                "spmm_x2_pair: misaligned pointer");
  X2Out xt, xs;
  xt.src_scale = scale_L; xt.hi = reinterpret_cast<char*>(hi_t); xt.lo = reinterpret_cast<char*>(lo_t); xt.scale = scale_t;
  xs.src_scale = scale_R; xs.hi = reinterpret_cast<char*>(hi_s); xs.lo = reinterpret_cast<char*>(lo_s); xs.scale = scale_s;
  SweepSide t{colptr, row_csc, val_csc, n, reinterpret_cast<const char*>(L), nullptr, 0, xt};
  SweepSide sd{rowptr, col, val, m, reinterpret_cast<const char*>(R), nullptr, 0, xs};
  return launch_sweep_sides<float, 2, 4, true>(t, sd, row_bytes / 16, nnz, (cudaStream_t)stream);
}
