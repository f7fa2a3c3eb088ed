// (a2+a3, input layer, 16-bit modes) conv1 of GCN_FC(8, 8, ...) in ONE kernel: aggregation of the 8-wide features,
// the (8 + 8) -> hids transform, bias, ReLU and the 16-bit store.
//
// Replaces PyG GraphConv.forward for the (p,q)->hids layer and the relu_ after it (reference arch.py:170, 75-80,
// 181-182).  The output write dominates ((m+n) * hids * 2 bytes: 307 MB at BASELINE C2, ~53 us at the measured
// 5.8 TB/s of a pure write stream); the previous form -- gather kernel, then the tcgen05 transform over one padded
// K block -- spent 155 us there because its TMEM -> register -> shared -> global epilogue was latency-bound and the
// gather was a separate, launch-latency-sized kernel.  With a reduction length of 16 the product is a single
// m16n8k16 MMA step per 16 x 8 outputs, so the accumulators can simply live in registers: every warp owns 16 rows,
//   A fragments  the rows' [aggregate | own features] (16 values, 16-bit), built once per row tile in shared memory
//   B fragments  [W_rel | W_root] as 16-bit, stored in shared memory in fragment order, one conflict-free 8-byte load
//                per lane and MMA; the column order inside a 32-column chunk is permuted so that a lane ends up with 8
//                CONSECUTIVE output features of a row and stores them as one 16-byte word (a warp store = 8 rows x 64 B)
//   C = bias     the bias rides in as the accumulator's initial value
// mma.sync (not tcgen05): TMEM accumulators would have to be drained to registers for the store anyway, which is
// exactly what bounded the old kernel; there is no K loop to pipeline.  Persistent blocks (weights staged once),
// 4 blocks per SM so the dependent gather chains of one block hide under the MMA / store phases of the others.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kMmaThreads = 256;
constexpr int kMmaRows = 128;      // rows a block covers per sweep of its warps: 8 warps x 16 rows

template <typename T> struct Mma16816;
template <> struct Mma16816<__nv_bfloat16> {
  __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct Mma16816<__half> {
  __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};

// four consecutive floats; one 16-byte load when the caller's array is 16-byte aligned (feature matrices carved out of
// a packed staging buffer are only 4-byte aligned)
__device__ __forceinline__ float4 ld4(const float* p, bool vec) {
  if (vec) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// One direction of the layer: destination rows of one side, sources on the other.
struct ConvInSide {
  const int32_t* ptr; const int32_t* idx; const float* val; int32_t rows;
  const float* Xsrc; const float* Xdst; const float* W_rel; const float* b_rel; const float* W_root;
  void* out; void* z16;   // z16: [rows,64] or null
};

// Shared memory: wfrag [N/32][4][32] uint2 | bias [N] float | z tile [128][16] T
// Blocks [0, blocks_a) work on side `sa`, the rest on side `sb` (both directions of the layer in ONE launch: the second
// direction's weight staging and first gathers hide under the first one's stores instead of following its tail).
// (the side's fields are read straight from the kernel parameter bank: selecting them into registers per block costs ~14
// registers of the 64 available at four blocks per SM, so each side gets its own copy of the body instead)
template <typename T>
__device__ __forceinline__ void conv_in_side(const ConvInSide& S, const int side_block, const int side_blocks, const int N,
                                             const int relu) {
  const int32_t* __restrict__ ptr = S.ptr;
  const int32_t* __restrict__ idx = S.idx;
  const float* __restrict__ val = S.val;
  const int32_t rows = S.rows;
  const float* __restrict__ Xsrc = S.Xsrc;
  const float* __restrict__ Xdst = S.Xdst;
  const float* __restrict__ W_rel = S.W_rel;
  const float* __restrict__ b_rel = S.b_rel;
  const float* __restrict__ W_root = S.W_root;
  T* __restrict__ out = reinterpret_cast<T*>(S.out);
  T* __restrict__ z16 = reinterpret_cast<T*>(S.z16);
  extern __shared__ __align__(16) uint8_t smem_m[];
  uint2* wfrag = reinterpret_cast<uint2*>(smem_m);
  float* bias_s = reinterpret_cast<float*>(smem_m + (size_t)N * 32);
  uint32_t* zt = reinterpret_cast<uint32_t*>(smem_m + (size_t)N * 36);      // [128][8] words = 16 T per row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // ---- weights in fragment order: entry (chunk, j, lane) = B fragment of MMA j of the 32-column chunk.
  //      fragment column n (= g) of MMA j is output feature chunk*32 + 8*(n/2) + 2*j + (n%2): lane (g,t) then owns the
  //      accumulators of features chunk*32 + 8t + 2j, +1 -> over j = 0..3 the 8 consecutive features 8t .. 8t+7.
  const bool wvec = ((reinterpret_cast<uintptr_t>(W_rel) | reinterpret_cast<uintptr_t>(W_root)) & 7) == 0;
#pragma unroll 4
  for (int i = tid; i < N * 4; i += kMmaThreads) {
    const int l = i & 31, j = (i >> 5) & 3, chunk = i >> 7;
    const int gg = l >> 2, tt = l & 3;
    const int col = chunk * 32 + 8 * (gg >> 1) + 2 * j + (gg & 1);
    float2 wr, wo;
    if (wvec) {
      wr = __ldg(reinterpret_cast<const float2*>(W_rel + (size_t)col * 8 + 2 * tt));     // k = 2t, 2t+1
      wo = __ldg(reinterpret_cast<const float2*>(W_root + (size_t)col * 8 + 2 * tt));    // k = 2t+8, +9
    } else {
      wr = make_float2(__ldg(W_rel + (size_t)col * 8 + 2 * tt), __ldg(W_rel + (size_t)col * 8 + 2 * tt + 1));
      wo = make_float2(__ldg(W_root + (size_t)col * 8 + 2 * tt), __ldg(W_root + (size_t)col * 8 + 2 * tt + 1));
    }
    wfrag[i] = make_uint2(Half16<T>::pack(wr.x, wr.y), Half16<T>::pack(wo.x, wo.y));
  }
  for (int i = tid; i < N; i += kMmaThreads) bias_s[i] = b_rel ? __ldg(b_rel + i) : 0.f;
  __syncthreads();

  // ---- every WARP walks its own 16-row tiles (no block-level barrier after the weight staging: tiles are small, so the
  //      150K rows of BASELINE C2 spread evenly over all resident warps and a warp's dependent gather chain hides under
  //      the MMA / store phases of the others)
  const int r = lane >> 1, h = lane & 1;             // gather role: row r of the tile, feature half h (4 of the 8)
  const bool vsrc = (reinterpret_cast<uintptr_t>(Xsrc) & 15) == 0, vdst = (reinterpret_cast<uintptr_t>(Xdst) & 15) == 0;
  const int nchunks = N >> 5;
  uint32_t* zw = zt + warp * (16 * 8);               // this warp's z tile: [16][8] words = 16 T per row
  const int64_t warps_total = (int64_t)side_blocks * (kMmaThreads / 32);
  for (int64_t row0 = ((int64_t)side_block * (kMmaThreads / 32) + warp) * 16; row0 < rows; row0 += warps_total * 16) {
    // ---- z[row] = [ sum_e val[e] * Xsrc[idx[e], :] | Xdst[row, :] ]  (fp32 accumulate in CSR order, then 16-bit)
    {
      const int64_t row = row0 + r;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), xd = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows) {
        const int32_t beg = __ldg(ptr + row), end = __ldg(ptr + row + 1);
        int32_t e = beg;
        for (; e + 2 <= end; e += 2) {
          const int32_t i0 = __ldg(idx + e), i1 = __ldg(idx + e + 1);
          const float w0 = __ldg(val + e), w1 = __ldg(val + e + 1);
          const float4 x0 = ld4(Xsrc + (int64_t)i0 * 8 + 4 * h, vsrc);
          const float4 x1 = ld4(Xsrc + (int64_t)i1 * 8 + 4 * h, vsrc);
          a.x = fmaf(w0, x0.x, a.x); a.y = fmaf(w0, x0.y, a.y); a.z = fmaf(w0, x0.z, a.z); a.w = fmaf(w0, x0.w, a.w);
          a.x = fmaf(w1, x1.x, a.x); a.y = fmaf(w1, x1.y, a.y); a.z = fmaf(w1, x1.z, a.z); a.w = fmaf(w1, x1.w, a.w);
        }
        if (e < end) {
          const float w0 = __ldg(val + e);
          const float4 x0 = ld4(Xsrc + (int64_t)__ldg(idx + e) * 8 + 4 * h, vsrc);
          a.x = fmaf(w0, x0.x, a.x); a.y = fmaf(w0, x0.y, a.y); a.z = fmaf(w0, x0.z, a.z); a.w = fmaf(w0, x0.w, a.w);
        }
        xd = ld4(Xdst + row * 8 + 4 * h, vdst);
      }
      uint32_t* zr = zw + r * 8;                     // words: [agg 0..7 | dst 0..7] as 16-bit pairs
      *reinterpret_cast<uint2*>(zr + 2 * h) = make_uint2(Half16<T>::pack(a.x, a.y), Half16<T>::pack(a.z, a.w));
      *reinterpret_cast<uint2*>(zr + 4 + 2 * h) = make_uint2(Half16<T>::pack(xd.x, xd.y), Half16<T>::pack(xd.z, xd.w));
    }
    __syncwarp();
    if (z16) {   // the transform input as the operand of the layer's weight gradient: [z | 1 | 0 ...] 16-bit [rows,64]
      const int64_t row = row0 + r;
      if (row < rows) {
        uint4* dst = reinterpret_cast<uint4*>(z16 + row * 64) + 4 * h;
        if (h == 0) {
          const uint4* src = reinterpret_cast<const uint4*>(zw + r * 8);
          dst[0] = src[0]; dst[1] = src[1];
          dst[2] = make_uint4(Half16<T>::pack(1.f, 0.f), 0u, 0u, 0u);
          dst[3] = make_uint4(0u, 0u, 0u, 0u);
        } else {
          dst[0] = dst[1] = dst[2] = dst[3] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    // ---- transform of the warp's 16 rows
    uint32_t a[4];
    {
      const uint32_t* z0 = zw + g * 8;
      const uint32_t* z1 = z0 + 8 * 8;
      a[0] = z0[t]; a[1] = z1[t]; a[2] = z0[4 + t]; a[3] = z1[4 + t];
    }
    __syncwarp();                                    // the z tile may be rewritten (next tile's gather) from here on
    const int64_t ra = row0 + g, rb = ra + 8;
    T* oa = out + ra * N + 8 * t;
    T* ob = out + rb * N + 8 * t;
    const bool va = ra < rows, vb = rb < rows;
#pragma unroll 2
    for (int c = 0; c < nchunks; ++c) {
      uint32_t lo[4], hi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 b = wfrag[(c * 4 + j) * 32 + lane];
        const float2 bs = *reinterpret_cast<const float2*>(bias_s + c * 32 + 8 * t + 2 * j);
        float d[4] = {bs.x, bs.y, bs.x, bs.y};
        Mma16816<T>::run(d, a, b.x, b.y);
        if (relu) { d[0] = fmaxf(d[0], 0.f); d[1] = fmaxf(d[1], 0.f); d[2] = fmaxf(d[2], 0.f); d[3] = fmaxf(d[3], 0.f); }
        lo[j] = Half16<T>::pack(d[0], d[1]);
        hi[j] = Half16<T>::pack(d[2], d[3]);
      }
      if (va) *reinterpret_cast<uint4*>(oa + c * 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      if (vb) *reinterpret_cast<uint4*>(ob + c * 32) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kMmaThreads, 4)
conv_in_mma_kernel(const __grid_constant__ ConvInSide sa, const __grid_constant__ ConvInSide sb, int blocks_a, int N, int relu) {
  if ((int)blockIdx.x < blocks_a) conv_in_side<T>(sa, (int)blockIdx.x, blocks_a, N, relu);
  else conv_in_side<T>(sb, (int)blockIdx.x - blocks_a, (int)gridDim.x - blocks_a, N, relu);
}

template <typename T>
int launch_mma(const ConvInSide& a, const ConvInSide& b, int N, int relu, cudaStream_t st) {
  static int blocks_per_sm = 0;
  static int smem_set = 0;
  const int smem = N * 36 + kMmaRows * 32;
  auto kern = conv_in_mma_kernel<T>;
  if (smem > smem_set) {
    LPGNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
    blocks_per_sm = 0;
  }
  if (blocks_per_sm == 0) {
    LPGNN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kMmaThreads, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  // blocks needed if every warp took one 16-row tile, capped by the co-resident grid; a pair shares the cap in
  // proportion to the sides' rows (every warp of either side then walks the same number of tiles)
  const int cap = sm_count() * blocks_per_sm;
  int ta = ceil_div(a.rows, kMmaRows), tb = ceil_div(b.rows, kMmaRows);
  if (ta + tb > cap) {
    const int64_t total = (int64_t)a.rows + b.rows;
    int ca = tb == 0 ? cap : (int)((int64_t)cap * a.rows / total);
    if (ta > 0 && ca < 1) ca = 1;
    if (tb > 0 && ca > cap - 1) ca = cap - 1;
    const int cb = cap - ca;
    ta = ta < ca ? ta : ca;
    tb = tb < cb ? tb : cb;
  }
  if (ta + tb == 0) return LPGNN_OK;
  kern<<<ta + tb, kMmaThreads, smem, st>>>(a, b, ta, N, relu);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

int check_side(const ConvInSide& s, const char* who) {
  // (idx / val may be null for a graph without entries: they are only read inside non-empty rows)
  LPGNN_REQUIRE(s.ptr && s.Xsrc && s.Xdst && s.W_rel && s.W_root && s.out, "%s: null pointer", who);
  LPGNN_REQUIRE((uintptr_t)s.Xsrc % 4 == 0 && (uintptr_t)s.Xdst % 4 == 0 && (uintptr_t)s.out % 16 == 0 && (uintptr_t)s.z16 % 16 == 0,
                "%s: misaligned pointer (out / z16 need 16 bytes)", who);
  return LPGNN_OK;
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" int lpgnn_conv_in_16(const int32_t* ptr, const int32_t* idx, const float* val, int32_t rows, const float* Xsrc,
                                const float* Xdst, const float* W_rel, const float* b_rel, const float* W_root, int32_t N,
                                void* out, int out_dtype, int epilogue, void* z16, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && N > 0 && N % 32 == 0 && N <= 4096, "conv_in_16: rows=%d, N=%d (N must be a multiple of 32, <= 4096)", rows, N);
  LPGNN_REQUIRE(is_16bit(out_dtype), "conv_in_16: out dtype %d is not a 16-bit type", out_dtype);
  if (rows == 0) return LPGNN_OK;
  const ConvInSide a{ptr, idx, val, rows, Xsrc, Xdst, W_rel, b_rel, W_root, out, z16};
  if (int rc = check_side(a, "conv_in_16")) return rc;
  const ConvInSide none{};
  const int relu = (epilogue & LPGNN_EPI_RELU) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == LPGNN_F16) return launch_mma<__half>(a, none, N, relu, st);
  return launch_mma<__nv_bfloat16>(a, none, N, relu, st);
}

extern "C" int lpgnn_conv_in_16_pair(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* colptr,
                                     const int32_t* row_csc, const float* val_csc, int32_t m, int32_t n, const float* x_s,
                                     const float* x_t, const float* l2r_wrel, const float* l2r_b, const float* l2r_wroot,
                                     const float* r2l_wrel, const float* r2l_b, const float* r2l_wroot, int32_t N,
                                     void* out_s, void* out_t, int out_dtype, int epilogue, void* z16_s, void* z16_t,
                                     lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m >= 0 && n >= 0 && N > 0 && N % 32 == 0 && N <= 4096,
                "conv_in_16_pair: m=%d, n=%d, N=%d (N must be a multiple of 32, <= 4096)", m, n, N);
  LPGNN_REQUIRE(is_16bit(out_dtype), "conv_in_16_pair: out dtype %d is not a 16-bit type", out_dtype);
  // variables side: destination = variables (CSC view), sources = constraints; constraints side: the CSR view
  ConvInSide t{colptr, row_csc, val_csc, n, x_s, x_t, l2r_wrel, l2r_b, l2r_wroot, out_t, z16_t};
  ConvInSide s{rowptr, col, val, m, x_t, x_s, r2l_wrel, r2l_b, r2l_wroot, out_s, z16_s};
  if (n > 0) { if (int rc = check_side(t, "conv_in_16_pair")) return rc; } else { t = ConvInSide{}; }
  if (m > 0) { if (int rc = check_side(s, "conv_in_16_pair")) return rc; } else { s = ConvInSide{}; }
  const int relu = (epilogue & LPGNN_EPI_RELU) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == LPGNN_F16) return launch_mma<__half>(t, s, N, relu, st);
  return launch_mma<__nv_bfloat16>(t, s, N, relu, st);
}
