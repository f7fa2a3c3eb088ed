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
//   B fragments  [W_rel | W_root] as 16-bit, stored in shared memory in fragment order, one conflict-free 16-byte load
//                per lane and PAIR of MMAs; the column order inside a 32-column chunk is permuted so that a lane ends
//                up with 8 CONSECUTIVE output features of a row and stores them as one 16-byte word (a warp store =
//                8 rows x 64 B)
//   C = bias     the bias rides in as the accumulator's initial value
// mma.sync (not tcgen05): TMEM accumulators would have to be drained to registers for the store anyway, which is
// exactly what bounded the old kernel; there is no K loop to pipeline.
// Two kernels: conv_in_mma_kernel (every warp walks its own 16-row tiles, B fragments re-read from shared memory per
// chunk; persistent blocks, 4 per SM: small inputs and any N % 32 == 0) and conv_in_mma_regb_kernel further down
// (large inputs, N = 256 .. 1024: B fragments in registers, producer warps gathering ahead).  Same arithmetic in the same
// order, so their results are bit-identical.
#include <stdlib.h>

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
    // (the fragments of MMAs 2jj and 2jj+1 of a lane sit side by side: one 16-byte load per lane fetches both)
    wfrag[(((chunk * 2 + (j >> 1)) * 32 + l) << 1) + (j & 1)] = make_uint2(Half16<T>::pack(wr.x, wr.y), Half16<T>::pack(wo.x, wo.y));
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
        for (; e + 4 <= end; e += 4) {     // four entries' dependent loads (idx -> X row) in flight per round trip
          const int32_t i0 = __ldg(idx + e), i1 = __ldg(idx + e + 1), i2 = __ldg(idx + e + 2), i3 = __ldg(idx + e + 3);
          const float w0 = __ldg(val + e), w1 = __ldg(val + e + 1), w2 = __ldg(val + e + 2), w3 = __ldg(val + e + 3);
          const float4 x0 = ld4(Xsrc + (int64_t)i0 * 8 + 4 * h, vsrc);
          const float4 x1 = ld4(Xsrc + (int64_t)i1 * 8 + 4 * h, vsrc);
          const float4 x2 = ld4(Xsrc + (int64_t)i2 * 8 + 4 * h, vsrc);
          const float4 x3 = ld4(Xsrc + (int64_t)i3 * 8 + 4 * h, vsrc);
          a.x = fmaf(w0, x0.x, a.x); a.y = fmaf(w0, x0.y, a.y); a.z = fmaf(w0, x0.z, a.z); a.w = fmaf(w0, x0.w, a.w);
          a.x = fmaf(w1, x1.x, a.x); a.y = fmaf(w1, x1.y, a.y); a.z = fmaf(w1, x1.z, a.z); a.w = fmaf(w1, x1.w, a.w);
          a.x = fmaf(w2, x2.x, a.x); a.y = fmaf(w2, x2.y, a.y); a.z = fmaf(w2, x2.z, a.z); a.w = fmaf(w2, x2.w, a.w);
          a.x = fmaf(w3, x3.x, a.x); a.y = fmaf(w3, x3.y, a.y); a.z = fmaf(w3, x3.z, a.z); a.w = fmaf(w3, x3.w, a.w);
        }
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
      const float4* bp = reinterpret_cast<const float4*>(bias_s + c * 32 + 8 * t);
      const float4 bq[2] = {bp[0], bp[1]};                 // this lane's 8 consecutive features of the chunk
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const uint4 b = reinterpret_cast<const uint4*>(wfrag)[(c * 2 + jj) * 32 + lane];
        float d0[4] = {bq[jj].x, bq[jj].y, bq[jj].x, bq[jj].y};
        float d1[4] = {bq[jj].z, bq[jj].w, bq[jj].z, bq[jj].w};
        Mma16816<T>::run(d0, a, b.x, b.y);
        Mma16816<T>::run(d1, a, b.z, b.w);
        if (relu) {
          d0[0] = fmaxf(d0[0], 0.f); d0[1] = fmaxf(d0[1], 0.f); d0[2] = fmaxf(d0[2], 0.f); d0[3] = fmaxf(d0[3], 0.f);
          d1[0] = fmaxf(d1[0], 0.f); d1[1] = fmaxf(d1[1], 0.f); d1[2] = fmaxf(d1[2], 0.f); d1[3] = fmaxf(d1[3], 0.f);
        }
        lo[2 * jj] = Half16<T>::pack(d0[0], d0[1]);     hi[2 * jj] = Half16<T>::pack(d0[2], d0[3]);
        lo[2 * jj + 1] = Half16<T>::pack(d1[0], d1[1]); hi[2 * jj + 1] = Half16<T>::pack(d1[2], d1[3]);
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

// ---------------------------------------------------------------------------------------------------------------------
// Large inputs: warp-specialised form with the B fragments in REGISTERS.
// Measured on B200 at BASELINE C2 (both sides, 78 us as one launch of the kernel above): without its stores that kernel
// still takes 45 us, without its MMAs 76 us -- per 32-column chunk every warp re-reads 1 KB of B fragments and 128 B of
// bias from shared memory for four MMAs, and those reads, not the tensor work or the store pattern, are what its chunk
// loop waits on; the gather chains of all warps also run in phase at the start.  Here
//   warps 0-7   consumers: warp w keeps the B fragments of columns [w*N/8, (w+1)*N/8) in registers for the whole side
//               (N = 256*CW: CW*8 registers; the bias slice, broadcast reads, stays in shared memory), and per 16-row tile reads only the A fragments (64 B per lane
//               group) from shared memory: CW*4 MMAs -> ReLU -> 16-bit -> CW*2 stores of 8 rows x 64 B
//   warps 8-15  producers, two groups of four warps taking alternate 128-row batches: one lane per row gathers
//               [A.x_src | x_dst] (fp32, CSR order) into a four-deep ring of z tiles (and writes the z16 operand of the
//               weight gradient); a row's dependent chain (ptr -> idx -> X row, cold in L2) is longer than the time the
//               consumers need to store a batch, hence two groups in flight
// Named barriers (full / empty per buffer: the consumers + the buffer's producer group) hand the batches over, so the
// gather loads of batches k+1, k+2 run under the stores of batch k.  One persistent block per SM; both directions of the layer are walked by every block.
constexpr int kRegBConsumers = 8, kRegBGroupWarps = 4, kRegBGroups = 2, kRegBBufs = 4;
constexpr int kRegBThreads = (kRegBConsumers + kRegBGroups * kRegBGroupWarps) * 32;   // 512
constexpr int kRegBBatch = kRegBGroupWarps * 32;                          // 128 rows: one producer lane per row
constexpr int kRegBBarThreads = (kRegBConsumers + kRegBGroupWarps) * 32;  // a buffer's barriers: the consumers + ONE producer group

__device__ __forceinline__ void nbar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kRegBBarThreads) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kRegBBarThreads) : "memory"); }

template <typename T, int CW>
__global__ void __launch_bounds__(kRegBThreads, 1)
conv_in_mma_regb_kernel(const __grid_constant__ ConvInSide sa, const __grid_constant__ ConvInSide sb, int relu) {
  constexpr int N = CW * 256;
  __shared__ __align__(16) uint32_t zt[kRegBBufs][kRegBBatch * 8];       // a ring of batches of z rows, 16 T per row
  __shared__ __align__(16) float bias_s[2][N];                           // per side (the consumers of side 1 may start while
                                                                         // a slower warp still works on side 0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int k = 0;                                                             // batches this block has handled (both roles count alike)
  if (warp >= kRegBConsumers) {
    // the two producer warpgroups hand registers to the two consumer warpgroups (128 per thread at launch: 96 / 160)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    const int group = (warp - kRegBConsumers) / kRegBGroupWarps;         // group g gathers the block's batches k = g, g + 2, ...
    for (int side = 0; side < 2; ++side) {
      const ConvInSide& S = side ? sb : sa;
      const int32_t rows = S.rows;
      const int nb = (rows + kRegBBatch - 1) / kRegBBatch;
      // ------------------------------------------------------------------------------------------ producers
      const int32_t* __restrict__ ptr = S.ptr;
      const int32_t* __restrict__ idx = S.idx;
      const float* __restrict__ val = S.val;
      const float* __restrict__ Xsrc = S.Xsrc;
      const float* __restrict__ Xdst = S.Xdst;
      T* __restrict__ z16 = reinterpret_cast<T*>(S.z16);
      const bool vsrc = (reinterpret_cast<uintptr_t>(Xsrc) & 15) == 0, vdst = (reinterpret_cast<uintptr_t>(Xdst) & 15) == 0;
      const int rl = ((warp - kRegBConsumers) % kRegBGroupWarps) * 32 + lane;   // row of the batch
      // This group's batches of the side: ordinals i with (k + i) odd / even = group.  The row pointers and the first
      // (idx, val) round of the group's NEXT batch are fetched while the current batch is finished and handed over, so a
      // batch's dependent chain is just its X-row rounds: ceil(nnz / 4) round trips for the longest row of a warp.
      const int first = (group - k) & 1;
      const int64_t stride = 2 * (int64_t)gridDim.x;
      int32_t beg = 0, end = 0;
      int32_t ci[4];
      float cw[4];
      {
        const int64_t row = ((int64_t)blockIdx.x + (int64_t)first * gridDim.x) * kRegBBatch + rl;
        if (row < rows) { beg = __ldg(ptr + row); end = __ldg(ptr + row + 1); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ci[u] = (beg + u < end) ? __ldg(idx + beg + u) : 0;
          cw[u] = (beg + u < end) ? __ldg(val + beg + u) : 0.f;
        }
      }
      int kk = k + first;
      for (int64_t b = (int64_t)blockIdx.x + (int64_t)first * gridDim.x; b < nb; b += stride, kk += 2) {
        const int buf = kk % kRegBBufs;
        const int64_t row = b * kRegBBatch + rl;
        const int64_t nrow = (b + stride) * kRegBBatch + rl;             // this lane's row in the group's next batch
        int32_t nbeg = 0, nend = 0;
        if (nrow < rows) { nbeg = __ldg(ptr + nrow); nend = __ldg(ptr + nrow + 1); }
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, d0 = a0, d1 = a0;
        if (row < rows) {
          d0 = ld4(Xdst + row * 8, vdst);
          d1 = ld4(Xdst + row * 8 + 4, vdst);
          // up to four entries per round, accumulated in CSR order; the next round's (idx, val) are fetched while this
          // round's X rows are in flight
          for (int32_t e = beg; e < end; e += 4) {
            float4 xa[4], xb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (e + u < end) { xa[u] = ld4(Xsrc + (int64_t)ci[u] * 8, vsrc); xb[u] = ld4(Xsrc + (int64_t)ci[u] * 8 + 4, vsrc); }
            float w_now[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              w_now[u] = cw[u];
              const int32_t nx = e + 4 + u;
              ci[u] = (nx < end) ? __ldg(idx + nx) : 0;
              cw[u] = (nx < end) ? __ldg(val + nx) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (e + u < end) {
                const float cf = w_now[u];
                a0.x = fmaf(cf, xa[u].x, a0.x); a0.y = fmaf(cf, xa[u].y, a0.y); a0.z = fmaf(cf, xa[u].z, a0.z); a0.w = fmaf(cf, xa[u].w, a0.w);
                a1.x = fmaf(cf, xb[u].x, a1.x); a1.y = fmaf(cf, xb[u].y, a1.y); a1.z = fmaf(cf, xb[u].z, a1.z); a1.w = fmaf(cf, xb[u].w, a1.w);
              }
          }
        }
        beg = nbeg; end = nend;                                          // first round of the next batch: in flight from here on
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ci[u] = (beg + u < end) ? __ldg(idx + beg + u) : 0;
          cw[u] = (beg + u < end) ? __ldg(val + beg + u) : 0.f;
        }
        const uint4 za = make_uint4(Half16<T>::pack(a0.x, a0.y), Half16<T>::pack(a0.z, a0.w), Half16<T>::pack(a1.x, a1.y),
                                    Half16<T>::pack(a1.z, a1.w));
        const uint4 zd = make_uint4(Half16<T>::pack(d0.x, d0.y), Half16<T>::pack(d0.z, d0.w), Half16<T>::pack(d1.x, d1.y),
                                    Half16<T>::pack(d1.z, d1.w));
        if (kk >= kRegBBufs) nbar_sync(1 + kRegBBufs + buf);             // the consumers have read batch kk-4 out of this buffer
        uint4* zr = reinterpret_cast<uint4*>(&zt[buf][rl * 8]);
        zr[0] = za; zr[1] = zd;
        nbar_arrive(1 + buf);                                            // batch kk is in shared memory
        if (z16 && row < rows) {   // [z | 1 | 0 ...] 16-bit [rows,64]: the operand of the layer's weight gradient
          uint4* dst = reinterpret_cast<uint4*>(z16 + row * 64);
          dst[0] = za; dst[1] = zd;
          dst[2] = make_uint4(Half16<T>::pack(1.f, 0.f), 0u, 0u, 0u);
          dst[3] = dst[4] = dst[5] = dst[6] = dst[7] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      k += nb > (int)blockIdx.x ? (nb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;   // batches of the side, both groups
    }
    // take the consumers' last releases, so that no barrier is left half-arrived when the block exits
    for (int j = (k >= kRegBBufs ? k - kRegBBufs : 0); j < k; ++j)
      if ((j & 1) == group) nbar_sync(1 + kRegBBufs + j % kRegBBufs);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    for (int side = 0; side < 2; ++side) {
      const ConvInSide& S = side ? sb : sa;
      const int32_t rows = S.rows;
      const int nb = (rows + kRegBBatch - 1) / kRegBBatch;
      // ------------------------------------------------------------------------------------------ consumers
      const float* __restrict__ W_rel = S.W_rel;
      const float* __restrict__ W_root = S.W_root;
      const float* __restrict__ b_rel = S.b_rel;
      T* __restrict__ out = reinterpret_cast<T*>(S.out);
      const int g = lane >> 2, t = lane & 3;
      // B fragments of this warp's CW chunks in registers (same fragment / column permutation as conv_in_side above);
      // the bias slice of the warp's columns in shared memory (only this warp reads it: __syncwarp suffices)
      uint2 breg[CW][4];
      float* bias_w = bias_s[side] + warp * (CW * 32);
      if (nb > 0) {
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
          const int chunk = warp * CW + cc;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = chunk * 32 + 8 * (g >> 1) + 2 * j + (g & 1);
            const float wr0 = __ldg(W_rel + (size_t)col * 8 + 2 * t), wr1 = __ldg(W_rel + (size_t)col * 8 + 2 * t + 1);
            const float wo0 = __ldg(W_root + (size_t)col * 8 + 2 * t), wo1 = __ldg(W_root + (size_t)col * 8 + 2 * t + 1);
            breg[cc][j] = make_uint2(Half16<T>::pack(wr0, wr1), Half16<T>::pack(wo0, wo1));
          }
          bias_w[cc * 32 + lane] = b_rel ? __ldg(b_rel + chunk * 32 + lane) : 0.f;
        }
        __syncwarp();
      }
      for (int b = blockIdx.x; b < nb; b += gridDim.x, ++k) {
        const int buf = k % kRegBBufs;
        nbar_sync(1 + buf);                                              // batch k has been gathered
        const uint32_t* zb = zt[buf];
#pragma unroll 1
        for (int tile = 0; tile < kRegBBatch / 16; ++tile) {
          uint32_t a[4];
          const uint32_t* z0 = zb + (tile * 16 + g) * 8;
          const uint32_t* z1 = z0 + 8 * 8;
          a[0] = z0[t]; a[1] = z1[t]; a[2] = z0[4 + t]; a[3] = z1[4 + t];
          // (barrier instructions order the prior shared-memory reads of the arriving threads before the barrier's
          // completion, PTX ISA "barrier": the producers may overwrite the buffer once all consumer warps have arrived)
          if (tile == kRegBBatch / 16 - 1) nbar_arrive(1 + kRegBBufs + buf);
          if ((int64_t)b * kRegBBatch + tile * 16 >= rows) continue;    // warp-uniform: whole tile beyond the last row
          const int64_t ra = (int64_t)b * kRegBBatch + tile * 16 + g, rb = ra + 8;
          T* oa = out + ra * N + (size_t)warp * (CW * 32) + 8 * t;
          T* ob = out + rb * N + (size_t)warp * (CW * 32) + 8 * t;
          const bool va = ra < rows, vb = rb < rows;
#pragma unroll
          for (int cc = 0; cc < CW; ++cc) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 bs = *reinterpret_cast<const float2*>(bias_w + cc * 32 + 8 * t + 2 * j);
              float d[4] = {bs.x, bs.y, bs.x, bs.y};
              Mma16816<T>::run(d, a, breg[cc][j].x, breg[cc][j].y);
              if (relu) { d[0] = fmaxf(d[0], 0.f); d[1] = fmaxf(d[1], 0.f); d[2] = fmaxf(d[2], 0.f); d[3] = fmaxf(d[3], 0.f); }
              lo[j] = Half16<T>::pack(d[0], d[1]);
              hi[j] = Half16<T>::pack(d[2], d[3]);
            }
            if (va) *reinterpret_cast<uint4*>(oa + cc * 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            if (vb) *reinterpret_cast<uint4*>(ob + cc * 32) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          }
        }
      }
    }
  }
}

// 0 keeps every size on the shared-memory-B kernel (environment LPGNN_CONV_IN_REGB=0 / lpgnn_set_conv_in_regb: A/B runs)
int g_conv_in_regb = [] { const char* e = getenv("LPGNN_CONV_IN_REGB"); return e ? atoi(e) != 0 : 1; }();

template <typename T, int CW>
int launch_regb(const ConvInSide& a, const ConvInSide& b, int relu, int batches, cudaStream_t st) {
  const int grid = batches < sm_count() ? batches : sm_count();
  conv_in_mma_regb_kernel<T, CW><<<grid, kRegBThreads, 0, st>>>(a, b, relu);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

template <typename T>
int launch_mma(const ConvInSide& a, const ConvInSide& b, int N, int relu, cudaStream_t st) {
  // large inputs with N = 256 * {1..4}: register-B form (one persistent block per SM, 128-row batches)
  const int batches = ceil_div(a.rows, kRegBBatch) + ceil_div(b.rows, kRegBBatch);
  if (g_conv_in_regb && N % 256 == 0 && N <= 1024 && batches >= 2 * sm_count()) {
    switch (N / 256) {
      case 1: return launch_regb<T, 1>(a, b, relu, batches, st);
      case 2: return launch_regb<T, 2>(a, b, relu, batches, st);
      case 3: return launch_regb<T, 3>(a, b, relu, batches, st);
      default: return launch_regb<T, 4>(a, b, relu, batches, st);
    }
  }
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

extern "C" int lpgnn_set_conv_in_regb(int enable) {
  const int prev = lpgnn::g_conv_in_regb;
  lpgnn::g_conv_in_regb = enable ? 1 : 0;
  return prev;
}
