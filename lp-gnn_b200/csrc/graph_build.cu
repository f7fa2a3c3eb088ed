// (a1) COO -> canonical CSR + CSC view of the LP matrix A.
//
// Replaces torch_sparse.SparseTensor.from_edge_index (reference dataset.py:301-304: storage
// sorted by row*ncols+col) and SparseTensor.t() (reference arch.py:71: csr2csc =
// argsort(col*nrows+row)).  All integer work; bit-exact with the reference ordering.
//
// Method: stable LSD radix sort (8-bit digits) of (key, payload) pairs, hand-written:
//   histogram kernel  -> per-block digit counts, laid out digit-major
//   scan kernel       -> exclusive scan over [256 x nblocks]  (global digit offsets per block)
//   scatter kernel    -> stable scatter; the rank of an item inside its digit is computed
//                        with warp match + an in-order cross-warp prefix, no atomics on
//                        ordered data, so the result is deterministic.
// CSR  = sort by column digits, then by row digits (LSD over the composite key).
// CSC  = stable sort of the CSR entries by column only (CSR order already breaks ties by row).
// HBM-bound integer work; roofline = bytes moved per pass (12 B in + 12 B out per entry... see DESIGN.md).
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;
constexpr int kSortTile = kSortThreads * kSortRounds;  // items per block
constexpr int kRadix = 256;

__global__ void iota_kernel(uint32_t* out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}

__global__ void narrow_i64_kernel(const int64_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)in[i];
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx,
                                  uint32_t* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

__global__ void gather_f32_kernel(const float* __restrict__ src, const uint32_t* __restrict__ idx,
                                  float* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ counts,
                  int nblocks) {
  __shared__ uint32_t h[kRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (kRadix - 1)], 1u);  // integer counts: order-free
  }
  __syncthreads();
  counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `counts` (length len) in place; single block.
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* __restrict__ counts, int64_t len) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const int64_t per = (len + 1023) / 1024;
  const int64_t lo = t * per, hi = min(lo + per, len);
  uint32_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += counts[i];
  part[t] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t v = (t >= off) ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = (t == 0) ? 0 : part[t - 1];
  for (int64_t i = lo; i < hi; ++i) {
    uint32_t c = counts[i];
    counts[i] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                     const uint32_t* __restrict__ offsets, int nblocks) {
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t goff[kRadix];          // global start of (digit, this block)
  __shared__ uint32_t run[kRadix];           // items of this digit already placed by earlier rounds
  __shared__ uint32_t wcnt[kWarps][kRadix];  // per-warp digit counts of the current round
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  goff[t] = offsets[(size_t)t * nblocks + blockIdx.x];
  run[t] = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) wcnt[w][t] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + r * kSortThreads + t;
    const bool valid = i < n;
    uint32_t key = 0, val = 0;
    if (valid) { key = keys_in[i]; val = vals_in[i]; }
    const uint32_t digit = valid ? ((key >> shift) & (kRadix - 1)) : kRadix;  // invalid lanes: own class
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1));
    if (valid && rank_in_warp == 0) wcnt[warp][digit] = __popc(peers);
    __syncthreads();
    uint32_t pos = 0;
    if (valid) {
      uint32_t before = run[digit];
      for (int w = 0; w < warp; ++w) before += wcnt[w][digit];
      pos = goff[digit] + before + rank_in_warp;
    }
    __syncthreads();
    {
      uint32_t s = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) { s += wcnt[w][t]; wcnt[w][t] = 0; }
      run[t] += s;
    }
    __syncthreads();
    if (valid) { keys_out[pos] = key; vals_out[pos] = val; }
  }
}

// ptr[q] = first position e with sorted_keys[e] >= q, for q in [0, rows]; ptr[rows] = n.
__global__ void fill_ptr_kernel(const uint32_t* __restrict__ sorted_keys, int64_t n, int32_t rows,
                                int32_t* __restrict__ ptr) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > n) return;
  int64_t prev = (e == 0) ? -1 : (int64_t)sorted_keys[e - 1];
  int64_t cur = (e == n) ? (int64_t)rows : (int64_t)sorted_keys[e];
  for (int64_t q = prev + 1; q <= cur; ++q) ptr[q] = (int32_t)e;
}

int bits_for(int64_t extent) {  // bits needed for values in [0, extent)
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < extent) ++b;
  return b;
}

struct SortBufs {
  uint32_t *k[2], *v[2];
  uint32_t* counts;
};

// Sorts (k[0], v[0]) by key bits [0, bits); returns the index (0/1) of the buffer holding the result.
int radix_sort(SortBufs& b, int64_t n, int bits, cudaStream_t st) {
  const int nblocks = ceil_div(n, kSortTile);
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(b.k[cur], n, shift, b.counts, nblocks);
    scan_kernel<<<1, 1024, 0, st>>>(b.counts, (int64_t)kRadix * nblocks);
    radix_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(b.k[cur], b.v[cur], b.k[cur ^ 1], b.v[cur ^ 1], n,
                                                           shift, b.counts, nblocks);
    cur ^= 1;
    count_launches(3);
  }
  return cur;
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" size_t lpgnn_graph_build_workspace_bytes(int64_t nnz, int32_t m, int32_t n) {
  (void)m; (void)n;
  const size_t z = (size_t)(nnz > 0 ? nnz : 1);
  const size_t words = align_up(z, 64);
  const size_t nblocks = (z + kSortTile - 1) / kSortTile;
  // k0,k1,v0,v1,r32,c32,rows_sorted + counts
  return (7 * words + align_up(kRadix * nblocks, 64)) * sizeof(uint32_t) + 256;
}

extern "C" int lpgnn_graph_build(const void* coo_row, const void* coo_col, int idx_is_i64, const float* coo_val,
                                 int64_t nnz, int32_t m, int32_t n, int32_t* rowptr, int32_t* col, float* val,
                                 int32_t* colptr, int32_t* row_csc, float* val_csc, int32_t* csr2csc,
                                 void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m >= 0 && n >= 0 && nnz >= 0, "graph_build: negative size");
  LPGNN_REQUIRE(nnz < ((int64_t)1 << 31), "graph_build: nnz must be < 2^31");
  LPGNN_REQUIRE(rowptr && colptr, "graph_build: null output");
  if (workspace_bytes < lpgnn_graph_build_workspace_bytes(nnz, m, n)) {
    set_error("graph_build: workspace %zu < required %zu", workspace_bytes,
              lpgnn_graph_build_workspace_bytes(nnz, m, n));
    return LPGNN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (nnz == 0) {
    LPGNN_CUDA_OK(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * ((size_t)m + 1), st));
    LPGNN_CUDA_OK(cudaMemsetAsync(colptr, 0, sizeof(int32_t) * ((size_t)n + 1), st));
    return LPGNN_OK;
  }
  LPGNN_REQUIRE(coo_row && coo_col && coo_val && col && val && row_csc && val_csc && csr2csc && workspace,
                "graph_build: null pointer");
  const int64_t z = nnz;
  const size_t words = align_up((size_t)z, 64);
  uint32_t* w = reinterpret_cast<uint32_t*>(workspace);
  SortBufs b;
  b.k[0] = w; b.k[1] = w + words; b.v[0] = w + 2 * words; b.v[1] = w + 3 * words;
  uint32_t* r32 = w + 4 * words;
  uint32_t* c32 = w + 5 * words;
  uint32_t* rows_sorted = w + 6 * words;
  b.counts = w + 7 * words;
  const int tb = 256, gb = ceil_div(z, tb);

  const uint32_t *rsrc, *csrc;
  if (idx_is_i64) {
    narrow_i64_kernel<<<gb, tb, 0, st>>>(reinterpret_cast<const int64_t*>(coo_row), r32, z);
    narrow_i64_kernel<<<gb, tb, 0, st>>>(reinterpret_cast<const int64_t*>(coo_col), c32, z);
    rsrc = r32; csrc = c32;
  } else {
    rsrc = reinterpret_cast<const uint32_t*>(coo_row);
    csrc = reinterpret_cast<const uint32_t*>(coo_col);
  }
  // ---- CSR: LSD over (row, col): column digits first, then row digits
  LPGNN_CUDA_OK(cudaMemcpyAsync(b.k[0], csrc, sizeof(uint32_t) * z, cudaMemcpyDeviceToDevice, st));
  iota_kernel<<<gb, tb, 0, st>>>(b.v[0], z);
  int cur = radix_sort(b, z, bits_for(n), st);
  if (cur != 0) std::swap(b.k[0], b.k[1]), std::swap(b.v[0], b.v[1]);
  gather_u32_kernel<<<gb, tb, 0, st>>>(rsrc, b.v[0], b.k[0], z);  // keys := row of each (col-sorted) entry
  cur = radix_sort(b, z, bits_for(m), st);
  // b.k[cur] = rows of the CSR entries (sorted), b.v[cur] = original COO index of each CSR entry
  LPGNN_CUDA_OK(cudaMemcpyAsync(rows_sorted, b.k[cur], sizeof(uint32_t) * z, cudaMemcpyDeviceToDevice, st));
  gather_u32_kernel<<<gb, tb, 0, st>>>(csrc, b.v[cur], reinterpret_cast<uint32_t*>(col), z);
  gather_f32_kernel<<<gb, tb, 0, st>>>(coo_val, b.v[cur], val, z);
  fill_ptr_kernel<<<ceil_div(z + 1, tb), tb, 0, st>>>(rows_sorted, z, m, rowptr);
  // ---- CSC view: stable sort of CSR entries by column
  LPGNN_CUDA_OK(cudaMemcpyAsync(b.k[0], col, sizeof(uint32_t) * z, cudaMemcpyDeviceToDevice, st));
  iota_kernel<<<gb, tb, 0, st>>>(b.v[0], z);
  cur = radix_sort(b, z, bits_for(n), st);
  LPGNN_CUDA_OK(cudaMemcpyAsync(csr2csc, b.v[cur], sizeof(uint32_t) * z, cudaMemcpyDeviceToDevice, st));
  gather_u32_kernel<<<gb, tb, 0, st>>>(rows_sorted, b.v[cur], reinterpret_cast<uint32_t*>(row_csc), z);
  gather_f32_kernel<<<gb, tb, 0, st>>>(val, b.v[cur], val_csc, z);
  fill_ptr_kernel<<<ceil_div(z + 1, tb), tb, 0, st>>>(b.k[cur], z, n, colptr);
  LPGNN_LAUNCH_OK();
  count_launches(idx_is_i64 ? 12 : 10);  // + 3 per radix pass, counted in radix_sort
  return LPGNN_OK;
}
