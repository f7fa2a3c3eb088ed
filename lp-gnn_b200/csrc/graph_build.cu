// (a1) COO -> canonical CSR + CSC view of the LP matrix A.
//
// Replaces torch_sparse.SparseTensor.from_edge_index (reference dataset.py:301-304: storage
// sorted by row*ncols+col) and SparseTensor.t() (reference arch.py:71: csr2csc =
// argsort(col*nrows+row)).  All integer work; bit-exact with the reference ordering.
//
// Method: stable LSD radix sort (up to 9-bit digits) of (key, payload) pairs, hand-written:
//   histogram kernel  -> per-block digit counts, laid out digit-major
//   scan kernel       -> one block per digit: exclusive scan over the tile counts + digit total
//   scatter kernel    -> stable scatter; the rank of an item inside its digit is computed
//                        with warp match + an in-order cross-warp prefix, no atomics on
//                        ordered data, so the result is deterministic.
// CSR  = sort by column digits, then by row digits (LSD over the composite key).
// CSC  = stable sort of the CSR entries by column only (CSR order already breaks ties by row).
// HBM-bound integer work; roofline = bytes moved per pass (12 B in + 12 B out per entry... see DESIGN.md).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortRounds = 8;                          // items per thread, held in registers
constexpr int kSortTile = kSortThreads * kSortRounds;   // items per block
constexpr int kMaxRadixBits = 9;
constexpr int kMaxRadix = 1 << kMaxRadixBits;

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

// CSR payload gather: col[e] = coo_col[perm[e]], val[e] = coo_val[perm[e]]
__global__ void gather_entry_kernel(const uint32_t* __restrict__ col_src, const float* __restrict__ val_src,
                                    const uint32_t* __restrict__ perm, uint32_t* __restrict__ col_out,
                                    float* __restrict__ val_out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t p = perm[i]; col_out[i] = col_src[p]; val_out[i] = val_src[p]; }
}

// Sorted-COO fast path in ONE pass over the entries: verify the (row, col) order and the index range, emit the
// CSR payload (col, val), fill rowptr from the row boundaries and set up the (key = col, payload = position)
// pairs of the CSC sort.
__device__ __forceinline__ void prep_sorted_entry(int64_t e, const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
                                                  const float* __restrict__ val, int64_t n, uint32_t m, uint32_t ncols,
                                                  uint32_t* __restrict__ col_out, float* __restrict__ val_out,
                                                  int32_t* __restrict__ rowptr, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ vals, uint32_t* __restrict__ status) {
  if (e > n) return;
  const int64_t prev = (e == 0) ? -1 : (int64_t)row[e - 1];
  int64_t cur = (int64_t)m;
  if (e < n) {
    const uint32_t r = row[e], c = col[e];
    cur = r;
    col_out[e] = c;
    val_out[e] = val[e];
    keys[e] = c;
    vals[e] = (uint32_t)e;
    if (status) {
      if (r >= m || c >= ncols) atomicOr(status, 2u);
      if (e > 0 && (r < (uint32_t)prev || (r == (uint32_t)prev && c < col[e - 1]))) atomicOr(status, 1u);
    }
    if (r >= m) cur = prev;  // out-of-range row: do not write past rowptr
  }
  for (int64_t q = prev + 1; q <= cur && q <= (int64_t)m; ++q) rowptr[q] = (int32_t)e;
}

__global__ void prep_sorted_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
                                   const float* __restrict__ val, int64_t n, uint32_t m, uint32_t ncols,
                                   uint32_t* __restrict__ col_out, float* __restrict__ val_out,
                                   int32_t* __restrict__ rowptr, uint32_t* __restrict__ keys,
                                   uint32_t* __restrict__ vals, uint32_t* __restrict__ status) {
  prep_sorted_entry(blockIdx.x * (int64_t)blockDim.x + threadIdx.x, row, col, val, n, m, ncols, col_out, val_out, rowptr, keys,
                    vals, status);
}

// CSC payload + colptr in one pass: csr2csc[k] = perm[k], row_csc[k] = rows[perm[k]], val_csc[k] = val[perm[k]],
// colptr from the boundaries of the sorted column keys.
__device__ __forceinline__ void finish_csc_entry(int64_t k, const uint32_t* __restrict__ rows, const float* __restrict__ val,
                                                 const uint32_t* __restrict__ sorted_cols, const uint32_t* __restrict__ perm,
                                                 int64_t n, int32_t ncols, uint32_t* __restrict__ csr2csc,
                                                 uint32_t* __restrict__ row_csc, float* __restrict__ val_csc,
                                                 int32_t* __restrict__ colptr) {
  if (k > n) return;
  const int64_t prev = (k == 0) ? -1 : (int64_t)sorted_cols[k - 1];
  int64_t cur = ncols;
  if (k < n) {
    const uint32_t p = perm[k];
    csr2csc[k] = p;
    row_csc[k] = rows[p];
    val_csc[k] = val[p];
    cur = min((int64_t)sorted_cols[k], (int64_t)ncols);
  }
  for (int64_t q = prev + 1; q <= cur; ++q) colptr[q] = (int32_t)k;
}

__global__ void finish_csc_kernel(const uint32_t* __restrict__ rows, const float* __restrict__ val,
                                  const uint32_t* __restrict__ sorted_cols, const uint32_t* __restrict__ perm,
                                  int64_t n, int32_t ncols, uint32_t* __restrict__ csr2csc,
                                  uint32_t* __restrict__ row_csc, float* __restrict__ val_csc,
                                  int32_t* __restrict__ colptr) {
  finish_csc_entry(blockIdx.x * (int64_t)blockDim.x + threadIdx.x, rows, val, sorted_cols, perm, n, ncols, csr2csc, row_csc,
                   val_csc, colptr);
}

// plain copy as a kernel (a cudaMemcpyAsync D2D costs ~10x a small kernel on the launch path)
__global__ void copy_u32_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// keys[i] = key_src[i], vals[i] = i   (start of a sort)
__global__ void init_pairs_kernel(const uint32_t* __restrict__ key_src, uint32_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = key_src[i]; vals[i] = (uint32_t)i; }
}

__global__ void expand_check_range_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col, int64_t n,
                                          uint32_t m, uint32_t ncols, uint32_t* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && (row[i] >= m || col[i] >= ncols)) atomicOr(flag, 2u);
}

__device__ __forceinline__ void radix_hist_tile(int bid, const uint32_t* __restrict__ keys, int64_t n, int shift, int radix_bits,
                                                uint32_t* __restrict__ counts, int nblocks) {
  __shared__ uint32_t h[kMaxRadix];
  const int radix = 1 << radix_bits;
  __syncthreads();                       // (a previous tile of the same block may still be reading h)
  for (int d = threadIdx.x; d < radix; d += kSortThreads) h[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)bid * kSortTile;
  uint32_t k[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + r * kSortThreads + threadIdx.x;
    k[r] = (i < n) ? keys[i] : 0xffffffffu;
  }
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(k[r] >> shift) & (radix - 1)], 1u);  // integer counts: order-free
  }
  __syncthreads();
  for (int d = threadIdx.x; d < radix; d += kSortThreads) counts[(size_t)d * nblocks + bid] = h[d];
}

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int radix_bits,
                  uint32_t* __restrict__ counts, int nblocks) {
  radix_hist_tile(blockIdx.x, keys, n, shift, radix_bits, counts, nblocks);
}

// One block per digit d: exclusive scan of counts[d][0..nblocks) in place (coalesced) and the digit's
// total into totals[d].  The scatter kernel turns the totals into digit bases itself.
__device__ __forceinline__ void scan_digit_row(int digit, uint32_t* __restrict__ counts, int nblocks, uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_sum[kSortThreads / 32];
  __shared__ uint32_t carry_s;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint32_t* row = counts + (size_t)digit * nblocks;
  __syncthreads();
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += kSortThreads) {
    const int i = base + t;
    const uint32_t c = (i < nblocks) ? row[i] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t before = carry_s;
    for (int w = 0; w < warp; ++w) before += warp_sum[w];
    if (i < nblocks) row[i] = before + incl - c;
    __syncthreads();
    if (t == kSortThreads - 1) carry_s = before + incl;
    __syncthreads();
  }
  if (t == 0) totals[digit] = carry_s;
}

__global__ void __launch_bounds__(kSortThreads)
scan_digit_rows_kernel(uint32_t* __restrict__ counts, int nblocks, uint32_t* __restrict__ totals) {
  scan_digit_row(blockIdx.x, counts, nblocks, totals);
}

// Stable scatter.  Warp w owns the contiguous slice [base + w*256, base + (w+1)*256) of the tile and
// walks it in 8 rounds of 32 consecutive items, keeping a private per-digit counter row in shared
// memory: the rank of an item inside its digit = (items of that digit in earlier rounds of the warp)
// + (match-any rank inside the round).  One block-wide exclusive prefix over the 8 warp rows then
// gives every item its global slot.  No atomics on ordered data -> deterministic and stable.
__device__ __forceinline__ void radix_scatter_tile(int bid, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                                                   int shift, int radix_bits, const uint32_t* __restrict__ offsets,
                                                   const uint32_t* __restrict__ totals, int nblocks) {
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t goff[kMaxRadix];          // global start of (digit, this block)
  __shared__ uint32_t dbase[kMaxRadix];         // exclusive scan of the digit totals
  __shared__ uint32_t wcnt[kWarps][kMaxRadix];  // per-warp digit counters -> exclusive prefix over warps
  const int radix = 1 << radix_bits;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  __syncthreads();                       // (a previous tile of the same block may still be reading the tables)
  for (int d = t; d < radix; d += kSortThreads) {
    dbase[d] = totals[d];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) wcnt[w][d] = 0;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of <= 512 totals by one warp, 16 per lane
    constexpr int kPer = kMaxRadix / 32;
    uint32_t loc[kPer];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { const int d = lane * kPer + j; loc[j] = (d < radix) ? dbase[d] : 0u; s += loc[j]; }
    uint32_t incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    uint32_t run = incl - s;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { const int d = lane * kPer + j; if (d < radix) dbase[d] = run; run += loc[j]; }
  }
  __syncthreads();
  for (int d = t; d < radix; d += kSortThreads) goff[d] = dbase[d] + offsets[(size_t)d * nblocks + bid];
  const int64_t base = (int64_t)bid * kSortTile + warp * (32 * kSortRounds);
  uint32_t key[kSortRounds], val[kSortRounds], rank[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + r * 32 + lane;
    key[r] = 0; val[r] = 0;
    if (i < n) { key[r] = keys_in[i]; val[r] = vals_in[i]; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const bool valid = base + r * 32 + lane < n;
    const uint32_t digit = valid ? ((key[r] >> shift) & (radix - 1)) : (uint32_t)radix;  // invalid lanes: own class
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t in_round = __popc(peers & ((1u << lane) - 1));
    uint32_t before = 0;
    if (valid) before = wcnt[warp][digit];
    __syncwarp();
    if (valid && in_round == 0) wcnt[warp][digit] = before + __popc(peers);
    __syncwarp();
    rank[r] = before + in_round;
  }
  __syncthreads();
  for (int d = t; d < radix; d += kSortThreads) {
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { const uint32_t c = wcnt[w][d]; wcnt[w][d] = acc; acc += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    if (base + r * 32 + lane < n) {
      const uint32_t digit = (key[r] >> shift) & (radix - 1);
      const uint32_t pos = goff[digit] + wcnt[warp][digit] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                     int radix_bits, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ totals,
                     int nblocks) {
  radix_scatter_tile(blockIdx.x, keys_in, vals_in, keys_out, vals_out, n, shift, radix_bits, offsets, totals, nblocks);
}

// ---- sorted-COO path in 1 + 2 * passes launches (was 2 + 3 * passes): the histogram of pass 0 and the per-column counts
// ride in the preparation kernel, the histogram of pass p + 1 is accumulated by the scatter of pass p (integer atomics on
// (digit, destination tile) counters: order-free), and the last scatter writes the CSC payload itself while extra blocks
// of the same launch turn the column counts into colptr (bucket base of the last digit + a block scan inside the bucket).
// Same ranks as the chain above, so the result is bit-identical (tests/test_gpu_graph_build.py).
__global__ void __launch_bounds__(kSortThreads)
prep_hist_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col, const float* __restrict__ val, int64_t n,
                 uint32_t m, uint32_t ncols, uint32_t* __restrict__ col_out, float* __restrict__ val_out,
                 int32_t* __restrict__ rowptr, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                 uint32_t* __restrict__ status, int radix_bits, uint32_t* __restrict__ counts, int nblocks,
                 uint32_t* __restrict__ colcount) {
  __shared__ uint32_t h[kMaxRadix];
  const int radix = 1 << radix_bits;
  for (int d = threadIdx.x; d < radix; d += kSortThreads) h[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t e = base + r * kSortThreads + threadIdx.x;
    prep_sorted_entry(e, row, col, val, n, m, ncols, col_out, val_out, rowptr, keys, vals, status);
    if (e < n) {
      const uint32_t c = col[e];
      atomicAdd(&h[c & (radix - 1)], 1u);
      if (c < ncols) atomicAdd(&colcount[c], 1u);
    }
  }
  __syncthreads();
  if ((int)blockIdx.x < nblocks)
    for (int d = threadIdx.x; d < radix; d += kSortThreads) counts[(size_t)d * nblocks + blockIdx.x] = h[d];
}

// scan of the digit rows (as scan_digit_rows_kernel) that also clears the counter table the NEXT scatter accumulates into
__global__ void __launch_bounds__(kSortThreads)
scan_digit_rows_zero_kernel(uint32_t* __restrict__ counts, int nblocks, uint32_t* __restrict__ totals,
                            uint32_t* __restrict__ zero, int64_t zero_words) {
  for (int64_t i = (int64_t)blockIdx.x * kSortThreads + threadIdx.x; i < zero_words; i += (int64_t)gridDim.x * kSortThreads) zero[i] = 0;
  scan_digit_row(blockIdx.x, counts, nblocks, totals);
}

// Stable scatter of one pass (ranks exactly as radix_scatter_tile).  kLast = false: writes the permuted pairs and counts
// the NEXT pass's digits per destination tile.  kLast = true: blocks [0, nblocks) write the CSC payload of their items,
// blocks [nblocks, nblocks + radix) write colptr of the columns whose last digit is blockIdx.x - nblocks.
template <bool kLast>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_fused_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                           uint32_t* __restrict__ vals_out, int64_t n, int shift, int radix_bits, const uint32_t* __restrict__ offsets,
                           const uint32_t* __restrict__ totals, int nblocks, uint32_t* __restrict__ next_counts, int next_bits,
                           const uint32_t* __restrict__ rows, const float* __restrict__ val, uint32_t* __restrict__ csr2csc,
                           uint32_t* __restrict__ row_csc, float* __restrict__ val_csc, const uint32_t* __restrict__ colcount,
                           int32_t* __restrict__ colptr, uint32_t ncols) {
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t goff[kMaxRadix];
  __shared__ uint32_t dbase[kMaxRadix];
  __shared__ uint32_t wcnt[kWarps][kMaxRadix];
  const int radix = 1 << radix_bits;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int bid = blockIdx.x;
  for (int d = t; d < radix; d += kSortThreads) {
    dbase[d] = totals[d];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) wcnt[w][d] = 0;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of <= 512 totals by one warp, 16 per lane
    constexpr int kPer = kMaxRadix / 32;
    uint32_t loc[kPer];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { const int d = lane * kPer + j; loc[j] = (d < radix) ? dbase[d] : 0u; s += loc[j]; }
    uint32_t incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    uint32_t run = incl - s;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { const int d = lane * kPer + j; if (d < radix) dbase[d] = run; run += loc[j]; }
  }
  __syncthreads();
  if (kLast && bid >= nblocks) {
    // ---- colptr of the bucket of last-pass digit g: columns [g << shift, (g + 1) << shift)
    const uint32_t g = (uint32_t)(bid - nblocks);
    const uint64_t cbeg = (uint64_t)g << shift, cend = min((uint64_t)(g + 1) << shift, (uint64_t)ncols);
    if (g == 0 && t == 0) colptr[ncols] = (int32_t)n;
    uint32_t* wsum = &wcnt[0][0];        // reuse: [kWarps] warp sums + carry
    uint32_t running = dbase[g];
    for (uint64_t c0 = cbeg; c0 < cend; c0 += kSortThreads) {
      const uint64_t c = c0 + t;
      const uint32_t cnt = (c < cend) ? colcount[c] : 0u;
      uint32_t incl = cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      __syncthreads();                   // (previous iteration's readers of wsum are done)
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      uint32_t before = running;
      for (int w = 0; w < warp; ++w) before += wsum[w];
      if (c < cend) colptr[c] = (int32_t)(before + incl - cnt);
      uint32_t tot = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += wsum[w];
      running += tot;
    }
    return;
  }
  for (int d = t; d < radix; d += kSortThreads) goff[d] = dbase[d] + offsets[(size_t)d * nblocks + bid];
  const int64_t base = (int64_t)bid * kSortTile + warp * (32 * kSortRounds);
  uint32_t key[kSortRounds], pay[kSortRounds], rank[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + r * 32 + lane;
    key[r] = 0; pay[r] = 0;
    if (i < n) { key[r] = keys_in[i]; pay[r] = vals_in[i]; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const bool valid = base + r * 32 + lane < n;
    const uint32_t digit = valid ? ((key[r] >> shift) & (radix - 1)) : (uint32_t)radix;  // invalid lanes: own class
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t in_round = __popc(peers & ((1u << lane) - 1));
    uint32_t before = 0;
    if (valid) before = wcnt[warp][digit];
    __syncwarp();
    if (valid && in_round == 0) wcnt[warp][digit] = before + __popc(peers);
    __syncwarp();
    rank[r] = before + in_round;
  }
  __syncthreads();
  for (int d = t; d < radix; d += kSortThreads) {
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { const uint32_t c = wcnt[w][d]; wcnt[w][d] = acc; acc += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    if (base + r * 32 + lane < n) {
      const uint32_t digit = (key[r] >> shift) & (radix - 1);
      const uint32_t pos = goff[digit] + wcnt[warp][digit] + rank[r];
      if (kLast) {
        const uint32_t p = pay[r];
        csr2csc[pos] = p;
        row_csc[pos] = rows[p];
        val_csc[pos] = val[p];
      } else {
        keys_out[pos] = key[r];
        vals_out[pos] = pay[r];
        const uint32_t nd = (key[r] >> (shift + radix_bits)) & ((1u << next_bits) - 1);
        atomicAdd(&next_counts[(size_t)nd * nblocks + pos / kSortTile], 1u);
      }
    }
  }
}

// ---- sorted-COO path as ONE cooperative launch: prep + the passes of the CSC radix sort + finish, with grid-wide barriers
// where the chain had kernel boundaries (8 launches of ~7 us each for ~24 MB of traffic at BASELINE C2 size).  Block b
// owns tile b of every pass (the grid covers all tiles: the caller checks the co-resident capacity); the digit rows of
// the scan and the element-wise first / last phases are strided over the grid.  Same device functions as the chain, so
// the results are bit-identical.
__global__ void __launch_bounds__(kSortThreads)
graph_sorted_fused_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col, const float* __restrict__ val, int64_t z,
                          uint32_t m, uint32_t ncols, uint32_t* __restrict__ col_out, float* __restrict__ val_out,
                          int32_t* __restrict__ rowptr, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1,
                          uint32_t* counts, uint32_t* totals, int nblocks, int passes, int digit_bits, uint32_t* __restrict__ status,
                          uint32_t* __restrict__ csr2csc, uint32_t* __restrict__ row_csc, float* __restrict__ val_csc,
                          int32_t* __restrict__ colptr) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int G = gridDim.x, b = blockIdx.x;
  for (int64_t e = (int64_t)b * kSortThreads + threadIdx.x; e <= z; e += (int64_t)G * kSortThreads)
    prep_sorted_entry(e, row, col, val, z, m, ncols, col_out, val_out, rowptr, k0, v0, status);
  uint32_t *ki = k0, *vi = v0, *ko = k1, *vo = v1;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * digit_bits;
    grid.sync();
    for (int tile = b; tile < nblocks; tile += G) radix_hist_tile(tile, ki, z, shift, digit_bits, counts, nblocks);
    grid.sync();
    for (int d = b; d < (1 << digit_bits); d += G) scan_digit_row(d, counts, nblocks, totals);
    grid.sync();
    for (int tile = b; tile < nblocks; tile += G) radix_scatter_tile(tile, ki, vi, ko, vo, z, shift, digit_bits, counts, totals, nblocks);
    uint32_t* tk = ki; ki = ko; ko = tk;
    uint32_t* tv = vi; vi = vo; vo = tv;
  }
  grid.sync();
  for (int64_t k = (int64_t)b * kSortThreads + threadIdx.x; k <= z; k += (int64_t)G * kSortThreads)
    finish_csc_entry(k, row, val_out, ki, vi, z, (int32_t)ncols, csr2csc, row_csc, val_csc, colptr);
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

// 1 builds sorted inputs with ONE cooperative launch (lpgnn_set_graph_fused; environment LPGNN_GRAPH_FUSED=1).  Off by
// default: measured on B200 the grid barriers cost what the launches did (C2: 68 vs 62 us) and small LPs lose the
// launch pipelining of the chain (C5, one call per LP: 4.7K vs 5.5K LPs/s); kept as an option and as a cross-check.
int g_graph_fused = [] { const char* e = getenv("LPGNN_GRAPH_FUSED"); return e ? atoi(e) != 0 : 0; }();

// co-resident blocks of the fused sorted-path kernel (0: cooperative launches unavailable)
int fused_blocks() {
  static int cap = -1;
  if (cap < 0) {
    int per_sm = 0, coop = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, graph_sorted_fused_kernel, kSortThreads, 0) != cudaSuccess)
      per_sm = 0;
    cudaGetLastError();
    cap = per_sm * sm_count();
  }
  return cap;
}

// 1 runs the sorted path in 1 + 2 * passes launches (lpgnn_set_graph_compact; environment LPGNN_GRAPH_COMPACT=1).  Off by
// default: measured on B200 (ncu launch list of one C2 step) the fused kernels take exactly what their parts took --
// prep + histogram 12.6 us (6.7 + 4.2), scatter + next histogram 18.0 (13.2 + 4.2), last scatter + payload + colptr 20.5
// (13.2 + 7.0) -- and with the stream fed ahead the three saved launches hide nothing: C2 1 325 vs 1 334 LPs/s, C5 one call
// per LP 7.1K vs 7.3K.  Kept as an option (13.6 instead of 16.4 launches per LP) and as a cross-check.
int g_graph_compact = [] { const char* e = getenv("LPGNN_GRAPH_COMPACT"); return e ? atoi(e) != 0 : 0; }();

struct SortBufs {
  uint32_t *k[2], *v[2];
  uint32_t* counts;
  uint32_t* totals;  // [kMaxRadix]
};

// Sorts (k[0], v[0]) by key bits [0, bits); returns the index (0/1) of the buffer holding the result.
int radix_sort(SortBufs& b, int64_t n, int bits, cudaStream_t st) {
  const int nblocks = ceil_div(n, kSortTile);
  const int passes = (bits + kMaxRadixBits - 1) / kMaxRadixBits;
  const int digit_bits = (bits + passes - 1) / passes;  // <= 9
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * digit_bits;
    radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(b.k[cur], n, shift, digit_bits, b.counts, nblocks);
    scan_digit_rows_kernel<<<1 << digit_bits, kSortThreads, 0, st>>>(b.counts, nblocks, b.totals);
    radix_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(b.k[cur], b.v[cur], b.k[cur ^ 1], b.v[cur ^ 1], n,
                                                           shift, digit_bits, b.counts, b.totals, nblocks);
    cur ^= 1;
    count_launches(3);
  }
  return cur;
}

// val[e] /= deg(row) for the entries of every row (one thread per row: LP rows are short); IEEE division, so the
// result equals numpy's float32 `val / deg` bit for bit
__global__ void mean_normalize_kernel(const int32_t* __restrict__ ptr, float* __restrict__ val, int32_t rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int32_t beg = ptr[r], end = ptr[r + 1];
  if (end - beg <= 1) return;
  const float d = (float)(end - beg);
  for (int32_t e = beg; e < end; ++e) val[e] = __fdiv_rn(val[e], d);
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" size_t lpgnn_graph_build_workspace_bytes(int64_t nnz, int32_t m, int32_t n) {
  (void)m;
  const size_t z = (size_t)(nnz > 0 ? nnz : 1);
  const size_t words = align_up(z, 64);
  const size_t nblocks = (z + kSortTile - 1) / kSortTile;
  // k0,k1,v0,v1,r32,c32,rows_sorted + counts + digit totals + [second counter table | per-column counts] (compact sorted path)
  return (7 * words + 2 * align_up(kMaxRadix * nblocks, 64) + kMaxRadix + 64 + align_up((size_t)(n > 0 ? n : 0) + 2, 64)) *
             sizeof(uint32_t) + 256;
}

extern "C" int lpgnn_graph_build(const void* coo_row, const void* coo_col, int idx_is_i64, const float* coo_val,
                                 int64_t nnz, int32_t m, int32_t n, int flags, int32_t* rowptr, int32_t* col,
                                 float* val, int32_t* colptr, int32_t* row_csc, float* val_csc, int32_t* csr2csc,
                                 int32_t* status, void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
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
  b.totals = b.counts + align_up((size_t)kMaxRadix * ceil_div(z, kSortTile), 64);
  const int tb = 256, gb = ceil_div(z, tb), gb1 = ceil_div(z + 1, tb);
  int launches = 0;
  uint32_t* u_status = reinterpret_cast<uint32_t*>(status);

  const uint32_t *rsrc, *csrc;
  if (idx_is_i64) {
    narrow_i64_kernel<<<gb, tb, 0, st>>>(reinterpret_cast<const int64_t*>(coo_row), r32, z);
    narrow_i64_kernel<<<gb, tb, 0, st>>>(reinterpret_cast<const int64_t*>(coo_col), c32, z);
    rsrc = r32; csrc = c32;
    launches += 2;
  } else {
    rsrc = reinterpret_cast<const uint32_t*>(coo_row);
    csrc = reinterpret_cast<const uint32_t*>(coo_col);
  }
  uint32_t* u_col = reinterpret_cast<uint32_t*>(col);
  const uint32_t* csr_rows;  // row of every CSR entry
  if ((flags & LPGNN_COO_SORTED) && !(flags & LPGNN_GRAPH_MEAN) && g_graph_fused && fused_blocks() > 0) {
    // sorted COO, one cooperative launch (see graph_sorted_fused_kernel); the grid is capped at the co-resident capacity
    // and strides over the tiles
    const int nblocks = ceil_div(z, kSortTile);
    const int bits = bits_for(n);
    int passes = (bits + kMaxRadixBits - 1) / kMaxRadixBits;
    int digit_bits = (bits + passes - 1) / passes;
    const int cap = fused_blocks();
    int grid = nblocks < cap ? (nblocks > 0 ? nblocks : 1) : cap;
    const int64_t zz = z;
    const uint32_t um = (uint32_t)m, un = (uint32_t)n;
    uint32_t *k0 = b.k[0], *v0 = b.v[0], *k1 = b.k[1], *v1 = b.v[1], *cnt = b.counts, *tot = b.totals;
    int nb = nblocks;
    uint32_t* c2c = reinterpret_cast<uint32_t*>(csr2csc);
    uint32_t* rcsc = reinterpret_cast<uint32_t*>(row_csc);
    void* args[] = {(void*)&rsrc, (void*)&csrc, (void*)&coo_val, (void*)&zz, (void*)&um, (void*)&un, (void*)&u_col, (void*)&val,
                    (void*)&rowptr, (void*)&k0, (void*)&v0, (void*)&k1, (void*)&v1, (void*)&cnt, (void*)&tot, (void*)&nb,
                    (void*)&passes, (void*)&digit_bits, (void*)&u_status, (void*)&c2c, (void*)&rcsc, (void*)&val_csc,
                    (void*)&colptr};
    LPGNN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)graph_sorted_fused_kernel, dim3(grid), dim3(kSortThreads), args, 0, st));
    count_launches(launches + 1);
    return LPGNN_OK;
  }
  if ((flags & LPGNN_COO_SORTED) && g_graph_compact) {
    // Caller asserts row-major order (dataset.py:251-252); CSR arrays = the input, CSC by a stable sort on the column, in
    // 1 + 2 * passes launches (see prep_hist_kernel / radix_scatter_fused_kernel)
    const int nblocks = ceil_div(z, kSortTile);
    const int bits = bits_for(n);
    const int passes = (bits + kMaxRadixBits - 1) / kMaxRadixBits;
    const int digit_bits = (bits + passes - 1) / passes;  // <= 9
    const size_t table = align_up((size_t)kMaxRadix * nblocks, 64);
    uint32_t* counts2 = b.totals + kMaxRadix + 64;         // second counter table, then the per-column counts
    uint32_t* colcount = counts2 + table;
    uint32_t* tab[2] = {b.counts, counts2};
    LPGNN_CUDA_OK(cudaMemsetAsync(counts2, 0, (table + (size_t)n + 2) * sizeof(uint32_t), st));
    prep_hist_kernel<<<ceil_div(z + 1, kSortTile), kSortThreads, 0, st>>>(rsrc, csrc, coo_val, z, (uint32_t)m, (uint32_t)n, u_col, val,
                                                                          rowptr, b.k[0], b.v[0], u_status, digit_bits, tab[0],
                                                                          nblocks, colcount);
    launches += 1;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
      const int shift = p * digit_bits;
      const bool last = p == passes - 1;
      // pass p reads table p % 2; the scatter of pass p fills table (p + 1) % 2, cleared by the memset (p = 0) or by the scan of
      // pass p (it held the counts of pass p - 1, fully consumed by then)
      if (p >= 1 && !last)
        scan_digit_rows_zero_kernel<<<1 << digit_bits, kSortThreads, 0, st>>>(tab[p & 1], nblocks, b.totals, tab[(p + 1) & 1],
                                                                              (int64_t)table);
      else
        scan_digit_rows_kernel<<<1 << digit_bits, kSortThreads, 0, st>>>(tab[p & 1], nblocks, b.totals);
      if (last)
        radix_scatter_fused_kernel<true><<<nblocks + (1 << digit_bits), kSortThreads, 0, st>>>(
            b.k[cur], b.v[cur], nullptr, nullptr, z, shift, digit_bits, tab[p & 1], b.totals, nblocks, nullptr, 0, rsrc, coo_val,
            reinterpret_cast<uint32_t*>(csr2csc), reinterpret_cast<uint32_t*>(row_csc), val_csc, colcount, colptr, (uint32_t)n);
      else
        radix_scatter_fused_kernel<false><<<nblocks, kSortThreads, 0, st>>>(
            b.k[cur], b.v[cur], b.k[cur ^ 1], b.v[cur ^ 1], z, shift, digit_bits, tab[p & 1], b.totals, nblocks, tab[(p + 1) & 1],
            digit_bits, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0u);
      cur ^= 1;
      launches += 2;
    }
    if (flags & LPGNN_GRAPH_MEAN) {
      mean_normalize_kernel<<<ceil_div(m, 256), 256, 0, st>>>(rowptr, val, m);
      mean_normalize_kernel<<<ceil_div(n, 256), 256, 0, st>>>(colptr, val_csc, n);
      launches += 2;
    }
    LPGNN_LAUNCH_OK();
    count_launches(launches);
    return LPGNN_OK;
  }
  if (flags & LPGNN_COO_SORTED) {
    // Caller asserts row-major order (what the reference's pipeline produces, dataset.py:251-252): the CSR
    // arrays are the input; the claim and the index range are verified on the device and reported in *status.
    prep_sorted_kernel<<<gb1, tb, 0, st>>>(rsrc, csrc, coo_val, z, (uint32_t)m, (uint32_t)n, u_col, val, rowptr, b.k[0],
                                           b.v[0], u_status);
    csr_rows = rsrc;
    launches += 1;
  } else {
    // ---- CSR: LSD over (row, col): column digits first, then row digits
    init_pairs_kernel<<<gb, tb, 0, st>>>(csrc, b.k[0], b.v[0], z);
    int cur = radix_sort(b, z, bits_for(n), st);
    if (cur != 0) std::swap(b.k[0], b.k[1]), std::swap(b.v[0], b.v[1]);
    gather_u32_kernel<<<gb, tb, 0, st>>>(rsrc, b.v[0], b.k[0], z);  // keys := row of each (col-sorted) entry
    cur = radix_sort(b, z, bits_for(m), st);
    // b.k[cur] = rows of the CSR entries (sorted), b.v[cur] = original COO index of each CSR entry
    copy_u32_kernel<<<gb, tb, 0, st>>>(b.k[cur], rows_sorted, z);
    gather_entry_kernel<<<gb, tb, 0, st>>>(csrc, coo_val, b.v[cur], u_col, val, z);
    csr_rows = rows_sorted;
    if (status) {
      expand_check_range_kernel<<<gb, tb, 0, st>>>(rsrc, csrc, z, (uint32_t)m, (uint32_t)n, u_status);
      ++launches;
    }
    fill_ptr_kernel<<<gb1, tb, 0, st>>>(csr_rows, z, m, rowptr);
    init_pairs_kernel<<<gb, tb, 0, st>>>(u_col, b.k[0], b.v[0], z);
    launches += 6;
  }
  // ---- CSC view: stable sort of the CSR entries by column, then payload + colptr in one pass
  const int cur = radix_sort(b, z, bits_for(n), st);
  finish_csc_kernel<<<gb1, tb, 0, st>>>(csr_rows, val, b.k[cur], b.v[cur], z, n, reinterpret_cast<uint32_t*>(csr2csc),
                                        reinterpret_cast<uint32_t*>(row_csc), val_csc, colptr);
  launches += 1;
  if (flags & LPGNN_GRAPH_MEAN) {
    // mean aggregation (PyG aggr='mean', the option left commented out at reference arch.py:57,60): every orientation's
    // values are divided by the degree of ITS destination node, so lpgnn_spmm returns neighbourhood means
    mean_normalize_kernel<<<ceil_div(m, 256), 256, 0, st>>>(rowptr, val, m);
    mean_normalize_kernel<<<ceil_div(n, 256), 256, 0, st>>>(colptr, val_csc, n);
    launches += 2;
  }
  LPGNN_LAUNCH_OK();
  count_launches(launches);
  return LPGNN_OK;
}

// ---------------------------------------------------------------------------------------------- block-diagonal packs
namespace lpgnn {
namespace {
// Entries of LP b sit at [edge_ptr[b], edge_ptr[b+1]) with LP-local indices; shift them to the pack's numbering.
__global__ void pack_offsets_kernel(int32_t* __restrict__ row, int32_t* __restrict__ col, int64_t nnz,
                                    const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ cons_ptr,
                                    const int32_t* __restrict__ vars_ptr, int n_seg) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int lo = 0, hi = n_seg;                       // largest b with edge_ptr[b] <= e
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (edge_ptr[mid] <= e) lo = mid; else hi = mid; }
  row[e] += cons_ptr[lo];
  col[e] += vars_ptr[lo];
}
}  // namespace
}  // namespace lpgnn

namespace lpgnn {
namespace {
// One pass from the STAGED LPs (each LP's host pack [row | col | val | x_s | x_t] copied verbatim, LP after LP) to the pack
// layout [row Z | col Z | val Z | x_s M*p | x_t N*q], indices shifted to the pack's numbering (the job of
// pack_offsets_kernel, fused).  One 4-byte word per thread; the LP of a word by binary search over the staging offsets.
__global__ void pack_scatter_kernel(const int32_t* __restrict__ staged, const int32_t* __restrict__ stage_off,
                                    const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ cons_ptr,
                                    const int32_t* __restrict__ vars_ptr, int n_seg, int p, int q, int64_t total_words,
                                    int32_t* __restrict__ row, int32_t* __restrict__ col, int32_t* __restrict__ val,
                                    int32_t* __restrict__ x_s, int32_t* __restrict__ x_t) {
  const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (w >= total_words) return;
  int lo = 0, hi = n_seg;                       // largest b with stage_off[b] <= w
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (stage_off[mid] <= w) lo = mid; else hi = mid; }
  const int b = lo;
  const int32_t e0 = edge_ptr[b], z = edge_ptr[b + 1] - e0;
  const int32_t c0 = cons_ptr[b], mb = cons_ptr[b + 1] - c0;
  const int32_t v0 = vars_ptr[b];
  int64_t k = w - stage_off[b];
  const int32_t x = staged[w];
  if (k < z) { row[e0 + k] = x + c0; return; }
  k -= z;
  if (k < z) { col[e0 + k] = x + v0; return; }
  k -= z;
  if (k < z) { val[e0 + k] = x; return; }
  k -= z;
  if (k < (int64_t)mb * p) { x_s[(int64_t)c0 * p + k] = x; return; }
  k -= (int64_t)mb * p;
  x_t[(int64_t)v0 * q + k] = x;
}
}  // namespace
}  // namespace lpgnn

extern "C" int lpgnn_pack_scatter(const int32_t* staged, const int32_t* stage_off, const int32_t* edge_ptr,
                                  const int32_t* cons_ptr, const int32_t* vars_ptr, int32_t n_segments, int32_t p, int32_t q,
                                  int64_t total_words, int32_t* row, int32_t* col, float* val, float* x_s, float* x_t,
                                  lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(total_words >= 0 && n_segments >= 1 && p > 0 && q > 0, "pack_scatter: bad arguments");
  if (total_words == 0) return LPGNN_OK;
  LPGNN_REQUIRE(staged && stage_off && edge_ptr && cons_ptr && vars_ptr && row && col && val && x_s && x_t, "pack_scatter: null pointer");
  pack_scatter_kernel<<<ceil_div(total_words, 256), 256, 0, (cudaStream_t)stream>>>(
      staged, stage_off, edge_ptr, cons_ptr, vars_ptr, n_segments, p, q, total_words, row, col, reinterpret_cast<int32_t*>(val),
      reinterpret_cast<int32_t*>(x_s), reinterpret_cast<int32_t*>(x_t));
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_pack_offsets(int32_t* row, int32_t* col, int64_t nnz, const int32_t* edge_ptr,
                                  const int32_t* cons_ptr, const int32_t* vars_ptr, int32_t n_segments,
                                  lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(nnz >= 0 && n_segments >= 1, "pack_offsets: bad arguments");
  if (nnz == 0) return LPGNN_OK;
  LPGNN_REQUIRE(row && col && edge_ptr && cons_ptr && vars_ptr, "pack_offsets: null pointer");
  pack_offsets_kernel<<<ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(row, col, nnz, edge_ptr, cons_ptr, vars_ptr,
                                                                           n_segments);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

// Tuning knob: 1 = sorted COO inputs are built by ONE cooperative launch, 0 (default) = the launch chain.  Results
// are bit-identical (same device functions).  Returns the previous setting.
extern "C" int lpgnn_set_graph_compact(int enable) {
  const int prev = lpgnn::g_graph_compact;
  lpgnn::g_graph_compact = enable ? 1 : 0;
  return prev;
}

extern "C" int lpgnn_set_graph_fused(int enable) {
  const int prev = lpgnn::g_graph_fused;
  lpgnn::g_graph_fused = enable ? 1 : 0;
  return prev;
}
