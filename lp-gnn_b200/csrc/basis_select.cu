// (a6) Basis decision: softmax -> global top-m on P(basic) -> status per node.
//
// Replaces val.inference_gnn (reference val.py:106-124): F.softmax, isnan mask, topk(m) over all
// m+n nodes (ATen sort-based for large k), two scatter writes, argmax and two .item() host syncs,
// with a sync-free sequence of small kernels:
//   1. softmax per node (fp32) -> key = bit pattern of p1 (p1 >= 0, so unsigned order == float
//      order), side = 0 if p0 >= p2 else 2 (argmax over {0,2} picks the first maximum);
//      fused with the histogram of the top radix digit
//   2. 4-pass MSB radix select of the k-th largest key (256-bin histograms, integer atomics only); every kernel
//      re-derives the threshold state from the histograms with one warp, so there are no state-advancing launches
//   3. ties at the threshold are taken in ascending node index (block counts + scan + in-block
//      ranks: deterministic; torch.topk leaves the tie order implementation-defined)
//   4. status = 1 if selected else side.
// HBM-bound integer/byte work: (m+n)*(12 + 4 + 1) bytes in pass 1, (m+n)*4 per select pass.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;                     // nodes per thread in the ordered passes
constexpr int kTile = kThreads * kItems;      // nodes per block in the ordered passes

// Threshold search state.  Every kernel re-derives it from the per-pass histograms (a chain of up to four
// 256-bin picks done by one warp), so no kernel exists only to advance the state and there is no host sync.
struct Threshold {
  uint32_t prefix;  // decided high bits of the k-th largest key
  uint32_t mask;    // which bits are decided
  uint32_t k_rem;   // how many keys are still to be taken among those matching the prefix
};

__device__ __forceinline__ float nan_to_zero(float v) { return (v != v) ? 0.f : v; }

// Executed by ONE warp.  hist: [npass][256].  Picks, pass by pass, the digit d with
// count(digit > d) < k_rem <= count(digit >= d) among the keys that match the prefix so far.
__device__ Threshold pick_chain(const uint32_t* __restrict__ hist, int npass, uint32_t k) {
  Threshold t{0u, 0u, k};
  const int lane = threadIdx.x & 31;
  if (k == 0) { t.prefix = 0xffffffffu; t.mask = 0xffffffffu; return t; }  // threshold above every key
  for (int p = 0; p < npass; ++p) {
    const uint32_t* h = hist + p * 256;
    const int shift = 24 - 8 * p;
    // lane l owns the 8 bins 255-8l .. 248-8l (descending digit order)
    uint32_t c[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = h[255 - 8 * lane - j]; s += c[j]; }
    uint32_t incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const uint32_t before = incl - s;                       // keys with a larger digit owned by lower lanes
    const bool mine = before < t.k_rem && t.k_rem <= incl;
    const uint32_t who = __ballot_sync(0xffffffffu, mine);
    uint32_t digit = 0, above = 0;
    if (who) {
      const int src = __ffs(who) - 1;
      if (lane == src) {
        uint32_t acc = before;
        int j = 0;
        for (; j < 7; ++j) { if (acc + c[j] >= t.k_rem) break; acc += c[j]; }
        digit = 255 - 8 * lane - j;
        above = acc;
      }
      digit = __shfl_sync(0xffffffffu, digit, src);
      above = __shfl_sync(0xffffffffu, above, src);
    } else {                                                 // k exceeds the population: take everything
      above = __shfl_sync(0xffffffffu, incl, 31) - h[0];
    }
    t.prefix |= digit << shift;
    t.mask |= 0xffu << shift;
    t.k_rem -= above;
  }
  return t;
}

__device__ __forceinline__ Threshold block_threshold(const uint32_t* hist, int npass, uint32_t k, Threshold* smem_t) {
  if (threadIdx.x < 32) {
    const Threshold t = pick_chain(hist, npass, k);
    if (threadIdx.x == 0) *smem_t = t;
  }
  __syncthreads();
  return *smem_t;
}

__global__ void init_hist_kernel(uint32_t* hist /*[4][256]*/, int32_t* counts) {
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) hist[i] = 0;
  if (counts && threadIdx.x < 4) counts[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(kThreads)
softmax_key_kernel(const float* __restrict__ lc, int32_t m, const float* __restrict__ lv, int32_t n,
                   uint32_t* __restrict__ keys, uint8_t* __restrict__ side, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t total = (int64_t)m + n;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const float* p = (i < m) ? (lc + i * 3) : (lv + (i - m) * 3);
    const float x0 = p[0], x1 = p[1], x2 = p[2];
    const float mx = fmaxf(x0, fmaxf(x1, x2));
    const float e0 = expf(x0 - mx), e1 = expf(x1 - mx), e2 = expf(x2 - mx);
    const float s = e0 + e1 + e2;
    const float p0 = nan_to_zero(e0 / s), p1 = nan_to_zero(e1 / s), p2 = nan_to_zero(e2 / s);
    const uint32_t key = __float_as_uint(p1);
    keys[i] = key;
    side[i] = (p0 >= p2) ? 0 : 2;
    atomicAdd(&h[key >> 24], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// histogram of digit `pass` (bits [24-8*pass, 32-8*pass)) over the keys matching the threshold prefix so far
__global__ void __launch_bounds__(kThreads)
digit_hist_kernel(const uint32_t* __restrict__ keys, int64_t total, uint32_t* __restrict__ hist, int pass, uint32_t k) {
  __shared__ uint32_t h[256];
  __shared__ Threshold ts;
  h[threadIdx.x] = 0;
  const Threshold t = block_threshold(hist, pass, k, &ts);   // includes a __syncthreads()
  const int shift = 24 - 8 * pass;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const uint32_t key = keys[i];
    if ((key & t.mask) == t.prefix) atomicAdd(&h[(key >> shift) & 0xff], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[pass * 256 + threadIdx.x], h[threadIdx.x]);
}

// number of keys equal to the threshold in each tile
__global__ void __launch_bounds__(kThreads)
tie_count_kernel(const uint32_t* __restrict__ keys, int64_t total, const uint32_t* __restrict__ hist, uint32_t k,
                 uint32_t* __restrict__ tile_ties) {
  __shared__ uint32_t cnt;
  __shared__ Threshold ts;
  if (threadIdx.x == 0) cnt = 0;
  const Threshold t = block_threshold(hist, 4, k, &ts);
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j)
    if (base + j < total && keys[base + j] == t.prefix) ++c;
  if (c) atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_ties[blockIdx.x] = cnt;
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads)
status_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ side, int64_t total, int32_t m,
              const uint32_t* __restrict__ hist, uint32_t k, const uint32_t* __restrict__ tile_ties,
              OutT* __restrict__ status, int32_t* __restrict__ counts) {
  __shared__ uint32_t warp_ties[kThreads / 32];
  __shared__ uint32_t red[kThreads / 32];
  __shared__ int32_t c_s[4];
  __shared__ Threshold ts;
  if (threadIdx.x < 4) c_s[threadIdx.x] = 0;
  const Threshold t = block_threshold(hist, 4, k, &ts);
  const uint32_t thr = t.prefix, k_rem = t.k_rem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ties in the tiles before this one (fixed order -> lowest node index first)
  uint32_t part = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += kThreads) part += tile_ties[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if (lane == 0) red[warp] = part;
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  uint32_t key[kItems];
  uint32_t my_ties = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    key[j] = (base + j < total) ? keys[base + j] : 0u;
    if (base + j < total && key[j] == thr) ++my_ties;
  }
  // exclusive prefix of my_ties over the block, in thread order (== node order)
  uint32_t incl = my_ties;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) warp_ties[warp] = incl;
  __syncthreads();
  uint32_t before = incl - my_ties;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) { before += red[w]; if (w < warp) before += warp_ties[w]; }
  int32_t n0 = 0, n1 = 0, n2 = 0, nbv = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int64_t i = base + j;
    if (i >= total) break;
    bool sel = key[j] > thr;
    if (key[j] == thr) { sel = before < k_rem; ++before; }
    const int s = sel ? 1 : (int)side[i];
    status[i] = (OutT)s;
    n0 += (s == 0); n1 += (s == 1); n2 += (s == 2); nbv += (s == 1 && i >= m);
  }
  if (counts) {
    if (n0) atomicAdd(&c_s[0], n0);
    if (n1) atomicAdd(&c_s[1], n1);
    if (n2) atomicAdd(&c_s[2], n2);
    if (nbv) atomicAdd(&c_s[3], nbv);
    __syncthreads();
    if (threadIdx.x < 4 && c_s[threadIdx.x]) atomicAdd(&counts[threadIdx.x], c_s[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------ segmented variant
// One CTA per LP of a block-diagonal pack: the whole decision (softmax keys, 4-pass radix select, ordered tie
// break, status) for that LP in a single launch.  Nodes of segment b: constraints [cptr[b], cptr[b+1]) of
// logits_cons and variables [vptr[b], vptr[b+1]) of logits_vars; k = its number of constraints.  Keys and side
// bytes live in the global scratch (L2-resident for the LP sizes this is used for); status is written in the packed
// layout [all constraints | all variables].
constexpr int kSegThreads = 1024;

template <typename OutT>
__global__ void __launch_bounds__(kSegThreads)
segmented_select_kernel(const float* __restrict__ lc, const float* __restrict__ lv, const int32_t* __restrict__ cptr,
                        const int32_t* __restrict__ vptr, int32_t total_cons, uint32_t* __restrict__ keys,
                        uint8_t* __restrict__ side, OutT* __restrict__ status, int lp_major) {
  __shared__ uint32_t hist[4 * 256];
  __shared__ Threshold ts;
  __shared__ uint32_t warp_ties[kSegThreads / 32];
  __shared__ uint32_t run_s;
  const int b = blockIdx.x;
  const int32_t c0 = cptr[b], c1 = cptr[b + 1], v0 = vptr[b], v1 = vptr[b + 1];
  const int32_t mb = c1 - c0, nb = v1 - v0, tot = mb + nb;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  hist[t] = 0;                                       // 4*256 == kSegThreads
  __syncthreads();
  // packed position of local node i
  auto pos = [&](int32_t i) -> int64_t { return i < mb ? (int64_t)c0 + i : (int64_t)total_cons + v0 + (i - mb); };
  for (int32_t i = t; i < tot; i += kSegThreads) {
    const float* p = i < mb ? lc + (int64_t)(c0 + i) * 3 : lv + (int64_t)(v0 + i - mb) * 3;
    const float x0 = p[0], x1 = p[1], x2 = p[2];
    const float mx = fmaxf(x0, fmaxf(x1, x2));
    const float e0 = expf(x0 - mx), e1 = expf(x1 - mx), e2 = expf(x2 - mx);
    const float sum = e0 + e1 + e2;
    const float p0 = nan_to_zero(e0 / sum), p1 = nan_to_zero(e1 / sum), p2 = nan_to_zero(e2 / sum);
    const uint32_t key = __float_as_uint(p1);
    const int64_t g = pos(i);
    keys[g] = key;
    side[g] = (p0 >= p2) ? 0 : 2;
    atomicAdd(&hist[key >> 24], 1u);
  }
  __syncthreads();
  for (int pass = 1; pass < 4; ++pass) {
    if (t < 32) { const Threshold th = pick_chain(hist, pass, (uint32_t)mb); if (t == 0) ts = th; }
    __syncthreads();
    const Threshold th = ts;
    const int shift = 24 - 8 * pass;
    for (int32_t i = t; i < tot; i += kSegThreads) {
      const uint32_t key = keys[pos(i)];
      if ((key & th.mask) == th.prefix) atomicAdd(&hist[pass * 256 + ((key >> shift) & 0xff)], 1u);
    }
    __syncthreads();
  }
  if (t < 32) { const Threshold th = pick_chain(hist, 4, (uint32_t)mb); if (t == 0) { ts = th; run_s = 0; } }
  __syncthreads();
  const uint32_t thr = ts.prefix, k_rem = ts.k_rem;
  // ordered pass: ties at the threshold go to the lowest local node index (constraints first)
  for (int32_t base = 0; base < tot; base += kSegThreads) {
    const int32_t i = base + t;
    uint32_t key = 0; bool tie = false;
    if (i < tot) { key = keys[pos(i)]; tie = key == thr; }
    const uint32_t ballot = __ballot_sync(0xffffffffu, tie);
    if (lane == 0) warp_ties[warp] = __popc(ballot);
    __syncthreads();
    uint32_t before = run_s + __popc(ballot & ((1u << lane) - 1));
    for (int w = 0; w < warp; ++w) before += warp_ties[w];
    if (i < tot) {
      const bool sel = key > thr || (tie && before < k_rem);
      const int64_t g = pos(i);
      // output position: packed [all constraints | all variables], or per LP [cons of LP b | vars of LP b] back to back
      status[lp_major ? (int64_t)c0 + v0 + i : g] = (OutT)(sel ? 1 : (int)side[g]);
    }
    __syncthreads();
    if (t == 0) { uint32_t s = 0; for (int w = 0; w < kSegThreads / 32; ++w) s += warp_ties[w]; run_s += s; }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ one-launch variant
// The whole decision as ONE cooperative kernel: every thread keeps its nodes' keys in registers (contiguous nodes per
// thread, so thread order == node order), the k-th largest key is found by a 3-digit MSB radix select (11 + 11 + 10
// bits) with a grid-wide barrier after each histogram, every block re-derives the threshold from the global histogram,
// and the ordered tie break (rare: only when the threshold key occurs more often than it is needed) costs one more
// barrier.  Replaces the 7-launch chain above for everything that fits the co-resident grid (~4.8M nodes); at BASELINE
// C2 size (150K nodes) the chain was pure launch latency (~50 us for 2.5 MB of traffic).
namespace cg = cooperative_groups;
constexpr int kSelThreads = 512;
constexpr int kBins = 2048;

struct Pick { uint32_t digit, above, at; };   // chosen digit, #keys with a larger digit, #keys in the chosen bin

// exclusive prefix of v over the block in thread order (512 threads); *total = block sum
__device__ __forceinline__ uint32_t sel_block_exclusive(uint32_t v, uint32_t* wsum /*[16]*/, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  __syncthreads();                       // wsum may still be read from a previous call
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kSelThreads / 32; ++w) { const uint32_t x = wsum[w]; if (w < warp) before += x; all += x; }
  *total = all;
  return before + incl - v;
}

// Every block: the digit d with count(digit > d) < k_rem <= count(digit >= d) in the global histogram of this pass.
__device__ __forceinline__ Pick block_pick(const uint32_t* __restrict__ ghist, uint32_t k_rem, uint32_t* wsum, Pick* pick_s) {
  const int t = threadIdx.x;
  uint32_t c[kBins / kSelThreads], s = 0;      // thread t owns bins 2047-4t .. 2044-4t (descending digits)
#pragma unroll
  for (int j = 0; j < kBins / kSelThreads; ++j) { c[j] = __ldcg(ghist + (kBins - 1 - (kBins / kSelThreads) * t - j)); s += c[j]; }
  uint32_t total;
  const uint32_t before = sel_block_exclusive(s, wsum, &total);
  if (before < k_rem && k_rem <= before + s) {
    uint32_t acc = before;
    int j = 0;
    for (; j < kBins / kSelThreads - 1; ++j) { if (acc + c[j] >= k_rem) break; acc += c[j]; }
    pick_s->digit = (uint32_t)(kBins - 1 - (kBins / kSelThreads) * t - j);
    pick_s->above = acc;
    pick_s->at = c[j];
  }
  __syncthreads();
  return *pick_s;
}

template <int MAXI, typename OutT>
__global__ void __launch_bounds__(kSelThreads, 2)
select_fused_kernel(const float* __restrict__ lc, int32_t m, const float* __restrict__ lv, int32_t n, uint32_t k, int ipt,
                    uint32_t* __restrict__ hist /*[3][kBins], zeroed*/, uint32_t* __restrict__ tile_ties /*[grid]*/,
                    OutT* __restrict__ status, int32_t* __restrict__ counts) {
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t h[kBins];
  __shared__ uint32_t wsum[kSelThreads / 32];
  __shared__ Pick pick_s;
  __shared__ int32_t c_s[4];
  const int t = threadIdx.x;
  const int64_t total = (int64_t)m + n;
  const int64_t base = ((int64_t)blockIdx.x * kSelThreads + t) * ipt;
  if (t < 4) c_s[t] = 0;
  for (int i = t; i < kBins; i += kSelThreads) h[i] = 0;
  __syncthreads();
  // ---- softmax keys (bit pattern of p1; p1 >= 0 so unsigned order == float order) + the {0,2} side bit
  uint32_t key[MAXI];
  uint32_t side2 = 0;
#pragma unroll
  for (int j = 0; j < MAXI; ++j) {
    key[j] = 0;
    const int64_t i = base + j;
    if (j < ipt && i < total) {
      const float* p = (i < m) ? (lc + i * 3) : (lv + (i - m) * 3);
      const float x0 = p[0], x1 = p[1], x2 = p[2];
      const float mx = fmaxf(x0, fmaxf(x1, x2));
      const float e0 = expf(x0 - mx), e1 = expf(x1 - mx), e2 = expf(x2 - mx);
      const float sum = e0 + e1 + e2;
      const float p0 = nan_to_zero(e0 / sum), p1 = nan_to_zero(e1 / sum), p2 = nan_to_zero(e2 / sum);
      key[j] = __float_as_uint(p1);
      if (!(p0 >= p2)) side2 |= 1u << j;
      atomicAdd(&h[key[j] >> 21], 1u);
    }
  }
  __syncthreads();
  for (int i = t; i < kBins; i += kSelThreads) if (h[i]) atomicAdd(&hist[i], h[i]);
  uint32_t thr = 0xffffffffu, k_rem = 0, n_ties = 0;
  if (k > 0) {                                                 // (k == 0: threshold above every key, nothing selected)
    grid.sync();
    const Pick p0 = block_pick(hist, k, wsum, &pick_s);
    k_rem = k - p0.above;
    // ---- second digit (bits 10..20) over the keys that share the first
#pragma unroll
    for (int j = 0; j < MAXI; ++j)
      if (j < ipt && base + j < total && (key[j] >> 21) == p0.digit) atomicAdd(&hist[kBins + ((key[j] >> 10) & 0x7ffu)], 1u);
    grid.sync();
    const Pick p1 = block_pick(hist + kBins, k_rem, wsum, &pick_s);
    k_rem -= p1.above;
    const uint32_t hi22 = (p0.digit << 11) | p1.digit;
#pragma unroll
    for (int j = 0; j < MAXI; ++j)
      if (j < ipt && base + j < total && (key[j] >> 10) == hi22) atomicAdd(&hist[2 * kBins + (key[j] & 0x3ffu)], 1u);
    grid.sync();
    const Pick p2 = block_pick(hist + 2 * kBins, k_rem, wsum, &pick_s);
    k_rem -= p2.above;
    thr = (hi22 << 10) | p2.digit;
    n_ties = p2.at;
  }
  // ---- ties at the threshold: all of them are taken unless the key occurs more often than needed; then the lowest
  //      node indices win (block counts + grid barrier + in-block ranks: deterministic)
  uint32_t before = 0;
  const bool ordered = k_rem < n_ties;
  if (ordered) {
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < MAXI; ++j) mine += (j < ipt && base + j < total && key[j] == thr);
    uint32_t blk;
    const uint32_t rank = sel_block_exclusive(mine, wsum, &blk);
    if (t == 0) tile_ties[blockIdx.x] = blk;
    grid.sync();
    uint32_t part = 0;
    for (int i = t; i < (int)blockIdx.x; i += kSelThreads) part += __ldcg(tile_ties + i);
    uint32_t prev;
    (void)sel_block_exclusive(part, wsum, &prev);
    before = prev + rank;
  }
  int32_t n0 = 0, n1 = 0, n2 = 0, nbv = 0;
#pragma unroll
  for (int j = 0; j < MAXI; ++j) {
    const int64_t i = base + j;
    if (j < ipt && i < total) {
      bool sel = key[j] > thr;
      if (key[j] == thr) { sel = !ordered || before < k_rem; ++before; }
      const int s = sel ? 1 : (int)((side2 >> j) & 1u) * 2;
      status[i] = (OutT)s;
      n0 += (s == 0); n1 += (s == 1); n2 += (s == 2); nbv += (s == 1 && i >= m);
    }
  }
  if (counts) {
    if (n0) atomicAdd(&c_s[0], n0);
    if (n1) atomicAdd(&c_s[1], n1);
    if (n2) atomicAdd(&c_s[2], n2);
    if (nbv) atomicAdd(&c_s[3], nbv);
    __syncthreads();
    if (t < 4 && c_s[t]) atomicAdd(&counts[t], c_s[t]);
  }
}

template <int MAXI, typename OutT>
int launch_fused(const float* lc, int32_t m, const float* lv, int32_t n, uint32_t k, int ipt, int grid, uint32_t* hist,
                 uint32_t* ties, void* status, int32_t* counts, cudaStream_t st) {
  OutT* out = reinterpret_cast<OutT*>(status);
  void* args[] = {(void*)&lc, (void*)&m, (void*)&lv, (void*)&n, (void*)&k, (void*)&ipt, (void*)&hist, (void*)&ties, (void*)&out,
                  (void*)&counts};
  LPGNN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)select_fused_kernel<MAXI, OutT>, dim3(grid), dim3(kSelThreads), args, 0, st));
  return LPGNN_OK;
}

// co-resident blocks of the fused kernel on this device (smallest over the instantiations that may be picked)
int fused_capacity() {
  static int cap = -1;
  if (cap < 0) {
    int per_sm = 0, coop = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_fused_kernel<16, int64_t>, kSelThreads, 0) != cudaSuccess)
      per_sm = 0;
    cudaGetLastError();
    cap = per_sm * sm_count();
  }
  return cap;
}

// 0 forces the multi-launch chain (lpgnn_set_select_fused; environment LPGNN_SELECT_FUSED=0 for A/B runs)
int g_select_fused = [] { const char* e = getenv("LPGNN_SELECT_FUSED"); return e ? atoi(e) != 0 : 1; }();

struct Layout {
  size_t keys, side, hist, state, ties, fhist, fties, total;
};
Layout layout(int64_t total_nodes) {
  Layout L;
  const size_t t = (size_t)(total_nodes > 0 ? total_nodes : 1);
  size_t off = 0;
  L.keys = off; off += align_up(t * 4, 256);
  L.side = off; off += align_up(t, 256);
  L.hist = off; off += 4 * 256 * 4;
  L.state = off; off += 256;
  L.ties = off; off += align_up(((t + kTile - 1) / kTile + 1) * 4, 256);
  L.fhist = off; off += 3 * kBins * 4;                 // one-launch variant: three digit histograms ...
  L.fties = off; off += 4096 * 4;                      // ... and per-block tie counts (grid <= 4096 blocks)
  L.total = off;
  return L;
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" size_t lpgnn_basis_select_workspace_bytes(int64_t total_nodes) { return layout(total_nodes).total; }

extern "C" int lpgnn_basis_select(const float* logits_cons, int32_t m, const float* logits_vars, int32_t n,
                                  int32_t k_basic, void* status, int status_is_i64, int32_t* counts_out,
                                  void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m >= 0 && n >= 0, "basis_select: negative size");
  const int64_t total = (int64_t)m + n;
  LPGNN_REQUIRE(k_basic >= 0 && k_basic <= total, "basis_select: k=%d outside [0, m+n=%lld]", k_basic, (long long)total);
  const Layout L = layout(total);
  if (workspace_bytes < L.total) {
    set_error("basis_select: workspace %zu < required %zu", workspace_bytes, L.total);
    return LPGNN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (total == 0) {
    if (counts_out) LPGNN_CUDA_OK(cudaMemsetAsync(counts_out, 0, 4 * sizeof(int32_t), st));
    return LPGNN_OK;
  }
  LPGNN_REQUIRE(status && workspace && (m == 0 || logits_cons) && (n == 0 || logits_vars), "basis_select: null pointer");
  char* w = reinterpret_cast<char*>(workspace);
  uint32_t* keys = reinterpret_cast<uint32_t*>(w + L.keys);
  uint8_t* side = reinterpret_cast<uint8_t*>(w + L.side);
  uint32_t* hist = reinterpret_cast<uint32_t*>(w + L.hist);
  void* state = w + L.state;
  uint32_t* ties = reinterpret_cast<uint32_t*>(w + L.ties);

  const uint32_t k = (uint32_t)k_basic;
  // ---- one cooperative launch when the nodes fit the co-resident grid with <= 16 keys per thread
  const int cap = g_select_fused ? min(fused_capacity(), 4096) : 0;
  if (cap > 0 && total <= (int64_t)cap * kSelThreads * 16) {
    uint32_t* fhist = reinterpret_cast<uint32_t*>(w + L.fhist);
    uint32_t* fties = reinterpret_cast<uint32_t*>(w + L.fties);
    const int want = ceil_div(total, kSelThreads);
    const int grid = want < cap ? want : cap;
    const int ipt = ceil_div(total, (int64_t)grid * kSelThreads);
    LPGNN_CUDA_OK(cudaMemsetAsync(fhist, 0, 3 * kBins * 4, st));
    if (counts_out) LPGNN_CUDA_OK(cudaMemsetAsync(counts_out, 0, 4 * sizeof(int32_t), st));
    int rc;
    if (status_is_i64)
      rc = ipt <= 1 ? launch_fused<1, int64_t>(logits_cons, m, logits_vars, n, k, ipt, grid, fhist, fties, status, counts_out, st)
         : ipt <= 4 ? launch_fused<4, int64_t>(logits_cons, m, logits_vars, n, k, ipt, grid, fhist, fties, status, counts_out, st)
                    : launch_fused<16, int64_t>(logits_cons, m, logits_vars, n, k, ipt, grid, fhist, fties, status, counts_out, st);
    else
      rc = ipt <= 1 ? launch_fused<1, uint8_t>(logits_cons, m, logits_vars, n, k, ipt, grid, fhist, fties, status, counts_out, st)
         : ipt <= 4 ? launch_fused<4, uint8_t>(logits_cons, m, logits_vars, n, k, ipt, grid, fhist, fties, status, counts_out, st)
                    : launch_fused<16, uint8_t>(logits_cons, m, logits_vars, n, k, ipt, grid, fhist, fties, status, counts_out, st);
    if (rc) return rc;
    count_launches(1);
    return LPGNN_OK;
  }
  const int grid_stride = min(ceil_div(total, kThreads), sm_count() * 8);
  const int ntiles = ceil_div(total, kTile);
  (void)state;
  init_hist_kernel<<<1, 256, 0, st>>>(hist, counts_out);
  softmax_key_kernel<<<grid_stride, kThreads, 0, st>>>(logits_cons, m, logits_vars, n, keys, side, hist);
  for (int pass = 1; pass < 4; ++pass)
    digit_hist_kernel<<<grid_stride, kThreads, 0, st>>>(keys, total, hist, pass, k);
  tie_count_kernel<<<ntiles, kThreads, 0, st>>>(keys, total, hist, k, ties);
  if (status_is_i64)
    status_kernel<int64_t><<<ntiles, kThreads, 0, st>>>(keys, side, total, m, hist, k, ties,
                                                        reinterpret_cast<int64_t*>(status), counts_out);
  else
    status_kernel<uint8_t><<<ntiles, kThreads, 0, st>>>(keys, side, total, m, hist, k, ties,
                                                        reinterpret_cast<uint8_t*>(status), counts_out);
  LPGNN_LAUNCH_OK();
  count_launches(7);
  return LPGNN_OK;
}

extern "C" int lpgnn_basis_select_segmented(const float* logits_cons, const float* logits_vars, const int32_t* cons_ptr,
                                            const int32_t* vars_ptr, int32_t n_segments, int32_t total_cons,
                                            int32_t total_vars, void* status, int status_is_i64, void* workspace,
                                            size_t workspace_bytes, lpgnn_stream_t stream) {
  return lpgnn_basis_select_segmented_ex(logits_cons, logits_vars, cons_ptr, vars_ptr, n_segments, total_cons, total_vars, status,
                                         status_is_i64, 0, workspace, workspace_bytes, stream);
}

extern "C" int lpgnn_basis_select_segmented_ex(const float* logits_cons, const float* logits_vars, const int32_t* cons_ptr,
                                               const int32_t* vars_ptr, int32_t n_segments, int32_t total_cons,
                                               int32_t total_vars, void* status, int status_is_i64, int lp_major,
                                               void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(n_segments >= 0 && total_cons >= 0 && total_vars >= 0, "basis_select_segmented: negative size");
  if (n_segments == 0) return LPGNN_OK;
  const int64_t total = (int64_t)total_cons + total_vars;
  const Layout L = layout(total);
  if (workspace_bytes < L.total) { set_error("basis_select_segmented: workspace too small"); return LPGNN_EWORKSPACE; }
  LPGNN_REQUIRE(logits_cons && logits_vars && cons_ptr && vars_ptr && status && workspace,
                "basis_select_segmented: null pointer");
  char* w = reinterpret_cast<char*>(workspace);
  uint32_t* keys = reinterpret_cast<uint32_t*>(w + L.keys);
  uint8_t* side = reinterpret_cast<uint8_t*>(w + L.side);
  cudaStream_t st = (cudaStream_t)stream;
  if (status_is_i64)
    segmented_select_kernel<int64_t><<<n_segments, kSegThreads, 0, st>>>(logits_cons, logits_vars, cons_ptr, vars_ptr,
                                                                         total_cons, keys, side,
                                                                         reinterpret_cast<int64_t*>(status), lp_major);
  else
    segmented_select_kernel<uint8_t><<<n_segments, kSegThreads, 0, st>>>(logits_cons, logits_vars, cons_ptr, vars_ptr,
                                                                         total_cons, keys, side,
                                                                         reinterpret_cast<uint8_t*>(status), lp_major);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

// Tuning knob: 1 (default) = the one-launch cooperative kernel where it fits, 0 = always the multi-launch chain.  Both
// produce the same statuses (same keys, same threshold, same ascending-index tie rule).  Returns the previous setting.
extern "C" int lpgnn_set_select_fused(int enable) {
  const int prev = lpgnn::g_select_fused;
  lpgnn::g_select_fused = enable ? 1 : 0;
  return prev;
}
