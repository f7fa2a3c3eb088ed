// (a6) Basis decision: softmax -> global top-m on P(basic) -> status per node.
//
// Replaces val.inference_gnn (reference val.py:106-124): F.softmax, isnan mask, topk(m) over all
// m+n nodes (ATen sort-based for large k), two scatter writes, argmax and two .item() host syncs,
// with a sync-free sequence of small kernels:
//   1. softmax per node (fp32) -> key = bit pattern of p1 (p1 >= 0, so unsigned order == float
//      order), side = 0 if p0 >= p2 else 2 (argmax over {0,2} picks the first maximum);
//      fused with the histogram of the top radix digit
//   2. 4-pass MSB radix select of the k-th largest key (256-bin histograms, integer atomics only)
//   3. ties at the threshold are taken in ascending node index (block counts + scan + in-block
//      ranks: deterministic; torch.topk leaves the tie order implementation-defined)
//   4. status = 1 if selected else side.
// HBM-bound integer/byte work: (m+n)*(12 + 4 + 1) bytes in pass 1, (m+n)*4 per select pass.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;                     // nodes per thread in the ordered passes
constexpr int kTile = kThreads * kItems;      // nodes per block in the ordered passes

struct SelectState {
  uint32_t prefix;    // decided high bits of the threshold key
  uint32_t mask;      // which bits are decided
  uint32_t k_rem;     // how many still to take among keys matching the prefix
  uint32_t pad;
};

__device__ __forceinline__ float nan_to_zero(float v) { return (v != v) ? 0.f : v; }

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

// One block.  Chooses the digit of the k-th largest key from hist, updates the state, clears hist.
__global__ void __launch_bounds__(256) pick_digit_kernel(uint32_t* __restrict__ hist, SelectState* __restrict__ st,
                                                         int shift) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = hist[threadIdx.x];
  hist[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t k = st->k_rem;
    if (k == 0) {  // nothing to select: threshold above every key
      st->prefix = 0xffffffffu; st->mask = 0xffffffffu;
    } else {
      uint32_t above = 0;
      int d = 255;
      for (; d > 0; --d) {
        if (above + h[d] >= k) break;
        above += h[d];
      }
      st->prefix |= ((uint32_t)d) << shift;
      st->mask |= 0xffu << shift;
      st->k_rem = k - above;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
digit_hist_kernel(const uint32_t* __restrict__ keys, int64_t total, const SelectState* __restrict__ st, int shift,
                  uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t prefix = st->prefix, mask = st->mask;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const uint32_t key = keys[i];
    if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 0xff], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void init_state_kernel(SelectState* st, uint32_t k, uint32_t* hist, int32_t* counts) {
  if (threadIdx.x == 0) { st->prefix = 0; st->mask = 0; st->k_rem = k; st->pad = 0; }
  hist[threadIdx.x] = 0;
  if (counts && threadIdx.x < 4) counts[threadIdx.x] = 0;
}

// number of keys equal to the threshold in each tile
__global__ void __launch_bounds__(kThreads)
tie_count_kernel(const uint32_t* __restrict__ keys, int64_t total, const SelectState* __restrict__ st,
                 uint32_t* __restrict__ tile_ties) {
  __shared__ uint32_t cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const uint32_t thr = st->prefix;
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j)
    if (base + j < total && keys[base + j] == thr) ++c;
  if (c) atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_ties[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) tie_scan_kernel(uint32_t* __restrict__ tile_ties, int ntiles) {
  // exclusive scan, single block, sequential chunks (ntiles is small: total/2048)
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const int per = (ntiles + 1023) / 1024;
  const int lo = t * per, hi = min(lo + per, ntiles);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += tile_ties[i];
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t v = (t >= off) ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = (t == 0) ? 0 : part[t - 1];
  for (int i = lo; i < hi; ++i) { uint32_t c = tile_ties[i]; tile_ties[i] = run; run += c; }
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads)
status_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ side, int64_t total, int32_t m,
              const SelectState* __restrict__ st, const uint32_t* __restrict__ tile_tie_offset,
              OutT* __restrict__ status, int32_t* __restrict__ counts) {
  __shared__ uint32_t warp_ties[kThreads / 32];
  __shared__ int32_t c_s[4];
  if (threadIdx.x < 4) c_s[threadIdx.x] = 0;
  const uint32_t thr = st->prefix, k_rem = st->k_rem;
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  uint32_t key[kItems];
  uint32_t my_ties = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    key[j] = (base + j < total) ? keys[base + j] : 0u;
    if (base + j < total && key[j] == thr) ++my_ties;
  }
  // exclusive prefix of my_ties over the block, in thread order (== node order)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = my_ties;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) warp_ties[warp] = incl;
  __syncthreads();
  uint32_t before = tile_tie_offset[blockIdx.x] + incl - my_ties;
  for (int w = 0; w < warp; ++w) before += warp_ties[w];
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

struct Layout {
  size_t keys, side, hist, state, ties, total;
};
Layout layout(int64_t total_nodes) {
  Layout L;
  const size_t t = (size_t)(total_nodes > 0 ? total_nodes : 1);
  size_t off = 0;
  L.keys = off; off += align_up(t * 4, 256);
  L.side = off; off += align_up(t, 256);
  L.hist = off; off += 256 * 4;
  L.state = off; off += 256;
  L.ties = off; off += align_up(((t + kTile - 1) / kTile + 1) * 4, 256);
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
  SelectState* state = reinterpret_cast<SelectState*>(w + L.state);
  uint32_t* ties = reinterpret_cast<uint32_t*>(w + L.ties);

  const int grid_stride = min(ceil_div(total, kThreads), sm_count() * 8);
  const int ntiles = ceil_div(total, kTile);
  init_state_kernel<<<1, 256, 0, st>>>(state, (uint32_t)k_basic, hist, counts_out);
  softmax_key_kernel<<<grid_stride, kThreads, 0, st>>>(logits_cons, m, logits_vars, n, keys, side, hist);
  pick_digit_kernel<<<1, 256, 0, st>>>(hist, state, 24);
  for (int shift = 16; shift >= 0; shift -= 8) {
    digit_hist_kernel<<<grid_stride, kThreads, 0, st>>>(keys, total, state, shift, hist);
    pick_digit_kernel<<<1, 256, 0, st>>>(hist, state, shift);
  }
  tie_count_kernel<<<ntiles, kThreads, 0, st>>>(keys, total, state, ties);
  tie_scan_kernel<<<1, 1024, 0, st>>>(ties, ntiles);
  if (status_is_i64)
    status_kernel<int64_t><<<ntiles, kThreads, 0, st>>>(keys, side, total, m, state, ties,
                                                        reinterpret_cast<int64_t*>(status), counts_out);
  else
    status_kernel<uint8_t><<<ntiles, kThreads, 0, st>>>(keys, side, total, m, state, ties,
                                                        reinterpret_cast<uint8_t*>(status), counts_out);
  LPGNN_LAUNCH_OK();
  count_launches(12);
  return LPGNN_OK;
}
