// (f-4) Sampled-subgraph path for LPs above `edge_num_thresh`: neighbour sampling and induced-subgraph extraction
// on the device, over the SAME resident CSR / CSC structures the full-graph path uses.
//
// Replaces torch_geometric.loader.NeighborLoader as the reference drives it (train.py:107-116: num_neighbors =
// [6]*depth, directed=False; val.py:22-27: num_neighbors = [-1]*depth) followed by MyToBipartite (dataset.py:275-332).
// The unipartite LP graph is bipartite, so one hop from a set of constraints is a walk over CSR rows and one hop from
// a set of variables a walk over CSC rows:
//   lpgnn_sample_mark        for every frontier node, choose min(deg, fanout) of its neighbours without replacement
//                            (Knuth's selection sampling driven by a counter hash: unbiased, reproducible from
//                            (seed, hop, node); fanout < 0 keeps all) and set their flag on the other side
//   lpgnn_induced_count/fill the subgraph induced by the sampled node sets (directed=False: every edge between two
//                            sampled nodes), relabelled to local ids, as a COO in local row order
// Flag writes are idempotent stores of 1, the extraction owns one output segment per row: no atomics, deterministic.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
sample_mark_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int32_t* __restrict__ frontier,
                   int32_t n_frontier, int32_t fanout, uint64_t seed, uint8_t* __restrict__ marks) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_frontier) return;
  const int32_t f = frontier[t];
  const int32_t beg = ptr[f], deg = ptr[f + 1] - beg;
  if (fanout < 0 || deg <= fanout) {
    for (int e = 0; e < deg; ++e) marks[idx[beg + e]] = 1;
    return;
  }
  // selection sampling (Knuth 3.4.2 S): entry e is taken with probability needed / (deg - e)
  int needed = fanout;
  const uint64_t key = seed ^ ((uint64_t)(uint32_t)f * 0x9E3779B97F4A7C15ull);
  for (int e = 0; e < deg && needed > 0; ++e) {
    const uint32_t u = hash_u64(key + (uint64_t)e * 0xD6E8FEB86659FD93ull);
    if ((uint64_t)u * (uint32_t)(deg - e) < ((uint64_t)needed << 32)) {
      marks[idx[beg + e]] = 1;
      --needed;
    }
  }
}

// rows: global ids of the sampled nodes of this side, in LOCAL order.  map_other[j] = local id of node j of the
// other side, or -1 when j is not sampled.
__global__ void __launch_bounds__(kThreads)
induced_count_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int32_t* __restrict__ rows,
                     int32_t n_rows, const int32_t* __restrict__ map_other, int32_t* __restrict__ counts) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_rows) return;
  const int32_t r = rows[t];
  int c = 0;
  for (int32_t e = ptr[r]; e < ptr[r + 1]; ++e) c += map_other[idx[e]] >= 0;
  counts[t] = c;
}

__global__ void __launch_bounds__(kThreads)
induced_fill_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                    const int32_t* __restrict__ rows, int32_t n_rows, const int32_t* __restrict__ map_other,
                    const int64_t* __restrict__ offsets, int32_t* __restrict__ out_row, int32_t* __restrict__ out_col,
                    float* __restrict__ out_val) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_rows) return;
  const int32_t r = rows[t];
  int64_t o = offsets[t];
  for (int32_t e = ptr[r]; e < ptr[r + 1]; ++e) {
    const int32_t j = map_other[idx[e]];
    if (j >= 0) { out_row[o] = t; out_col[o] = j; out_val[o] = val[e]; ++o; }
  }
}


// ------------------------------------------------------------------------------------------------ one-sync mini-batches
// The node sets of a mini-batch without host round trips: every list lives at full capacity on the device, its fill
// level in a small device array, and the ordered compactions (seeds in seed order, then the nodes first reached at hop
// 1, 2, ... in ascending id order -- the local numbering MyToBipartite would produce, dataset.py:288-294) are stable
// stream compactions whose cross-tile prefix comes from a decoupled look-back: tiles are handed out by a ticket
// counter, a tile publishes its count (flag A) as soon as it knows it and its inclusive prefix (flag P) once it has
// summed its predecessors' words, so no tile ever waits for a tile that has not started.
//   state byte per node: 0 = untouched, t in 1..254 = first marked during hop t, 255 = member of the node list
constexpr int kScanItems = 16;                       // state bytes per thread = one 16-byte load
constexpr int kScanTile = kThreads * kScanItems;
constexpr unsigned long long kFlagA = 1ull << 62, kFlagP = 2ull << 62, kValMask = (1ull << 62) - 1;

struct ScanWs {                                      // zeroed (cudaMemsetAsync) before every scanning kernel
  unsigned int ticket;
  unsigned int pad;
  unsigned long long tile[1];                        // [ntiles]
};

// Block-wide: exclusive prefix of `aggregate` over all tiles before `tile` (ticket order).  Called by every thread.
__device__ __forceinline__ uint32_t lookback_prefix(ScanWs* ws, int tile, uint32_t aggregate) {
  __shared__ uint32_t prefix_s;
  if (threadIdx.x == 0) {
    volatile unsigned long long* words = ws->tile;
    uint32_t running = 0;
    if (tile > 0) {
      words[tile] = kFlagA | aggregate;
      __threadfence();
      for (int j = tile - 1; j >= 0; --j) {
        unsigned long long w;
        do { w = words[j]; } while ((w >> 62) == 0);
        running += (uint32_t)(w & kValMask);
        if ((w >> 62) == 2) break;
      }
    }
    words[tile] = kFlagP | (unsigned long long)(running + aggregate);
    __threadfence();
    prefix_s = running;
  }
  __syncthreads();
  return prefix_s;
}

__device__ __forceinline__ int take_ticket(ScanWs* ws) {
  __shared__ int tile_s;
  if (threadIdx.x == 0) tile_s = (int)atomicAdd(&ws->ticket, 1u);
  __syncthreads();
  return tile_s;
}

// exclusive prefix of `v` over the block in thread order + block total (kThreads = 256)
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) { const uint32_t s = wsum[w]; if (w < warp) before += s; all += s; }
  __syncthreads();
  *total = all;
  return before + incl - v;
}

// sizes (device int32): [0..1] = members per side so far (cons, vars), [2..3] = seeds per side,
// [4 + 2*h + side] = members after hop h (h = 0 is "after the seeds"), [kSizesNnz] = nnz of the induced subgraph
constexpr int kSizesHop0 = 4;
constexpr int kSizesNnz = 4 + 2 * 34;
constexpr int kSizesLen = kSizesNnz + 2;

// Seeds of one side, in seed order: list[k] = side-local id, map[id] = k, state[id] = 255.
__global__ void __launch_bounds__(kThreads)
seed_partition_kernel(const int64_t* __restrict__ seeds, int n_seeds, int m, int side, int32_t* __restrict__ list,
                      int32_t* __restrict__ map, uint8_t* __restrict__ state, int32_t* __restrict__ sizes, ScanWs* ws) {
  const int tile = take_ticket(ws);
  const int i = tile * kThreads + threadIdx.x;
  int64_t s = -1;
  bool mine = false;
  if (i < n_seeds) { s = seeds[i]; mine = side ? (s >= m) : (s < m); }
  uint32_t total;
  const uint32_t rank = block_exclusive(mine ? 1u : 0u, &total);
  const uint32_t prefix = lookback_prefix(ws, tile, total);
  if (mine) {
    const int32_t id = (int32_t)(side ? s - m : s);
    const int32_t k = (int32_t)(prefix + rank);
    list[k] = id; map[id] = k; state[id] = 255;
  }
  if (tile == (int)gridDim.x - 1 && threadIdx.x == 0) {
    sizes[side] = sizes[2 + side] = sizes[kSizesHop0 + side] = (int32_t)(prefix + total);
  }
}

// Neighbour sampling from the frontier of hop `hop` (the nodes appended by hop - 1; hop = 1 walks the seeds): marks the
// other side's untouched nodes with the hop number.
__global__ void __launch_bounds__(kThreads)
sample_hop_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int32_t* __restrict__ list,
                  const int32_t* __restrict__ sizes, int side, int hop, int32_t fanout, uint64_t seed,
                  uint8_t* __restrict__ state_other) {
  const int beg = hop >= 2 ? sizes[kSizesHop0 + 2 * (hop - 2) + side] : 0;
  const int end = sizes[kSizesHop0 + 2 * (hop - 1) + side];
  const uint8_t tag = (uint8_t)hop;
  for (int t = beg + blockIdx.x * kThreads + threadIdx.x; t < end; t += gridDim.x * kThreads) {
    const int32_t f = list[t];
    const int32_t b = ptr[f], deg = ptr[f + 1] - b;
    if (fanout < 0 || deg <= fanout) {
      for (int e = 0; e < deg; ++e) { const int32_t j = idx[b + e]; if (state_other[j] == 0) state_other[j] = tag; }
      continue;
    }
    int needed = fanout;
    const uint64_t key = seed ^ ((uint64_t)(uint32_t)f * 0x9E3779B97F4A7C15ull);
    for (int e = 0; e < deg && needed > 0; ++e) {
      const uint32_t u = hash_u64(key + (uint64_t)e * 0xD6E8FEB86659FD93ull);
      if ((uint64_t)u * (uint32_t)(deg - e) < ((uint64_t)needed << 32)) {
        const int32_t j = idx[b + e];
        if (state_other[j] == 0) state_other[j] = tag;     // racing writers store the same tag
        --needed;
      }
    }
  }
}

// Appends the nodes first marked during hop `hop` to the side's list in ascending id order.
__global__ void __launch_bounds__(kThreads)
append_marked_kernel(uint8_t* __restrict__ state, int n_nodes, int hop, int side, int32_t* __restrict__ list,
                     int32_t* __restrict__ map, int32_t* __restrict__ sizes, ScanWs* ws) {
  const int tile = take_ticket(ws);
  const int old = sizes[kSizesHop0 + 2 * (hop - 1) + side];      // written by an earlier kernel
  const int base = tile * kScanTile + threadIdx.x * kScanItems;
  const uint32_t tag = (uint32_t)hop;
  uint32_t st[kScanItems];
  if (base + kScanItems <= n_nodes) {                          // one 16-byte load (state is 256-byte aligned)
    const uint4 v = *reinterpret_cast<const uint4*>(state + base);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) st[j] = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
  } else {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) st[j] = (base + j < n_nodes) ? state[base + j] : 0u;
  }
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) cnt += st[j] == tag;
  uint32_t total;
  const uint32_t rank = block_exclusive(cnt, &total);
  const uint32_t prefix = lookback_prefix(ws, tile, total);
  if (cnt) {
    int32_t k = old + (int32_t)(prefix + rank);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
      if (st[j] == tag) { list[k] = base + j; map[base + j] = k; state[base + j] = 255; ++k; }
  }
  if (tile == (int)gridDim.x - 1 && threadIdx.x == 0)
    sizes[side] = sizes[kSizesHop0 + 2 * hop + side] = old + (int32_t)(prefix + total);
}

// offsets[t] = number of induced entries in the rows before local row t; offsets[n_rows] and sizes[kSizesNnz] = total
__global__ void __launch_bounds__(kThreads)
induced_offsets_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int32_t* __restrict__ rows,
                       const int32_t* __restrict__ sizes, const int32_t* __restrict__ map_other,
                       int32_t* __restrict__ offsets, int32_t* __restrict__ sizes_out, ScanWs* ws) {
  const int n_rows = sizes[0];
  const int ntiles = (n_rows + kThreads - 1) / kThreads;
  const int tile = take_ticket(ws);
  if (tile >= ntiles) {                                   // (n_rows == 0: tile 0 still reports an empty subgraph)
    if (tile == 0 && threadIdx.x == 0) { offsets[0] = 0; sizes_out[kSizesNnz] = 0; }
    return;
  }
  const int t = tile * kThreads + threadIdx.x;
  uint32_t c = 0;
  if (t < n_rows) {
    const int32_t r = rows[t];
    for (int32_t e = ptr[r]; e < ptr[r + 1]; ++e) c += map_other[idx[e]] >= 0;
  }
  uint32_t total;
  const uint32_t rank = block_exclusive(c, &total);
  const uint32_t prefix = lookback_prefix(ws, tile, total);
  if (t < n_rows) offsets[t] = (int32_t)(prefix + rank);
  if (tile == ntiles - 1 && threadIdx.x == 0) { offsets[n_rows] = (int32_t)(prefix + total); sizes_out[kSizesNnz] = (int32_t)(prefix + total); }
}

// COO of the induced subgraph in CANONICAL order (local row, then ascending local column), so that lpgnn_graph_build
// takes its sorted path: one thread per row gathers the surviving entries and orders them by local column (rows of an
// LP are short; the local numbering is not monotone in the global one, so the order of the full matrix's row is lost).
__global__ void __launch_bounds__(kThreads)
induced_fill_sorted_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                           const int32_t* __restrict__ rows, int32_t n_rows, const int32_t* __restrict__ map_other,
                           const int32_t* __restrict__ offsets, int32_t* __restrict__ out_row, int32_t* __restrict__ out_col,
                           float* __restrict__ out_val) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_rows) return;
  const int32_t r = rows[t];
  const int32_t o0 = offsets[t];
  int32_t o = o0;
  for (int32_t e = ptr[r]; e < ptr[r + 1]; ++e) {
    const int32_t j = map_other[idx[e]];
    if (j < 0) continue;
    // insertion into the sorted prefix [o0, o): duplicates of a column keep their input order (stable)
    const float v = val[e];
    int32_t p = o;
    while (p > o0 && out_col[p - 1] > j) { out_col[p] = out_col[p - 1]; out_val[p] = out_val[p - 1]; --p; }
    out_col[p] = j; out_val[p] = v; out_row[o] = t;
    ++o;
  }
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" int lpgnn_sample_mark(const int32_t* ptr, const int32_t* idx, const int32_t* frontier, int32_t n_frontier,
                                 int32_t fanout, uint64_t seed, uint8_t* marks_other, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(n_frontier >= 0, "sample_mark: bad frontier size %d", n_frontier);
  if (n_frontier == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && idx && frontier && marks_other, "sample_mark: null pointer");
  LPGNN_REQUIRE(fanout != 0, "sample_mark: fanout must be positive, or negative for all neighbours");
  sample_mark_kernel<<<ceil_div(n_frontier, kThreads), kThreads, 0, (cudaStream_t)stream>>>(ptr, idx, frontier, n_frontier,
                                                                                          fanout, seed, marks_other);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_induced_count(const int32_t* ptr, const int32_t* idx, const int32_t* rows, int32_t n_rows,
                                   const int32_t* map_other, int32_t* counts, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(n_rows >= 0, "induced_count: bad row count %d", n_rows);
  if (n_rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && idx && rows && map_other && counts, "induced_count: null pointer");
  induced_count_kernel<<<ceil_div(n_rows, kThreads), kThreads, 0, (cudaStream_t)stream>>>(ptr, idx, rows, n_rows, map_other,
                                                                                        counts);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_induced_fill(const int32_t* ptr, const int32_t* idx, const float* val, const int32_t* rows,
                                  int32_t n_rows, const int32_t* map_other, const int64_t* offsets, int32_t* out_row,
                                  int32_t* out_col, float* out_val, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(n_rows >= 0, "induced_fill: bad row count %d", n_rows);
  if (n_rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && idx && val && rows && map_other && offsets && out_row && out_col && out_val, "induced_fill: null pointer");
  induced_fill_kernel<<<ceil_div(n_rows, kThreads), kThreads, 0, (cudaStream_t)stream>>>(ptr, idx, val, rows, n_rows, map_other,
                                                                                       offsets, out_row, out_col, out_val);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

// ---------------------------------------------------------------------------------------------- one-sync mini-batches
namespace {
size_t scan_ws_bytes(int64_t items, int per_tile) {
  return align_up(sizeof(ScanWs) + (size_t)((items + per_tile - 1) / per_tile + 1) * sizeof(unsigned long long), 256);
}
}  // namespace

extern "C" int32_t lpgnn_sample_sizes_len(void) { return kSizesLen; }

extern "C" size_t lpgnn_sample_nodes_workspace_bytes(int32_t m, int32_t n, int32_t n_seeds) {
  const size_t a = scan_ws_bytes(n_seeds, kThreads), b = scan_ws_bytes(m > n ? m : n, kScanTile), c = scan_ws_bytes(m, kThreads);
  const size_t scan = a > b ? (a > c ? a : c) : (b > c ? b : c);
  return 2 * scan + align_up((size_t)m, 256) + align_up((size_t)n, 256);       // two scan areas + the two state arrays
}

extern "C" int lpgnn_sample_nodes(const int32_t* rowptr, const int32_t* col, const int32_t* colptr, const int32_t* row_csc,
                                  int32_t m, int32_t n, const int64_t* seeds, int32_t n_seeds, const int32_t* fanouts_host,
                                  int32_t n_hops, uint64_t seed, int32_t* cons_nodes, int32_t* var_nodes, int32_t* map_cons,
                                  int32_t* map_vars, int32_t* sizes, void* workspace, size_t workspace_bytes,
                                  lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(m > 0 && n > 0 && n_seeds > 0 && n_hops >= 0 && n_hops <= 32, "sample_nodes: bad sizes (hops <= 32)");
  LPGNN_REQUIRE(rowptr && col && colptr && row_csc && seeds && cons_nodes && var_nodes && map_cons && map_vars && sizes && workspace &&
                (n_hops == 0 || fanouts_host), "sample_nodes: null pointer");
  for (int h = 0; h < n_hops; ++h) LPGNN_REQUIRE(fanouts_host[h] != 0, "sample_nodes: fanout must be positive, or negative for all neighbours");
  if (workspace_bytes < lpgnn_sample_nodes_workspace_bytes(m, n, n_seeds)) { set_error("sample_nodes: workspace too small"); return LPGNN_EWORKSPACE; }
  LPGNN_REQUIRE((uintptr_t)workspace % 256 == 0, "sample_nodes: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* w = reinterpret_cast<char*>(workspace);
  const size_t scan = (lpgnn_sample_nodes_workspace_bytes(m, n, n_seeds) - align_up((size_t)m, 256) - align_up((size_t)n, 256)) / 2;
  ScanWs* ws[2] = {reinterpret_cast<ScanWs*>(w), reinterpret_cast<ScanWs*>(w + scan)};
  uint8_t* state_c = reinterpret_cast<uint8_t*>(w + 2 * scan);
  uint8_t* state_v = state_c + align_up((size_t)m, 256);
  LPGNN_CUDA_OK(cudaMemsetAsync(state_c, 0, align_up((size_t)m, 256) + align_up((size_t)n, 256), st));
  LPGNN_CUDA_OK(cudaMemsetAsync(map_cons, 0xff, sizeof(int32_t) * (size_t)m, st));
  LPGNN_CUDA_OK(cudaMemsetAsync(map_vars, 0xff, sizeof(int32_t) * (size_t)n, st));
  LPGNN_CUDA_OK(cudaMemsetAsync(sizes, 0, sizeof(int32_t) * kSizesLen, st));
  int launches = 0;
  const int seed_tiles = ceil_div(n_seeds, kThreads);
  for (int side = 0; side < 2; ++side) {
    LPGNN_CUDA_OK(cudaMemsetAsync(ws[side], 0, scan_ws_bytes(n_seeds, kThreads), st));
    seed_partition_kernel<<<seed_tiles, kThreads, 0, st>>>(seeds, n_seeds, m, side, side ? var_nodes : cons_nodes,
                                                            side ? map_vars : map_cons, side ? state_v : state_c, sizes, ws[side]);
    ++launches;
  }
  const int walk_grid = sm_count() * 4;
  for (int hop = 1; hop <= n_hops; ++hop) {
    const uint64_t s = seed + (uint64_t)hop * 0xC2B2AE3D27D4EB4Full;
    // a hop from the constraints walks CSR rows and marks variables; a hop from the variables walks CSC rows
    sample_hop_kernel<<<walk_grid, kThreads, 0, st>>>(rowptr, col, cons_nodes, sizes, 0, hop, fanouts_host[hop - 1], s, state_v);
    sample_hop_kernel<<<walk_grid, kThreads, 0, st>>>(colptr, row_csc, var_nodes, sizes, 1, hop, fanouts_host[hop - 1],
                                                      s ^ 0x5555555555555555ull, state_c);
    LPGNN_CUDA_OK(cudaMemsetAsync(ws[0], 0, scan_ws_bytes(m, kScanTile), st));
    LPGNN_CUDA_OK(cudaMemsetAsync(ws[1], 0, scan_ws_bytes(n, kScanTile), st));
    append_marked_kernel<<<ceil_div(m, kScanTile), kThreads, 0, st>>>(state_c, m, hop, 0, cons_nodes, map_cons, sizes, ws[0]);
    append_marked_kernel<<<ceil_div(n, kScanTile), kThreads, 0, st>>>(state_v, n, hop, 1, var_nodes, map_vars, sizes, ws[1]);
    launches += 4;
  }
  LPGNN_LAUNCH_OK();
  count_launches(launches);
  return LPGNN_OK;
}

extern "C" int lpgnn_induced_offsets(const int32_t* ptr, const int32_t* idx, const int32_t* rows, int32_t rows_capacity,
                                     const int32_t* map_other, int32_t* offsets, int32_t* sizes, void* workspace,
                                     size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows_capacity > 0 && ptr && idx && rows && map_other && offsets && sizes && workspace, "induced_offsets: bad arguments");
  LPGNN_REQUIRE(workspace_bytes >= scan_ws_bytes(rows_capacity, kThreads) && (uintptr_t)workspace % 256 == 0,
                "induced_offsets: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  ScanWs* ws = reinterpret_cast<ScanWs*>(workspace);
  LPGNN_CUDA_OK(cudaMemsetAsync(ws, 0, scan_ws_bytes(rows_capacity, kThreads), st));
  induced_offsets_kernel<<<ceil_div(rows_capacity, kThreads), kThreads, 0, st>>>(ptr, idx, rows, sizes, map_other, offsets, sizes, ws);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_induced_fill_sorted(const int32_t* ptr, const int32_t* idx, const float* val, const int32_t* rows,
                                         int32_t n_rows, const int32_t* map_other, const int32_t* offsets, int32_t* out_row,
                                         int32_t* out_col, float* out_val, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(n_rows >= 0, "induced_fill_sorted: bad row count %d", n_rows);
  if (n_rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(ptr && idx && val && rows && map_other && offsets && out_row && out_col && out_val, "induced_fill_sorted: null pointer");
  induced_fill_sorted_kernel<<<ceil_div(n_rows, kThreads), kThreads, 0, (cudaStream_t)stream>>>(ptr, idx, val, rows, n_rows, map_other,
                                                                                              offsets, out_row, out_col, out_val);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

// ---- node data of a sampled mini-batch in one launch: feature rows, labels and ids of the sampled constraints and variables
// (train.py:117-123 reads batch.x_s / x_t / y_s / y_t; PyG's NeighborLoader slices them with n_id)
namespace lpgnn {
namespace {
__global__ void sample_gather_kernel(const float* __restrict__ x_s, const float* __restrict__ x_t, const int64_t* __restrict__ y_s,
                                     const int64_t* __restrict__ y_t, const int32_t* __restrict__ cons_nodes, int32_t mc,
                                     const int32_t* __restrict__ var_nodes, int32_t nv, int32_t p, int32_t q,
                                     float* __restrict__ out_xs, float* __restrict__ out_xt, int64_t* __restrict__ out_ys,
                                     int64_t* __restrict__ out_yt, int32_t* __restrict__ ids_s, int32_t* __restrict__ ids_t) {
  const int64_t a = (int64_t)mc * p, b = a + (int64_t)nv * q;
  const int64_t total = b + mc + nv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < a) {
      const int32_t r = (int32_t)(i / p), c = (int32_t)(i % p);
      out_xs[i] = __ldg(x_s + (int64_t)__ldg(cons_nodes + r) * p + c);
    } else if (i < b) {
      const int64_t k = i - a;
      const int32_t r = (int32_t)(k / q), c = (int32_t)(k % q);
      out_xt[k] = __ldg(x_t + (int64_t)__ldg(var_nodes + r) * q + c);
    } else if (i < b + mc) {
      const int32_t r = (int32_t)(i - b);
      const int32_t node = __ldg(cons_nodes + r);
      ids_s[r] = node;
      if (y_s) out_ys[r] = __ldg(y_s + node);
    } else {
      const int32_t r = (int32_t)(i - b - mc);
      const int32_t node = __ldg(var_nodes + r);
      ids_t[r] = node;
      if (y_t) out_yt[r] = __ldg(y_t + node);
    }
  }
}
}  // namespace
}  // namespace lpgnn

extern "C" int lpgnn_sample_gather(const float* x_s, const float* x_t, const int64_t* y_s, const int64_t* y_t,
                                   const int32_t* cons_nodes, int32_t mc, const int32_t* var_nodes, int32_t nv, int32_t p,
                                   int32_t q, float* out_xs, float* out_xt, int64_t* out_ys, int64_t* out_yt, int32_t* ids_s,
                                   int32_t* ids_t, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(mc >= 0 && nv >= 0 && p > 0 && q > 0, "sample_gather: bad sizes mc=%d nv=%d p=%d q=%d", mc, nv, p, q);
  if (mc + nv == 0) return LPGNN_OK;
  LPGNN_REQUIRE(x_s && x_t && cons_nodes && var_nodes && out_xs && out_xt && ids_s && ids_t, "sample_gather: null pointer");
  LPGNN_REQUIRE((y_s == nullptr) == (out_ys == nullptr) && (y_t == nullptr) == (out_yt == nullptr), "sample_gather: label pointers must come in pairs");
  const int64_t total = (int64_t)mc * p + (int64_t)nv * q + mc + nv;
  const int64_t want = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  lpgnn::sample_gather_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
      x_s, x_t, y_s, y_t, cons_nodes, mc, var_nodes, nv, p, q, out_xs, out_xt, out_ys, out_yt, ids_s, ids_t);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}
