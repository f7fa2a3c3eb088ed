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
