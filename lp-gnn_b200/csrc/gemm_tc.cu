// (a3, bf16 fast path) Node transform of a hidden GraphConv layer on the 5th-gen tensor cores:
//   out[M,N] = epi( A1[M,K1]*W1[N,K1]^T + A2[M,K2]*W2[N,K2]^T + bias[N] )     bf16 in/out, fp32 accumulate
//
// Replaces lin_rel(agg) + lin_root(x_dst) + relu_ of PyG GraphConv (reference arch.py:75-80, 188:
// two cuBLAS GEMMs, a bias add, an add and an elementwise kernel) with one persistent,
// warp-specialised tcgen05 kernel; the two (A,W) pairs are walked as ONE concatenated reduction so
// the accumulator never leaves TMEM between them.
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A (128 x 64) and W (BN x 64), 128-byte
//               swizzle, into a kStages-deep shared-memory ring (full/empty mbarriers)
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, kind::f16),
//               accumulators in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//               overlaps the main loop of tile i+1; tcgen05.commit releases ring slots
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld (32 lanes x 32 columns), + bias, ReLU, pack to bf16, 64-byte
//               row segments to global
// Tiles are ordered n-fastest so CTAs resident together share the same A row block through L2.
// Bound: tensor pipe (2*M*N*(K1+K2) flops against the measured bf16 peak); see DESIGN.md.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#include <type_traits>

namespace lpgnn {
namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // 64 bf16 = 128 bytes = one swizzle row
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;                         // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiWarp0 * 32 + kEpiThreads;  // 384

template <int BN, typename OutT, int kCl = 1> struct Cfg {
  static constexpr int kRowBytes = 32 * (int)sizeof(OutT);   // one staged piece: 32 columns of one row
  static constexpr int kStagingBytes = kEpiWarps * 32 * kRowBytes;
  static constexpr bool kWide = sizeof(OutT) == 4;   // fp32 output needs twice the staging space
  static constexpr int kStages = (kCl == 2) ? (kWide ? 5 : 6) : (BN == 256) ? (kWide ? 3 : 4) : (BN == 128 ? (kWide ? 5 : 6) : (kWide ? 6 : 8));
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / kCl) * BK * 2;       // a CTA pair keeps half of the W tile in each CTA
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // power of two >= 32 for BN in {64,128,256}
  static constexpr int kBarBytes = 5120;  // mbarriers + tmem ptr + bias slice (BN floats) + head weight slice (3*BN floats)
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kStagingBytes + 1024 /*align slack*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

// The reduction is a concatenation of up to kMaxSegs (A, W) operand pairs: 2 for a hidden GraphConv layer
// (lin_rel on the aggregate, lin_root on the node's own features), 6 or 12 for its fp32-accurate forms, where every
// fp32 operand is split into 2 or 3 bf16 parts and the significant cross products accumulate in the same TMEM tile.
constexpr int kMaxSegs = 12;
struct Segs {
  CUtensorMap a[kMaxSegs];
  CUtensorMap w[kMaxSegs];
  int kb_end[kMaxSegs];   // cumulative K-block count at the end of each segment
  int count;
};

// 0xffff per 16-bit lane whose bf16 value is > 0 (positive, non-zero), else 0
__device__ __forceinline__ uint32_t bf16x2_pos_mask(uint32_t a) {
  const uint32_t lo = ((a & 0x8000u) == 0u && (a & 0x7fffu) != 0u) ? 0x0000ffffu : 0u;
  const uint32_t hi = ((a & 0x80000000u) == 0u && (a & 0x7fff0000u) != 0u) ? 0xffff0000u : 0u;
  return lo | hi;
}

// kMN = true: both operands are MN-major (out = A^T B with A [Kred, M], B [Kred, N] row-major; the weight gradient
// dW = dY^T X without materialising any transpose): segment 0's maps describe A and B with 64 x 64 boxes.
// kCl = 2: CTA pair (tcgen05 cta_group::2).  The two CTAs of a cluster own two vertically adjacent 128-row blocks of
// the same column block; the leader (rank 0) issues ONE 256 x BN x 16 MMA per step that reads each CTA's A tile and
// each CTA's HALF of the W tile from their shared memories and accumulates into both TMEMs.  Per SM and K block the
// shared-memory traffic drops from 48 KB written + 48 KB read to 32 + 32 KB (one-CTA tiles are bound by exactly that:
// 96 + 96 B/clk against the 128 B/clk of an SM's shared memory), and the ring holds 6 stages instead of 4.
//   barriers: every TMA of the pair completes on the LEADER's full barrier; the leader's tcgen05.commit multicasts the
//   "slot free" / "accumulator ready" arrivals to both CTAs; both epilogues arrive on the leader's tmem_empty barrier.
// kEpi selects the epilogue features that are COMPILED IN (bit 0: fused basis-status head, bit 1: training extras --
// keep-mask, dropout, output scale).  Predicated-off code still costs issue slots, and with one or two K blocks per
// tile the kernel is bound by the epilogue's instruction stream, so the plain transform carries none of it.
template <int BN, typename OutT, bool kMN, int kCl, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ Segs segs, const float* __restrict__ bias, OutT* __restrict__ out, int M, int N,
               int relu, const float* __restrict__ head_w /*[3,N] or null*/,
               float* __restrict__ head_partial /*[N/BN][M][3] or null*/, int ksplit, const EpiX epx) {
  using C = Cfg<BN, OutT, kCl>;
  constexpr bool kHead = (kEpi & 1) != 0, kEpx = (kEpi & 2) != 0;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  // (aligned by OFFSET, not by rounding the pointer through an integer: that would turn every access to the staging area
  // and the epilogue constants into generic LD / ST -- long-scoreboard latency on each -- instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + C::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * C::kStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * C::kStages + 2;// [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * C::kStages + 6);  // [BN]
  float* headw_s = bias_s + BN;                                         // [3][BN]
  uint8_t* staging = smem + C::kStages * C::kStageBytes + C::kBarBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = N / BN;
  const int num_m = (M + BM - 1) / BM;
  // split-K (weight gradients: few output tiles, reduction over all nodes): work item = (output tile, K slice);
  // slice s writes its partial tile to out + s*M*N (fp32), summed afterwards in a fixed order.
  const int num_out_tiles = num_m * num_n;
  const int num_tiles = num_out_tiles * ksplit;
  const int kblocks_all = segs.kb_end[segs.count - 1];
  const int kb_per = (kblocks_all + ksplit - 1) / ksplit;
  // Work loop.  kCl = 1: CTA b takes tiles b, b + grid, ...  kCl = 2 (no split-K): cluster c takes the tile PAIRS
  // c, c + clusters, ...; pair p = rows blocks (2*(p / num_n), +1) of column block p % num_n, one per CTA rank.  A rank
  // whose row block lies beyond M still runs the loads and MMAs (TMA zero-fills, stores are masked): the two CTAs must
  // stay in lock step because they fill each other's shared memory.
  const uint32_t crank = (kCl > 1) ? ptx::cluster_ctarank() : 0u;
  const int w_first = (kCl > 1) ? (int)(blockIdx.x / kCl) : (int)blockIdx.x;
  const int w_step = (kCl > 1) ? (int)(gridDim.x / kCl) : (int)gridDim.x;
  const int num_pairs = ((num_m + kCl - 1) / kCl) * num_n;            // output tile pairs (kCl = 2)
  const int w_count = (kCl > 1) ? num_pairs * ksplit : num_tiles;
  auto tile_of = [&](int wi, int& m_blk, int& n_blk, int& split) {
    if (kCl > 1) { const int op = wi % num_pairs; split = wi / num_pairs; m_blk = kCl * (op / num_n) + (int)crank; n_blk = op % num_n; }
    else { const int ot = wi % num_out_tiles; split = wi / num_out_tiles; m_blk = ot / num_n; n_blk = ot % num_n; }
  };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < segs.count; ++i) { ptx::prefetch_tensormap(&segs.a[i]); ptx::prefetch_tensormap(&segs.w[i]); }
  }
  if (warp == 1 && lane == 0) {
    // pair: the leader's full barrier counts the bytes of both producers, its tmem_empty barrier hears both epilogues
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tmem_full[b], 1); ptx::mbar_init(&tmem_empty[b], kCl * kEpiThreads); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kCl > 1) { ptx::tmem_alloc_pair(tmem_ptr, C::kTmemCols); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(tmem_ptr, C::kTmemCols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCl > 1) ptx::cluster_sync();      // the peer's barriers exist before anything is sent to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = w_first; tile < w_count; tile += w_step) {
        int m_blk, n_blk, split;
        tile_of(tile, m_blk, n_blk, split);
        const int kb_begin = split * kb_per, kb_end = min(kb_begin + kb_per, kblocks_all);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (kCl > 1) {
            // both CTAs' tiles complete on the leader's barrier (address of my barrier mapped into CTA 0)
            // (the leader alone arrives, expecting the bytes of both CTAs: a remote arrive per K block would put a
            // cluster round trip on the peer producer's critical path; bytes that land before the leader's
            // expect_tx just run the transaction count negative until it is posted)
            const uint32_t lead_bar = ptx::mapa_shared(ptx::smem_u32(&full_bar[stage]), 0);
            if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], kCl * C::kStageBytes);
            if constexpr (kMN) {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                ptx::tma_load_2d_pair(smem_a + stage * C::kABytes + i * 8192, &segs.a[0], lead_bar, m_blk * BM + 64 * i, kb * BK);
#pragma unroll
              for (int i = 0; i < BN / kCl / 64; ++i)
                ptx::tma_load_2d_pair(smem_b + stage * C::kBBytes + i * 8192, &segs.w[0], lead_bar,
                                      n_blk * BN + (int)crank * (BN / kCl) + 64 * i, kb * BK);
            } else {
              int sg = 0;
              while (kb >= segs.kb_end[sg]) ++sg;
              const int kc = (kb - (sg ? segs.kb_end[sg - 1] : 0)) * BK;
              ptx::tma_load_2d_pair(smem_a + stage * C::kABytes, &segs.a[sg], lead_bar, kc, m_blk * BM);
              ptx::tma_load_2d_pair(smem_b + stage * C::kBBytes, &segs.w[sg], lead_bar, kc, n_blk * BN + (int)crank * (BN / kCl));
            }
          } else {
          ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          if constexpr (kMN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              ptx::tma_load_2d(smem_a + stage * C::kABytes + i * 8192, &segs.a[0], &full_bar[stage], m_blk * BM + 64 * i,
                               kb * BK);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              ptx::tma_load_2d(smem_b + stage * C::kBBytes + i * 8192, &segs.w[0], &full_bar[stage], n_blk * BN + 64 * i,
                               kb * BK);
          } else {
            int sg = 0;
            while (kb >= segs.kb_end[sg]) ++sg;
            const int kc = (kb - (sg ? segs.kb_end[sg - 1] : 0)) * BK;
            ptx::tma_load_2d(smem_a + stage * C::kABytes, &segs.a[sg], &full_bar[stage], kc, m_blk * BM);
            ptx::tma_load_2d(smem_b + stage * C::kBBytes, &segs.w[sg], &full_bar[stage], kc, n_blk * BN);
          }
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    if (lane == 0 && (kCl == 1 || crank == 0)) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM * kCl, BN, kMN, std::is_same<OutT, __half>::value);
      int stage = 0; uint32_t phase = 0;
      int t = 0;
      for (int tile = w_first; tile < w_count; tile += w_step, ++t) {
        const int buf = t & 1;
        const uint32_t use_phase = (t >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[buf], use_phase ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        int m_blk_u, n_blk_u, split;
        tile_of(tile, m_blk_u, n_blk_u, split);
        const int kb_begin = split * kb_per, kb_end = min(kb_begin + kb_per, kblocks_all);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + stage * C::kABytes);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * C::kBBytes);
          const uint64_t adesc = kMN ? ptx::umma_desc_mn_sw128(a_addr, 8192) : ptx::umma_desc_k_sw128(a_addr);
          const uint64_t bdesc = kMN ? ptx::umma_desc_mn_sw128(b_addr, 8192) : ptx::umma_desc_k_sw128(b_addr);
          // K-major: 16 elements = 32 bytes along the swizzle row (+2 in 16-byte units);
          // MN-major: 16 K-rows of 128 bytes = 2048 bytes (+128)
          constexpr uint32_t kstep = kMN ? 128 : 2;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            if constexpr (kCl > 1) ptx::umma_bf16_pair(tmem_d, adesc + kstep * k, bdesc + kstep * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            else ptx::umma_bf16(tmem_d, adesc + kstep * k, bdesc + kstep * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          if constexpr (kCl > 1) ptx::umma_commit_pair(&empty_bar[stage]);   // slot free in both CTAs
          else ptx::umma_commit(&empty_bar[stage]);  // slot free once these MMAs have read it
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (kCl > 1) ptx::umma_commit_pair(&tmem_full[buf]);    // both halves of the accumulator complete
        else ptx::umma_commit(&tmem_full[buf]);      // accumulator complete
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - kEpiWarp0;              // 0..7
    const int q = ew & 3;                         // TMEM lane quarter of this warp (== warp % 4)
    const int hsel = ew >> 2;                     // which half of the tile's columns this warp drains
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..255
    constexpr int kPieces = BN / 32;              // 32-column pieces per tile row
    constexpr int kPiecesPerWarp = kPieces / 2;
    constexpr int P = C::kRowBytes / 16;          // 16-byte chunks per staged row: 4 (bf16) or 8 (fp32)
    constexpr int kRowsPerInstr = 32 / P;         // rows covered by one warp-wide 16-byte store: 8 or 4
    uint8_t* my_stage = staging + ew * (32 * C::kRowBytes);
    int t = 0;
    for (int tile = w_first; tile < w_count; tile += w_step, ++t) {
      int m_blk, n_blk, split;
      tile_of(tile, m_blk, n_blk, split);
      OutT* const out_t = out ? out + (int64_t)split * M * N : nullptr;
      const int buf = t & 1;
      const uint32_t use_phase = (t >> 1) & 1;
      // stage the bias (and head weight) slice of this tile
      for (int j = et; j < BN; j += kEpiThreads) bias_s[j] = bias ? __ldg(bias + n_blk * BN + j) : 0.f;
      if (kHead && head_w)
        for (int j = et; j < 3 * BN; j += kEpiThreads)
          headw_s[j] = __ldg(head_w + (int64_t)(j / BN) * N + n_blk * BN + (j % BN));
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      ptx::mbar_wait(&tmem_full[buf], use_phase);
      ptx::tc_fence_after();
      // Each lane owns accumulator row (q*32 + lane).  A 32-column piece of the row (64 B bf16 / 128 B fp32) is
      // staged through shared memory with an XOR swizzle (conflict-free both ways) so that one warp-wide 16-byte
      // store covers 8 (4) rows x 64 (128) contiguous bytes instead of 32 rows x 16 bytes.
      const int64_t row_base = (int64_t)m_blk * BM + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
      const int wsw = (sizeof(OutT) == 2) ? ((lane >> 1) & 3) : (lane & 7);   // write-side swizzle of this lane's row
      float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;   // fused basis-status head: partial dot products of this row
#pragma unroll 1
      for (int pi = 0; pi < kPiecesPerWarp; ++pi) {
        const int pc = hsel * kPiecesPerWarp + pi;
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr + pc * 32, r);
        // keep-mask activations of the chunks this thread will store: fetched now, used after the staging round trip
        constexpr int kStoreIters = 32 / kRowsPerInstr;
        uint4 am[(kEpx && sizeof(OutT) == 2) ? kStoreIters : 1];
        if constexpr (kEpx && sizeof(OutT) == 2) {
          if (epx.mask_act && out) {
#pragma unroll
            for (int i = 0; i < kStoreIters; ++i) {
              const int64_t grow = row_base + kRowsPerInstr * i + lane / P;
              am[i] = make_uint4(0u, 0u, 0u, 0u);
              if (grow < M)
                am[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(epx.mask_act) +
                                                             ((grow * N + (int64_t)n_blk * BN + pc * 32) * 2 + (lane % P) * 16)));
            }
          }
        }
        ptx::tmem_ld_wait();
        uint4* srow = reinterpret_cast<uint4*>(my_stage + lane * C::kRowBytes);
        if constexpr (sizeof(OutT) == 2) {
          uint32_t packed[16];
          // dropout keep bits of this lane's 32 consecutive elements (8 hash quads), applied BEFORE the values are packed
          // or enter the fused head, so the head sees exactly the activations that are stored
          uint32_t keep = 0xffffffffu;
          if constexpr (kEpx) {
            if (epx.drop_threshold) {
              const uint64_t q0 = (uint64_t)((row_base + lane) * N + (int64_t)n_blk * BN + pc * 32) >> 2;
              keep = 0u;
#pragma unroll
              for (int qd = 0; qd < 8; ++qd) keep |= (dropout_keep4(epx.drop_seed, q0 + qd, epx.drop_threshold) & 15u) << (4 * qd);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v0 = __uint_as_float(r[2 * j]) + bias_s[pc * 32 + 2 * j];
            float v1 = __uint_as_float(r[2 * j + 1]) + bias_s[pc * 32 + 2 * j + 1];
            if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            if constexpr (kEpx) {
              v0 *= epx.out_scale; v1 *= epx.out_scale;
              if (!((keep >> (2 * j)) & 1u)) v0 = 0.f;
              if (!((keep >> (2 * j + 1)) & 1u)) v1 = 0.f;
            }
            packed[j] = Half16<OutT>::pack(v0, v1);
            if (kHead && head_w) {
              const int col = pc * 32 + 2 * j;
              hd0 = fmaf(v0, headw_s[col], hd0);          hd0 = fmaf(v1, headw_s[col + 1], hd0);
              hd1 = fmaf(v0, headw_s[BN + col], hd1);     hd1 = fmaf(v1, headw_s[BN + col + 1], hd1);
              hd2 = fmaf(v0, headw_s[2 * BN + col], hd2); hd2 = fmaf(v1, headw_s[2 * BN + col + 1], hd2);
            }
          }
          if (out) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              srow[j ^ wsw] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(r[j]) + bias_s[pc * 32 + j];
            if (relu) v = fmaxf(v, 0.f);
            r[j] = __float_as_uint(v);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) srow[j ^ wsw] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
        if (!out) continue;                       // head-only mode: the activation tile is never written
        __syncwarp();
        const int sub_row = lane / P, piece = lane % P;
#pragma unroll
        for (int i = 0; i < 32 / kRowsPerInstr; ++i) {
          const int rr = kRowsPerInstr * i + sub_row;
          const int rsw = (sizeof(OutT) == 2) ? ((rr >> 1) & 3) : (rr & 7);
          uint4 v = *reinterpret_cast<const uint4*>(my_stage + rr * C::kRowBytes + ((piece ^ rsw) * 16));
          const int64_t grow = row_base + rr;
          if (grow < M) {
            const int64_t e0 = grow * N + (int64_t)n_blk * BN + pc * 32;   // first element of this row piece
            if constexpr (sizeof(OutT) == 2) {
              // keep-masks act on whole bf16 lanes of the 16-byte chunk this thread stores (elements e0 + piece*8 ..)
              if (kEpx && epx.mask_act) {
                const uint4 a = am[i];
                v.x &= bf16x2_pos_mask(a.x); v.y &= bf16x2_pos_mask(a.y);
                v.z &= bf16x2_pos_mask(a.z); v.w &= bf16x2_pos_mask(a.w);
              }
            }
            uint8_t* dst = reinterpret_cast<uint8_t*>(out_t + e0);
            *reinterpret_cast<uint4*>(dst + piece * 16) = v;
          }
        }
        __syncwarp();
      }
      if (kHead && head_partial) {
        const int64_t grow = row_base + lane;
        if (grow < M) {
          float* hp = head_partial + ((int64_t)(n_blk * 2 + hsel) * M + grow) * 3;
          hp[0] = hd0; hp[1] = hd1; hp[2] = hd2;
        }
      }
      ptx::tc_fence_before();
      if constexpr (kCl > 1) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&tmem_empty[buf]), 0));
      else ptx::mbar_arrive(&tmem_empty[buf]);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // bias_s may be overwritten for the next tile
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCl > 1) ptx::cluster_sync();      // nothing of the peer is still in flight towards this CTA
  if (warp == 2) {
    if constexpr (kCl > 1) ptx::tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------- host side
// [rows, K] bf16 / half row-major, box = [box_rows, 64], 128-byte swizzle, OOB rows read as zero.
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int box_rows, bool f16 = false) {
  return make_map_16bit(map, base, rows, K, K, box_rows, f16, "node_transform(bf16)");
}

template <int BN, typename OutT, bool kMN, int kEpi>
int launch_epi(const Segs& segs, const float* bias, void* out, int M, int N, int relu, const float* head_w,
               float* head_partial, int ksplit, cudaStream_t st, const EpiX& epx) {
  using C = Cfg<BN, OutT>;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, OutT, kMN, 1, kEpi>;
  if (!attr_set) {
    LPGNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int tiles = ceil_div(M, BM) * (N / BN) * ksplit;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(segs, bias, reinterpret_cast<OutT*>(out), M, N, relu, head_w, head_partial,
                                              ksplit, epx);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

inline int epi_features(const float* head_w, const EpiX& epx) {
  return (head_w ? 1 : 0) | ((epx.mask_act || epx.drop_threshold || epx.out_scale != 1.f) ? 2 : 0);
}

template <int BN, typename OutT, bool kMN = false>
int launch(const Segs& segs, const float* bias, void* out, int M, int N, int relu, const float* head_w,
           float* head_partial, int ksplit, cudaStream_t st, const EpiX& epx = EpiX()) {
  // the wide bf16 tile gets one instantiation per feature set; fp32 outputs and MN-major operands never use the
  // extras, the narrow tiles (rare, small) share the full-featured one
  if constexpr (sizeof(OutT) == 4 || kMN) {
    return launch_epi<BN, OutT, kMN, 0>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
  } else if constexpr (std::is_same<OutT, __half>::value) {   // inference only: plain transform or fused head
    if (head_w) return launch_epi<BN, OutT, kMN, 1>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
    return launch_epi<BN, OutT, kMN, 0>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
  } else if constexpr (BN != 256) {
    return launch_epi<BN, OutT, kMN, 3>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
  } else {
    switch (epi_features(head_w, epx)) {
      case 0: return launch_epi<BN, OutT, kMN, 0>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
      case 1: return launch_epi<BN, OutT, kMN, 1>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
      case 2: return launch_epi<BN, OutT, kMN, 2>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
      default: return launch_epi<BN, OutT, kMN, 3>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx);
    }
  }
}

// CTA pairs (K-major, no split-K).  `segs.w` must have been encoded with BN / 2 box rows.
template <int BN, typename OutT, bool kMN, int kEpi>
int launch_cluster2_epi(const Segs& segs, const float* bias, void* out, int M, int N, int relu, const float* head_w,
                        float* head_partial, cudaStream_t st, const EpiX& epx, int ksplit) {
  using C = Cfg<BN, OutT, 2>;
  static int max_clusters = -1;
  auto kern = gemm_tc_kernel<BN, OutT, kMN, 2, kEpi>;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    LPGNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    cfg.gridDim = dim3(sm_count() / 2 * 2);
    int n = 0;
    LPGNN_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    max_clusters = n > 0 ? n : 1;
    if (getenv("LPGNN_DEBUG")) fprintf(stderr, "lpgnn: gemm pair kernel: max active clusters = %d (SMs %d)\n", n, sm_count());
  }
  const int pairs = ceil_div(ceil_div(M, BM), 2) * (N / BN) * ksplit;
  const int clusters = pairs < max_clusters ? pairs : max_clusters;
  cfg.gridDim = dim3(2 * clusters);
  OutT* out_t = reinterpret_cast<OutT*>(out);
  LPGNN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, segs, bias, out_t, M, N, relu, head_w, head_partial, ksplit, epx));
  count_launches(1);
  return LPGNN_OK;
}

template <int BN, typename OutT, bool kMN = false>
int launch_cluster2(const Segs& segs, const float* bias, void* out, int M, int N, int relu, const float* head_w,
                    float* head_partial, cudaStream_t st, const EpiX& epx, int ksplit = 1) {
  if constexpr (sizeof(OutT) == 4 || kMN) {
    return launch_cluster2_epi<BN, OutT, kMN, 0>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx, ksplit);
  } else if constexpr (std::is_same<OutT, __half>::value) {
    if (head_w) return launch_cluster2_epi<BN, OutT, kMN, 1>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx, ksplit);
    return launch_cluster2_epi<BN, OutT, kMN, 0>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx, ksplit);
  } else {
    switch (epi_features(head_w, epx)) {
      case 0: return launch_cluster2_epi<BN, OutT, kMN, 0>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx, ksplit);
      case 1: return launch_cluster2_epi<BN, OutT, kMN, 1>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx, ksplit);
      case 2: return launch_cluster2_epi<BN, OutT, kMN, 2>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx, ksplit);
      default: return launch_cluster2_epi<BN, OutT, kMN, 3>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx, ksplit);
    }
  }
}

int g_gemm_cluster = 1;   // 0 disables the 2-CTA cluster kernel (lpgnn_set_gemm_cluster, for A/B measurements)

}  // namespace

// out[M,N] = epi( sum_i A_i[M,K_i] * W_i[N,K_i]^T + bias ) over nseg <= 12 bf16 operand pairs.
// fmt: 0 = bf16 operands and output, 1 = bf16 operands, fp32 output, 2 = IEEE half operands and output (inference only).
int gemm_tc_run(const void* const* A, const void* const* W, const int* K, int nseg, const float* bias, int M, int N,
                void* out, int fmt, int relu, const float* head_w, float* head_partial, int ksplit, cudaStream_t st,
                const EpiX& epx) {
  const int out_f32 = fmt == 1;
  const bool f16 = fmt == 2;
  LPGNN_REQUIRE(!f16 || (!epx.mask_act && !epx.drop_threshold && epx.out_scale == 1.f && ksplit == 1),
                "node_transform(f16): the half format is inference-only (no mask / dropout epilogues, no split-K)");
  LPGNN_REQUIRE(!(out_f32 || !out) || (!epx.mask_act && !epx.drop_threshold && epx.out_scale == 1.f),
                "node_transform(bf16): mask / dropout epilogues need a bf16 output");
  LPGNN_REQUIRE(nseg >= 1 && nseg <= kMaxSegs, "gemm_tc: %d operand pairs (max %d)", nseg, kMaxSegs);
  LPGNN_REQUIRE(N % 64 == 0, "node_transform(bf16): N=%d must be a multiple of 64", N);
  LPGNN_REQUIRE((uintptr_t)out % 16 == 0, "node_transform(bf16): out must be 16-byte aligned");
  LPGNN_REQUIRE(out || (head_w && head_partial), "node_transform(bf16): no output requested");
  LPGNN_REQUIRE(!head_w || (head_partial && !out_f32), "node_transform(bf16): fused head needs head_partial and bf16 mode");
  const int BN = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  // wide layers with many row blocks: pairs of CTAs share the W tiles (see the kernel); bf16 output or fused head only
  // (long reductions only: with one or two K blocks per tile the kernel is epilogue-bound and the pair's extra
  // synchronisation costs more than it saves -- measured 103 vs 88 us on the C2 input layer)
  int kb_total = 0;
  for (int i = 0; i < nseg; ++i) kb_total += K[i] / BK;
  const bool use_cluster = g_gemm_cluster && BN == 256 && ksplit == 1 && !out_f32 && M >= 16 * BM && kb_total >= 8;
  Segs segs;
  int kb = 0;
  for (int i = 0; i < nseg; ++i) {
    LPGNN_REQUIRE(A[i] && W[i] && K[i] > 0 && K[i] % BK == 0, "node_transform(bf16): K=%d must be a positive multiple of 64",
                  K[i]);
    LPGNN_REQUIRE((uintptr_t)A[i] % 16 == 0 && (uintptr_t)W[i] % 16 == 0, "node_transform(bf16): operands must be 16-byte aligned");
    if (int rc = make_map(&segs.a[i], A[i], M, K[i], BM, f16)) return rc;
    if (int rc = make_map(&segs.w[i], W[i], N, K[i], use_cluster ? BN / 2 : BN, f16)) return rc;
    kb += K[i] / BK;
    segs.kb_end[i] = kb;
  }
  for (int i = nseg; i < kMaxSegs; ++i) { segs.a[i] = segs.a[0]; segs.w[i] = segs.w[0]; segs.kb_end[i] = kb; }
  segs.count = nseg;
  LPGNN_REQUIRE(ksplit >= 1 && (ksplit == 1 || (out_f32 && !bias && !relu && !head_w && ksplit <= kb)),
                "node_transform(bf16): split-K needs fp32 partial outputs, no epilogue, and ksplit <= K/64");
#define LPGNN_GO(BNV, T) return launch<BNV, T>(segs, bias, out, M, N, relu, head_w, head_partial, ksplit, st, epx)
  if (out_f32) {
    if (BN == 256) LPGNN_GO(256, float);
    if (BN == 128) LPGNN_GO(128, float);
    LPGNN_GO(64, float);
  }
  if (f16) {
    if (use_cluster) return launch_cluster2<256, __half>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx);
    if (BN == 256) LPGNN_GO(256, __half);
    if (BN == 128) LPGNN_GO(128, __half);
    LPGNN_GO(64, __half);
  }
  if (use_cluster) return launch_cluster2<256, __nv_bfloat16>(segs, bias, out, M, N, relu, head_w, head_partial, st, epx);
  if (BN == 256) LPGNN_GO(256, __nv_bfloat16);
  if (BN == 128) LPGNN_GO(128, __nv_bfloat16);
  LPGNN_GO(64, __nv_bfloat16);
#undef LPGNN_GO
}

// out[M,N] (fp32) = A^T B for A [Kred, M], B [Kred, N] row-major bf16 (MN-major operands), split-K over Kred.
// `out` is the [ksplit][M][N] partial buffer when ksplit > 1.
int gemm_tc_mn(const void* A, const void* B, int64_t Kred, int M, int N, void* out, int ksplit, cudaStream_t st) {
  LPGNN_REQUIRE(A && B && out && Kred > 0 && M > 0 && N % 64 == 0 && M % 8 == 0 && N % 8 == 0,
                "gemm_mn: bad arguments (M=%d, N=%d must be multiples of 8; N of 64)", M, N);
  LPGNN_REQUIRE((uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0 && (uintptr_t)out % 16 == 0, "gemm_mn: misaligned operand");
  const int BN = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("gemm_mn: cuTensorMapEncodeTiled unavailable"); return LPGNN_ECUDA; }
  Segs segs;
  auto mk = [&](CUtensorMap* map, const void* base, int cols) -> int {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)Kred};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gemm_mn: cuTensorMapEncodeTiled failed (%d)", (int)r); return LPGNN_ECUDA; }
    return LPGNN_OK;
  };
  if (int rc = mk(&segs.a[0], A, M)) return rc;
  if (int rc = mk(&segs.w[0], B, N)) return rc;
  const int kb = (int)((Kred + BK - 1) / BK);
  for (int i = 0; i < kMaxSegs; ++i) { if (i) { segs.a[i] = segs.a[0]; segs.w[i] = segs.w[0]; } segs.kb_end[i] = kb; }
  segs.count = 1;
  LPGNN_REQUIRE(ksplit >= 1 && ksplit <= kb, "gemm_mn: bad ksplit %d for %d K blocks", ksplit, kb);
  // wide weight gradients: CTA pairs (M = output rows = dY columns must give whole pairs of 128-row blocks)
  if (g_gemm_cluster && BN == 256 && M % (2 * BM) == 0 && kb / ksplit >= 8)
    return launch_cluster2<256, float, true>(segs, nullptr, out, M, N, 0, nullptr, nullptr, st, EpiX(), ksplit);
  if (BN == 256) return launch<256, float, true>(segs, nullptr, out, M, N, 0, nullptr, nullptr, ksplit, st);
  if (BN == 128) return launch<128, float, true>(segs, nullptr, out, M, N, 0, nullptr, nullptr, ksplit, st);
  return launch<64, float, true>(segs, nullptr, out, M, N, 0, nullptr, nullptr, ksplit, st);
}

int node_transform_bf16(const void* A1, int K1, const void* W1, const void* A2, int K2, const void* W2,
                        const float* bias, int M, int N, void* out, int fmt, int relu, const float* head_w,
                        float* head_partial, int ksplit, cudaStream_t st, const EpiX& epx) {
  const void* A[2] = {A1, A2};
  const void* W[2] = {W1, W2};
  const int K[2] = {K1, K2};
  const int nseg = (A2 != nullptr && K2 > 0) ? 2 : 1;
  return gemm_tc_run(A, W, K, nseg, bias, M, N, out, fmt, relu, head_w, head_partial, ksplit, st, epx);
}

}  // namespace lpgnn

// Tuning knob: enable (default) / disable the 2-CTA cluster form of the wide bf16 transform; returns the previous
// setting.  Results are identical either way (same MMAs in the same order per output tile).
extern "C" int lpgnn_set_gemm_cluster(int enable) {
  const int prev = lpgnn::g_gemm_cluster;
  lpgnn::g_gemm_cluster = enable ? 1 : 0;
  return prev;
}
