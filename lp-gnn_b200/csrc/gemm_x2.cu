// (a3, fp32 contract on the tensor cores) Node transform of a hidden GraphConv layer with fp32-level accuracy:
//   out[M,N] = epi( A1[M,K1]*W1[N,K1]^T + A2[M,K2]*W2[N,K2]^T + bias[N] )        fp32 in / out
//
// Replaces lin_rel(agg) + lin_root(x_dst) + relu_ of PyG GraphConv at the reference's DEFAULT precision
// (reference arch.py:75-80, 185-188 with `--fp16 0`, utils.py:770): two cuBLAS SGEMMs, a bias add, an add and an
// elementwise kernel there; here ONE tcgen05 kernel over "x2" operands.
//
// x2 format (lpgnn_split_x2): every fp32 row x[i,:] is stored as two IEEE-half rows and one power-of-two row scale,
//     x[i,k] = s_i * ( hi[i,k] + 2^-11 * lo[i,k] ),   hi = half(x / s_i),  lo = half( (x / s_i - hi) * 2^11 )
// with s_i chosen so that max_k |x[i,k]| / s_i lies in [2^12, 2^13): 22 significant bits per element (relative to the
// element, not to the row maximum, until the half subnormal range 2^-24 / 2^12 = 2^-36 of the row maximum), no
// overflow for any fp32 row.  A product a*w then is
//     a*w = s_a s_w ( a_hi w_hi  +  2^-11 (a_hi w_lo + a_lo w_hi)  +  2^-22 a_lo w_lo )
// and the kernel runs the first three terms as half x half -> fp32 MMAs (products of two 11-bit mantissas are exact
// in fp32); the dropped term is 2^-22 relative.  Three tensor-core passes instead of the six of a 3-way bf16 split.
//
// Accumulation.  The tensor core adds into its fp32 accumulator with truncation, so a long reduction in TMEM drifts
// (measured round 1: 1.5e-5 relative at K = 2048, which the row normalisation of add_knowledge amplifies past the
// 1e-4 logit bar).  Here a TMEM accumulator only ever holds a CHUNK of the reduction (chunk_kb K-blocks of 64,
// default 4 = 16 MMAs); the epilogue warps drain every chunk with tcgen05.ld and add it to per-thread fp32 REGISTER
// accumulators with one round-to-nearest FFMA per element (the 2^-11 of the correction passes rides in that FFMA,
// so the correction terms never meet the main term inside the truncating adder).  Two TMEM buffers of BN columns
// alternate, so the drain of chunk c overlaps the MMAs of chunk c+1.
//
//   warp 0      TMA producer (A tile 128 x 64 halves, W tile BN x 64, 128-byte swizzle, mbarrier ring)
//   warp 1      MMA issuer: tcgen05.mma kind::f16 (half operands), one commit per K-block (ring slot) and per chunk
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: thread = one accumulator row x BN/2 columns held in registers (128 for BN = 256; the
//               producer warpgroup hands its registers over with setmaxnreg), final scale / bias / ReLU, then
//               fp32 store and / or the fused basis-status head (partial dot products with head_w [3,N])
// CTA pairs (cta_group::2) as in gemm_tc.cu for wide layers.  Bound: tensor pipe, 3 * 2*M*N*(K1+K2) half flops.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace lpgnn {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UK = 16;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiWarp0 * 32 + kEpiThreads;   // 384
constexpr int kMaxSegs = 6;
constexpr int kRegsLight = 56, kRegsHeavy = 224;          // 128*56 + 256*224 = 64512 = 384*168

template <int BN, int kCl> struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / kCl) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kStageBytes == 49152) ? 4 : (kStageBytes == 32768 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;                // two chunk buffers
  static constexpr int kBarBytes = 256 + 5 * BN * 4;      // mbarriers + tmem ptr | bias, column scale, 3 x head weights
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024 /*align slack*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static_assert(2 * kStages * 8 + 48 <= 256, "barrier block too small");
};

struct X2Segs {
  CUtensorMap a[kMaxSegs];
  CUtensorMap w[kMaxSegs];
  int kb_end[kMaxSegs];     // cumulative K-block count at the end of each segment
  int chunk_kb[kMaxSegs];   // K-blocks per TMEM chunk inside the segment
  float scale[kMaxSegs];    // weight of the segment's partial sums in the register accumulator (1 or 2^-11)
  int rs_sel[kMaxSegs];     // which row-scale array the segment's A operand belongs to (0: rowscale, 1: rowscale2)
  int count;
};

__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLight)); }
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsHeavy)); }

template <int BN, int kCl, bool kHead>
__global__ void __launch_bounds__(kThreads, 1)
gemm_x2_kernel(const __grid_constant__ X2Segs segs, const float* __restrict__ bias, const float* __restrict__ rowscale,
               const float* __restrict__ rowscale2, const float* __restrict__ colscale, float* __restrict__ out, int M, int N, int relu,
               const float* __restrict__ head_w /*[3,N]*/, float* __restrict__ head_partial /*[2N/BN][M][3]*/) {
  using C = Cfg<BN, kCl>;
  constexpr int NACC = BN / 2;                     // accumulator columns per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = bars + 2 * C::kStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);
  float* bias_s = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + 256);   // [BN]
  float* cs_s = bias_s + BN;                                                             // [BN]
  float* headw_s = cs_s + BN;                                                            // [3][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = N / BN;
  const int num_m = (M + BM - 1) / BM;
  const uint32_t crank = (kCl > 1) ? ptx::cluster_ctarank() : 0u;
  const int w_first = (kCl > 1) ? (int)(blockIdx.x / kCl) : (int)blockIdx.x;
  const int w_step = (kCl > 1) ? (int)(gridDim.x / kCl) : (int)gridDim.x;
  const int w_count = ((num_m + kCl - 1) / kCl) * num_n;       // tiles (kCl = 1) or vertically adjacent tile pairs
  auto tile_of = [&](int wi, int& m_blk, int& n_blk) { m_blk = kCl * (wi / num_n) + (int)crank; n_blk = wi % num_n; };
  const int kblocks_all = segs.kb_end[segs.count - 1];

  if (warp == 0 && lane == 0)
    for (int i = 0; i < segs.count; ++i) { ptx::prefetch_tensormap(&segs.a[i]); ptx::prefetch_tensormap(&segs.w[i]); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tmem_full[b], 1); ptx::mbar_init(&tmem_empty[b], kCl * kEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kCl > 1) { ptx::tmem_alloc_pair(tmem_ptr, C::kTmemCols); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(tmem_ptr, C::kTmemCols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCl > 1) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kEpiWarp0) {
    setmaxnreg_dec();
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int stage = 0; uint32_t phase = 0;
      for (int tile = w_first; tile < w_count; tile += w_step) {
        int m_blk, n_blk;
        tile_of(tile, m_blk, n_blk);
        int sg = 0, seg_begin = 0;
        for (int kb = 0; kb < kblocks_all; ++kb) {
          while (kb >= segs.kb_end[sg]) { seg_begin = segs.kb_end[sg]; ++sg; }
          const int kc = (kb - seg_begin) * BK;
          ptx::mbar_wait_quiet(&empty_bar[stage], phase ^ 1);
          if constexpr (kCl > 1) {
            // both CTAs' tiles complete on the leader's barrier; the leader alone posts the expected byte count
            const uint32_t lead_bar = ptx::mapa_shared(ptx::smem_u32(&full_bar[stage]), 0);
            if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], kCl * C::kStageBytes);
            ptx::tma_load_2d_pair(smem_a + stage * C::kABytes, &segs.a[sg], lead_bar, kc, m_blk * BM);
            ptx::tma_load_2d_pair(smem_b + stage * C::kBBytes, &segs.w[sg], lead_bar, kc, n_blk * BN + (int)crank * (BN / kCl));
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
            ptx::tma_load_2d(smem_a + stage * C::kABytes, &segs.a[sg], &full_bar[stage], kc, m_blk * BM);
            ptx::tma_load_2d(smem_b + stage * C::kBBytes, &segs.w[sg], &full_bar[stage], kc, n_blk * BN);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0 && (kCl == 1 || crank == 0)) {
      // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM * kCl, BN, false, /*f16=*/true);
      int stage = 0; uint32_t phase = 0;
      uint32_t c = 0;                               // chunks issued so far (all tiles): buffer c & 1, use (c >> 1)
      for (int tile = w_first; tile < w_count; tile += w_step) {
        int seg_begin = 0;
        for (int sg = 0; sg < segs.count; ++sg) {
          const int seg_end = segs.kb_end[sg], ck = segs.chunk_kb[sg];
          for (int cb = seg_begin; cb < seg_end; cb += ck, ++c) {
            const uint32_t buf = c & 1u;
            ptx::mbar_wait_quiet(&tmem_empty[buf], ((c >> 1) & 1u) ^ 1u);   // the epilogue has drained this buffer
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * BN;
            const int ce = min(cb + ck, seg_end);
            for (int kb = cb; kb < ce; ++kb) {
              ptx::mbar_wait_quiet(&full_bar[stage], phase);
              ptx::tc_fence_after();
              const uint64_t adesc = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_a + stage * C::kABytes));
              const uint64_t bdesc = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_b + stage * C::kBBytes));
#pragma unroll
              for (int k = 0; k < BK / UK; ++k) {
                if constexpr (kCl > 1) ptx::umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > cb || k > 0) ? 1u : 0u);
                else ptx::umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > cb || k > 0) ? 1u : 0u);
              }
              if constexpr (kCl > 1) ptx::umma_commit_pair(&empty_bar[stage]);
              else ptx::umma_commit(&empty_bar[stage]);
              if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
            if constexpr (kCl > 1) ptx::umma_commit_pair(&tmem_full[buf]);
            else ptx::umma_commit(&tmem_full[buf]);
          }
          seg_begin = seg_end;
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue
    setmaxnreg_inc();
    const int ew = warp - kEpiWarp0;              // 0..7
    const int q = ew & 3;                         // TMEM lane quarter (== warp % 4)
    const int hsel = ew >> 2;                     // column half of the tile
    const int et = threadIdx.x - kEpiWarp0 * 32;
    uint32_t c = 0;
    for (int tile = w_first; tile < w_count; tile += w_step) {
      int m_blk, n_blk;
      tile_of(tile, m_blk, n_blk);
      for (int j = et; j < BN; j += kEpiThreads) {
        bias_s[j] = bias ? __ldg(bias + n_blk * BN + j) : 0.f;
        cs_s[j] = colscale ? __ldg(colscale + n_blk * BN + j) : 1.f;
      }
      if (kHead)
        for (int j = et; j < 3 * BN; j += kEpiThreads)
          headw_s[j] = __ldg(head_w + (int64_t)(j / BN) * N + n_blk * BN + (j % BN));
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");

      float acc[NACC];
#pragma unroll
      for (int j = 0; j < NACC; ++j) acc[j] = 0.f;
      // this thread's row and the power-of-two scales of its two A operands (exact factors; rows past M contribute 0)
      const int64_t row = (int64_t)m_blk * BM + q * 32 + lane;
      const bool row_ok = row < M;
      const float rs1 = row_ok ? (rowscale ? __ldg(rowscale + row) : 1.f) : 0.f;
      const float rs2 = row_ok ? (rowscale2 ? __ldg(rowscale2 + row) : rs1) : 0.f;
      int seg_begin = 0;
      for (int sg = 0; sg < segs.count; ++sg) {
        const int seg_end = segs.kb_end[sg], ck = segs.chunk_kb[sg];
        const float sc = segs.scale[sg] * (segs.rs_sel[sg] ? rs2 : rs1);
        for (int cb = seg_begin; cb < seg_end; cb += ck, ++c) {
          const uint32_t buf = c & 1u;
          ptx::mbar_wait_quiet(&tmem_full[buf], (c >> 1) & 1u);
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + hsel * NACC;
#pragma unroll
          for (int p = 0; p < NACC / 32; p += 2) {
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32(taddr + p * 32, r0);
            if (p + 1 < NACC / 32) ptx::tmem_ld_32x32(taddr + (p + 1) * 32, r1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[p * 32 + j] = fmaf(sc, __uint_as_float(r0[j]), acc[p * 32 + j]);
            if (p + 1 < NACC / 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[(p + 1) * 32 + j] = fmaf(sc, __uint_as_float(r1[j]), acc[(p + 1) * 32 + j]);
            }
          }
          // buffer drained: one arrival per warp (pair: on the leader's barrier, which hears both CTAs)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCl > 1) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&tmem_empty[buf]), 0));
            else ptx::mbar_arrive(&tmem_empty[buf]);
          }
        }
        seg_begin = seg_end;
      }

      // out = epi( s_col * acc + bias ): the row scales rode in with the chunks, the column scale is a power of two (exact)
      const int col0 = hsel * NACC;
      float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;
      float* orow = out ? out + row * N + (int64_t)n_blk * BN + col0 : nullptr;
#pragma unroll
      for (int j4 = 0; j4 < NACC; j4 += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int col = col0 + j4 + u;
          float t = fmaf(acc[j4 + u], cs_s[col], bias_s[col]);
          if (relu) t = fmaxf(t, 0.f);
          v[u] = t;
          if (kHead) {
            hd0 = fmaf(t, headw_s[col], hd0);
            hd1 = fmaf(t, headw_s[BN + col], hd1);
            hd2 = fmaf(t, headw_s[2 * BN + col], hd2);
          }
        }
        if (orow && row_ok) *reinterpret_cast<float4*>(orow + j4) = make_float4(v[0], v[1], v[2], v[3]);
      }
      if (kHead && row_ok) {
        float* hp = head_partial + ((int64_t)(n_blk * 2 + hsel) * M + row) * 3;
        hp[0] = hd0; hp[1] = hd1; hp[2] = hd2;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");   // bias_s is rewritten for the next tile
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCl > 1) ptx::cluster_sync();
  if (warp == 2) {
    if constexpr (kCl > 1) ptx::tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int BN, int kCl, bool kHead>
int launch_x2(const X2Segs& segs, const float* bias, const float* rowscale, const float* rowscale2, const float* colscale, float* out,
              int M, int N, int relu, const float* head_w, float* head_partial, cudaStream_t st) {
  using C = Cfg<BN, kCl>;
  auto kern = gemm_x2_kernel<BN, kCl, kHead>;
  static int max_clusters = -1;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    LPGNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    if (kCl > 1) {
      cfg.gridDim = dim3(sm_count() / kCl * kCl);
      int n = 0;
      LPGNN_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      max_clusters = n > 0 ? n : 1;
    } else {
      max_clusters = sm_count();
    }
  }
  const int items = ceil_div(ceil_div(M, BM), kCl) * (N / BN);
  const int clusters = items < max_clusters ? items : max_clusters;
  cfg.gridDim = dim3(kCl * clusters);
  LPGNN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, segs, bias, rowscale, rowscale2, colscale, out, M, N, relu, head_w, head_partial));
  count_launches(1);
  return LPGNN_OK;
}

int g_x2_chunk_kb = 4;        // K-blocks of 64 per TMEM chunk of the main (hi x hi) passes
int g_x2_corr_chunk_kb = 16;  // ... of the correction passes (their drift is scaled by 2^-11)

// One warp per row: the row maximum over both tensors fixes the power-of-two scale, then every element is split.
__global__ void __launch_bounds__(256)
split_x2_kernel(const float* __restrict__ x1, int K1, const float* __restrict__ x2, int K2, int64_t rows,
                __half* __restrict__ hi1, __half* __restrict__ lo1, __half* __restrict__ hi2, __half* __restrict__ lo2,
                float* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float4* p1 = reinterpret_cast<const float4*>(x1 + r * K1);
    const float4* p2 = x2 ? reinterpret_cast<const float4*>(x2 + r * K2) : nullptr;
    const int q1 = K1 >> 2, q2 = x2 ? (K2 >> 2) : 0;
    float mx = 0.f;
    for (int i = lane; i < q1; i += 32) { const float4 v = __ldg(p1 + i); mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)))); }
    for (int i = lane; i < q2; i += 32) { const float4 v = __ldg(p2 + i); mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)))); }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // mx = f * 2^e, f in [1,2): scaled row maximum in [2^12, 2^13).  Zero / non-finite rows keep scale 1; the exponent
    // is clamped so that both 2^(12-e) and 2^(e-12) are normal floats.
    int e = 12;
    if (mx > 0.f && mx < __int_as_float(0x7f800000)) e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;
    e = max(-100, min(e, 112));
    const float down = __int_as_float((uint32_t)(127 + 12 - e) << 23);   // 2^(12-e)
    if (lane == 0) scale[r] = __int_as_float((uint32_t)(127 + e - 12) << 23);
    auto emit = [&](const float4* src, int quads, __half* hi, __half* lo, int K) {
      uint2* ho = reinterpret_cast<uint2*>(hi + r * K);
      uint2* lw = reinterpret_cast<uint2*>(lo + r * K);
      for (int i = lane; i < quads; i += 32) {
        const float4 v = __ldg(src + i);
        const float s[4] = {v.x * down, v.y * down, v.z * down, v.w * down};
        __half h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          h[k] = __float2half_rn(s[k]);
          l[k] = __float2half_rn((s[k] - __half2float(h[k])) * 2048.f);
        }
        const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
        const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
        ho[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        lw[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
    };
    emit(p1, q1, hi1, lo1, K1);
    if (x2) emit(p2, q2, hi2, lo2, K2);
  }
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" int lpgnn_split_x2(const float* x1, int32_t K1, const float* x2, int32_t K2, int64_t rows, void* hi1, void* lo1,
                              void* hi2, void* lo2, float* scale, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && K1 > 0 && K1 % 4 == 0 && K2 >= 0 && K2 % 4 == 0, "split_x2: K1=%d, K2=%d must be multiples of 4", K1, K2);
  if (rows == 0) return LPGNN_OK;
  if (K2 == 0) x2 = nullptr;
  LPGNN_REQUIRE(x1 && hi1 && lo1 && scale && (!x2 || (hi2 && lo2)), "split_x2: null pointer");
  LPGNN_REQUIRE((uintptr_t)x1 % 16 == 0 && (uintptr_t)x2 % 16 == 0 && (uintptr_t)hi1 % 8 == 0 && (uintptr_t)lo1 % 8 == 0 &&
                (uintptr_t)hi2 % 8 == 0 && (uintptr_t)lo2 % 8 == 0, "split_x2: misaligned pointer");
  const int64_t want = (rows + 7) / 8, cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want < cap ? want : cap);
  split_x2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x1, K1, x2, K2, rows, (__half*)hi1, (__half*)lo1, (__half*)hi2,
                                                        (__half*)lo2, scale);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_node_transform_x2(const void* A1_hi, const void* A1_lo, int32_t K1, const void* W1_hi, const void* W1_lo,
                                       const void* A2_hi, const void* A2_lo, int32_t K2, const void* W2_hi, const void* W2_lo,
                                       const float* rowscale, const float* rowscale2, const float* colscale, const float* bias,
                                       int32_t M, int32_t N, float* out, int epilogue, const float* head_w, float* head_partial,
                                       lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(M >= 0 && N > 0 && N % 64 == 0 && K1 > 0 && K1 % BK == 0 && K2 >= 0 && K2 % BK == 0,
                "node_transform_x2: bad shape M=%d N=%d K1=%d K2=%d (N, K multiples of 64)", M, N, K1, K2);
  if (M == 0) return LPGNN_OK;
  if (K2 == 0) { A2_hi = A2_lo = W2_hi = W2_lo = nullptr; }
  LPGNN_REQUIRE(A1_hi && A1_lo && W1_hi && W1_lo && (K2 == 0 || (A2_hi && A2_lo && W2_hi && W2_lo)), "node_transform_x2: null operand");
  LPGNN_REQUIRE(out || (head_w && head_partial), "node_transform_x2: no output requested");
  LPGNN_REQUIRE(!head_w || head_partial, "node_transform_x2: fused head needs head_partial");
  LPGNN_REQUIRE((uintptr_t)out % 16 == 0, "node_transform_x2: out must be 16-byte aligned");
  const void* ops[8] = {A1_hi, A1_lo, W1_hi, W1_lo, A2_hi, A2_lo, W2_hi, W2_lo};
  for (int i = 0; i < 8; ++i) LPGNN_REQUIRE((uintptr_t)ops[i] % 16 == 0, "node_transform_x2: operands must be 16-byte aligned");
  const int BN = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  const int kb_total = 3 * (K1 + K2) / BK;
  const bool pair = BN == 256 && M >= 16 * BM && kb_total >= 8;
  // segment order: the correction passes (weight 2^-11) first, then the main passes
  X2Segs segs;
  int n = 0, kb = 0;
  auto add = [&](const void* a, const void* w, int K, float scale, int chunk, int rs_sel) -> int {
    if (int rc = make_map_16bit(&segs.a[n], a, M, K, K, BM, true, "node_transform_x2")) return rc;
    if (int rc = make_map_16bit(&segs.w[n], w, N, K, K, pair ? BN / 2 : BN, true, "node_transform_x2")) return rc;
    kb += K / BK;
    segs.kb_end[n] = kb; segs.chunk_kb[n] = chunk; segs.scale[n] = scale; segs.rs_sel[n] = rs_sel;
    ++n;
    return LPGNN_OK;
  };
  const float corr = 1.f / 2048.f;
  const int cm = g_x2_chunk_kb, cc = g_x2_corr_chunk_kb;
  if (int rc = add(A1_hi, W1_lo, K1, corr, cc, 0)) return rc;
  if (int rc = add(A1_lo, W1_hi, K1, corr, cc, 0)) return rc;
  if (K2) {
    if (int rc = add(A2_hi, W2_lo, K2, corr, cc, 1)) return rc;
    if (int rc = add(A2_lo, W2_hi, K2, corr, cc, 1)) return rc;
  }
  if (int rc = add(A1_hi, W1_hi, K1, 1.f, cm, 0)) return rc;
  if (K2) if (int rc = add(A2_hi, W2_hi, K2, 1.f, cm, 1)) return rc;
  for (int i = n; i < kMaxSegs; ++i) { segs.a[i] = segs.a[0]; segs.w[i] = segs.w[0]; segs.kb_end[i] = kb; segs.chunk_kb[i] = 1; segs.scale[i] = 0.f; segs.rs_sel[i] = 0; }
  segs.count = n;
  const int relu = (epilogue & LPGNN_EPI_RELU) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
#define LPGNN_X2_GO(BNV, CL)                                                                                          \
  return head_w ? launch_x2<BNV, CL, true>(segs, bias, rowscale, rowscale2, colscale, out, M, N, relu, head_w, head_partial, st) \
                : launch_x2<BNV, CL, false>(segs, bias, rowscale, rowscale2, colscale, out, M, N, relu, head_w, head_partial, st)
  if (pair) LPGNN_X2_GO(256, 2);
  if (BN == 256) LPGNN_X2_GO(256, 1);
  if (BN == 128) LPGNN_X2_GO(128, 1);
  LPGNN_X2_GO(64, 1);
#undef LPGNN_X2_GO
}

// Tuning knob (process-wide): K-blocks of 64 accumulated inside TMEM before the epilogue adds the chunk to its fp32
// registers (main passes, 1..64; the correction passes use 4x that, at least 16).  Returns the previous setting.
extern "C" int lpgnn_set_x2_chunk(int kblocks) {
  const int prev = g_x2_chunk_kb;
  if (kblocks >= 1 && kblocks <= 64) {
    g_x2_chunk_kb = kblocks;
    g_x2_corr_chunk_kb = kblocks * 4 > 16 ? kblocks * 4 : 16;
  }
  return prev;
}
