// (a3, fp32 parity mode) Node transform on CUDA cores:
//   out[M,N] = epi( A1[M,K1]*W1[N,K1]^T + A2[M,K2]*W2[N,K2]^T + bias[N] )       all float32
//
// Replaces lin_rel(agg) + lin_root(x_dst) (+ relu_) of PyG GraphConv reached from reference
// arch.py:75-80 / 188 in the mode that has to match the reference's fp32 arithmetic to 1e-4
// (bf16 tensor-core products cannot).  Classic register-tiled SGEMM: 128x128 block tile, 8x8 per
// thread, K step 16, operands staged through shared memory transposed to [k][m] so the inner
// product reads conflict-free float4 rows; global loads are 128-bit along K and prefetched into
// registers while the previous step is multiplied.  The two (A,W) pairs are walked as one
// concatenated reduction, so the layer is one launch and `out` is written once.
// Bound: fp32 FMA pipe (2*M*N*(K1+K2) flops); the bf16 tcgen05 kernel in gemm_tc.cu is the fast path.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
constexpr int kThreads = (BM / TM) * (BN / TN);  // 256

struct Seg {
  const float* A;
  const float* W;
  int K;
};

__global__ void __launch_bounds__(kThreads)
sgemm_cat_kernel(Seg s0, Seg s1, const float* __restrict__ bias, int M, int N, float* __restrict__ out, int relu) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  // loader mapping: 128 rows x 16 k = 512 float4; thread loads float4 #tid and #tid+256
  const int l_row = tid >> 2;         // 0..63 (+64 for the second)
  const int l_k4 = (tid & 3) * 4;     // 0,4,8,12

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int steps0 = (s0.K + BK - 1) / BK, steps1 = (s1.K + BK - 1) / BK;
  const int steps = steps0 + steps1;
  float4 ra[2], rw[2];

  auto fetch = [&](int step) {
    const Seg& s = (step < steps0) ? s0 : s1;
    const int k = ((step < steps0) ? step : step - steps0) * BK + l_k4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = m0 + l_row + h * 64;
      const int c = n0 + l_row + h * 64;
      ra[h] = (r < M && k < s.K) ? __ldg(reinterpret_cast<const float4*>(s.A + r * s.K + k)) : make_float4(0, 0, 0, 0);
      rw[h] = (c < N && k < s.K) ? __ldg(reinterpret_cast<const float4*>(s.W + (int64_t)c * s.K + k))
                                 : make_float4(0, 0, 0, 0);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = l_row + h * 64;
      As[buf][l_k4 + 0][r] = ra[h].x; As[buf][l_k4 + 1][r] = ra[h].y;
      As[buf][l_k4 + 2][r] = ra[h].z; As[buf][l_k4 + 3][r] = ra[h].w;
      Ws[buf][l_k4 + 0][r] = rw[h].x; Ws[buf][l_k4 + 1][r] = rw[h].y;
      Ws[buf][l_k4 + 2][r] = rw[h].z; Ws[buf][l_k4 + 3][r] = rw[h].w;
    }
  };

  if (steps > 0) { fetch(0); stash(0); }
  __syncthreads();
  for (int step = 0; step < steps; ++step) {
    const int buf = step & 1;
    if (step + 1 < steps) fetch(step + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * TN]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * TN + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (step + 1 < steps) stash(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t r = m0 + ty * TM + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = n0 + tx * TN + j;
      if (c >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + c) : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      out[r * N + c] = v;
    }
  }
}

}  // namespace

int node_transform_f32(const float* A1, int K1, const float* W1, const float* A2, int K2, const float* W2,
                       const float* bias, int M, int N, float* out, int relu, cudaStream_t st) {
  LPGNN_REQUIRE(K1 % 4 == 0 && K2 % 4 == 0, "node_transform(f32): K1=%d and K2=%d must be multiples of 4", K1, K2);
  LPGNN_REQUIRE((uintptr_t)A1 % 16 == 0 && (uintptr_t)W1 % 16 == 0 && (uintptr_t)A2 % 16 == 0 &&
                    (uintptr_t)W2 % 16 == 0,
                "node_transform(f32): operands must be 16-byte aligned");
  Seg s0{A1, W1, K1}, s1{A2, W2, A2 ? K2 : 0};
  LPGNN_REQUIRE(ceil_div(M, BM) <= 65535, "node_transform(f32): M=%d exceeds the %d rows one launch covers", M, 65535 * BM);
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  sgemm_cat_kernel<<<grid, kThreads, 0, st>>>(s0, s1, bias, M, N, out, relu);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

}  // namespace lpgnn
