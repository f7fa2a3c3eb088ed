// C-ABI plumbing: error state, device checks, dtype dispatch of the node transform.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace lpgnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

static int g_sm_count = 0;
static int g_cc_major = 0, g_cc_minor = 0;
static int g_dev_checked = -1;

int check_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) {
    cudaGetLastError();
    set_error("no CUDA device available: liblpgnn has no CPU fallback");
    return LPGNN_ENODEVICE;
  }
  if (dev != g_dev_checked) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaGetDeviceProperties failed");
      return LPGNN_ENODEVICE;
    }
    g_sm_count = p.multiProcessorCount;
    g_cc_major = p.major;
    g_cc_minor = p.minor;
    g_dev_checked = dev;
  }
  if (g_cc_major != 10) {
    set_error("device compute capability %d.%d is not sm_100 (Blackwell B200); liblpgnn is sm_100a-only", g_cc_major,
              g_cc_minor);
    return LPGNN_ENODEVICE;
  }
  return LPGNN_OK;
}

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

int node_transform_f32(const float* A1, int K1, const float* W1, const float* A2, int K2, const float* W2,
                       const float* bias, int M, int N, float* out, int relu, cudaStream_t st);
int node_transform_bf16(const void* A1, int K1, const void* W1, const void* A2, int K2, const void* W2,
                        const float* bias, int M, int N, void* out, int out_f32, int relu, const float* head_w,
                        float* head_partial, int ksplit, cudaStream_t st, const EpiX& epx = EpiX());

int gemm_tc_run(const void* const* A, const void* const* W, const int* K, int nseg, const float* bias, int M, int N,
                void* out, int out_f32, int relu, const float* head_w, float* head_partial, int ksplit, cudaStream_t st,
                const EpiX& epx = EpiX());

// x = p0 + p1 (+ p2) with p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1): the operand form of the
// fp32-accurate tensor-core transform (2 parts: ~2^-17 relative, 3 parts: ~2^-25)
template <int PARTS>
__global__ void split_bf16_kernel(const float4* __restrict__ x, int64_t quads, uint2* __restrict__ p0,
                                  uint2* __restrict__ p1, uint2* __restrict__ p2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < quads; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    float r[4] = {v.x, v.y, v.z, v.w};
    uint2* outs[3] = {p0, p1, p2};
#pragma unroll
    for (int part = 0; part < PARTS; ++part) {
      __nv_bfloat16 h[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { h[k] = __float2bfloat16_rn(r[k]); r[k] -= __bfloat162float(h[k]); }
      __nv_bfloat162 a = __halves2bfloat162(h[0], h[1]), b = __halves2bfloat162(h[2], h[3]);
      outs[part][i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
  }
}

}  // namespace lpgnn

using namespace lpgnn;

extern "C" int lpgnn_split_bf16(const float* x, int64_t count, int parts, void* const* out_parts, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(count >= 0 && count % 4 == 0 && (parts == 2 || parts == 3), "split_bf16: count %% 4 != 0 or parts not in {2,3}");
  if (count == 0) return LPGNN_OK;
  LPGNN_REQUIRE(x && out_parts && (uintptr_t)x % 16 == 0, "split_bf16: null or misaligned pointer");
  for (int i = 0; i < parts; ++i)
    LPGNN_REQUIRE(out_parts[i] && (uintptr_t)out_parts[i] % 8 == 0, "split_bf16: null or misaligned output %d", i);
  const int64_t quads = count / 4;
  const int64_t want = (quads + 255) / 256, cap = (int64_t)sm_count() * 16;
  const int grid = (int)(want < cap ? want : cap);
  cudaStream_t st = (cudaStream_t)stream;
  uint2 *p0 = reinterpret_cast<uint2*>(out_parts[0]), *p1 = reinterpret_cast<uint2*>(out_parts[1]);
  if (parts == 2) split_bf16_kernel<2><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(x), quads, p0, p1, nullptr);
  else split_bf16_kernel<3><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(x), quads, p0, p1,
                                                  reinterpret_cast<uint2*>(out_parts[2]));
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

// fp32-accurate node transform on the tensor cores: every fp32 operand arrives split into `parts` bf16 tensors.
//   parts = 2: a*w ~ a0*w0 + a0*w1 + a1*w0                                   (3 passes, ~2^-17 relative)
//   parts = 3: a*w ~ a0*w0 + a0*w1 + a1*w0 + a0*w2 + a2*w0 + a1*w1           (6 passes, ~2^-24 relative)
extern "C" int lpgnn_node_transform_split(int parts, const void* const* A1, int32_t K1, const void* const* W1,
                                          const void* const* A2, int32_t K2, const void* const* W2,
                                          const float* bias, int32_t M, int32_t N, float* out, int epilogue,
                                          lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE((parts == 2 || parts == 3) && M >= 0 && N > 0 && K1 > 0 && K2 >= 0, "node_transform_split: bad arguments");
  if (M == 0) return LPGNN_OK;
  LPGNN_REQUIRE(A1 && W1 && out && (K2 == 0 || (A2 && W2)), "node_transform_split: null pointer");
  static const int pa[6] = {0, 0, 1, 0, 2, 1}, pw[6] = {0, 1, 0, 2, 0, 1};
  const int nprod = parts == 2 ? 3 : 6;
  const void* A[12]; const void* W[12]; int K[12];
  int n = 0;
  for (int i = 0; i < nprod; ++i, ++n) { A[n] = A1[pa[i]]; W[n] = W1[pw[i]]; K[n] = K1; }
  if (K2 > 0) for (int i = 0; i < nprod; ++i, ++n) { A[n] = A2[pa[i]]; W[n] = W2[pw[i]]; K[n] = K2; }
  return gemm_tc_run(A, W, K, n, bias, M, N, out, 1, (epilogue & LPGNN_EPI_RELU) ? 1 : 0, nullptr, nullptr, 1,
                     (cudaStream_t)stream);
}

extern "C" int lpgnn_version(void) { return LPGNN_VERSION; }
extern "C" const char* lpgnn_last_error(void) { return g_err; }
extern "C" uint64_t lpgnn_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int lpgnn_device_info(int* sm, int* major, int* minor) {
  if (int rc = check_device()) return rc;
  if (sm) *sm = g_sm_count;
  if (major) *major = g_cc_major;
  if (minor) *minor = g_cc_minor;
  return LPGNN_OK;
}

extern "C" int lpgnn_node_transform(const void* A1, int32_t K1, const void* W1, const void* A2, int32_t K2,
                                    const void* W2, const float* bias, int32_t M, int32_t N, void* out, int dtype,
                                    int out_dtype, int epilogue, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(M >= 0 && N > 0 && K1 > 0 && K2 >= 0, "node_transform: bad shape M=%d N=%d K1=%d K2=%d", M, N, K1, K2);
  LPGNN_REQUIRE(dtype_ok(dtype), "node_transform: bad dtype %d", dtype);
  LPGNN_REQUIRE((out_dtype == LPGNN_F32 && dtype != LPGNN_F16) || (is_16bit(out_dtype) && out_dtype == dtype),
                "node_transform: out_dtype %d not available for operand dtype %d", out_dtype, dtype);
  if (M == 0) return LPGNN_OK;
  LPGNN_REQUIRE(A1 && W1 && out, "node_transform: null pointer");
  LPGNN_REQUIRE(K2 == 0 || (A2 && W2), "node_transform: K2=%d but A2/W2 is null", K2);
  const int relu = (epilogue & LPGNN_EPI_RELU) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (K2 == 0) { A2 = nullptr; W2 = nullptr; }
  if (dtype == LPGNN_F32)
    return node_transform_f32((const float*)A1, K1, (const float*)W1, (const float*)A2, K2, (const float*)W2, bias, M,
                              N, (float*)out, relu, st);
  return node_transform_bf16(A1, K1, W1, A2, K2, W2, bias, M, N, out, dtype == LPGNN_F16 ? 2 : (out_dtype == LPGNN_F32 ? 1 : 0),
                             relu, nullptr, nullptr, 1, st);
}

// node_transform + fused keep-masks (training): see include/lpgnn.h
extern "C" int lpgnn_node_transform_ex(const void* A1, int32_t K1, const void* W1, const void* A2, int32_t K2,
                                       const void* W2, const float* bias, int32_t M, int32_t N, void* out, int dtype,
                                       const lpgnn_epilogue_args* epi, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(epi, "node_transform_ex: null epilogue arguments");
  LPGNN_REQUIRE(dtype == LPGNN_F32 || dtype == LPGNN_BF16, "node_transform_ex: bad dtype %d (training epilogues: f32 / bf16)", dtype);
  LPGNN_REQUIRE(epi->dropout_p >= 0.f && epi->dropout_p < 1.f, "node_transform_ex: dropout_p=%f outside [0,1)", epi->dropout_p);
  const float mscale = epi->mask_act ? epi->mask_scale : 1.f;
  if (dtype == LPGNN_F32) {   // CUDA-core transform, then the stand-alone mask / dropout kernels (same semantics)
    if (int rc = lpgnn_node_transform(A1, K1, W1, A2, K2, W2, bias, M, N, out, dtype, dtype, epi->epilogue, stream)) return rc;
    if (epi->mask_act)
      if (int rc = lpgnn_relu_bwd(out, nullptr, epi->mask_act, (int64_t)M * N, dtype, mscale, out, stream)) return rc;
    if (epi->dropout_p > 0.f) return lpgnn_dropout(out, (int64_t)M * N, dtype, epi->dropout_p, epi->dropout_seed, stream);
    return LPGNN_OK;
  }
  LPGNN_REQUIRE(M >= 0 && N > 0 && K1 > 0 && K2 >= 0, "node_transform_ex: bad shape M=%d N=%d K1=%d K2=%d", M, N, K1, K2);
  if (M == 0) return LPGNN_OK;
  LPGNN_REQUIRE(A1 && W1 && out, "node_transform_ex: null pointer");
  LPGNN_REQUIRE(K2 == 0 || (A2 && W2), "node_transform_ex: K2=%d but A2/W2 is null", K2);
  LPGNN_REQUIRE(!epi->mask_act || (uintptr_t)epi->mask_act % 16 == 0, "node_transform_ex: mask_act must be 16-byte aligned");
  if (K2 == 0) { A2 = nullptr; W2 = nullptr; }
  EpiX epx;
  epx.mask_act = epi->mask_act;
  epx.out_scale = mscale;
  if (epi->dropout_p > 0.f) {
    epx.drop_threshold = (uint32_t)((double)epi->dropout_p * 4294967296.0);
    epx.drop_seed = epi->dropout_seed;
    epx.out_scale *= 1.f / (1.f - epi->dropout_p);
  }
  return node_transform_bf16(A1, K1, W1, A2, K2, W2, bias, M, N, out, 0, (epi->epilogue & LPGNN_EPI_RELU) ? 1 : 0, nullptr,
                             nullptr, 1, (cudaStream_t)stream, epx);
}

// Training forward of the LAST hidden layer: transform + ReLU + dropout + 16-bit store as lpgnn_node_transform_ex, with the
// basis-status head accumulated in the same epilogue on the values that are stored (after dropout).
extern "C" int lpgnn_node_transform_head_train(const void* A1, int32_t K1, const void* W1, const void* A2, int32_t K2,
                                               const void* W2, const float* bias, int32_t M, int32_t N, void* out,
                                               const lpgnn_epilogue_args* epi, const float* head_w, float* head_partial,
                                               lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(epi, "node_transform_head_train: null epilogue arguments");
  LPGNN_REQUIRE(epi->dropout_p >= 0.f && epi->dropout_p < 1.f, "node_transform_head_train: dropout_p=%f outside [0,1)", epi->dropout_p);
  LPGNN_REQUIRE(M >= 0 && N > 0 && K1 > 0 && K2 >= 0, "node_transform_head_train: bad shape M=%d N=%d K1=%d K2=%d", M, N, K1, K2);
  if (M == 0) return LPGNN_OK;
  LPGNN_REQUIRE(A1 && W1 && out && head_w && head_partial, "node_transform_head_train: null pointer");
  LPGNN_REQUIRE(K2 == 0 || (A2 && W2), "node_transform_head_train: K2=%d but A2/W2 is null", K2);
  LPGNN_REQUIRE(!epi->mask_act, "node_transform_head_train: the keep-mask epilogue belongs to the backward pass");
  if (K2 == 0) { A2 = nullptr; W2 = nullptr; }
  EpiX epx;
  if (epi->dropout_p > 0.f) {
    epx.drop_threshold = (uint32_t)((double)epi->dropout_p * 4294967296.0);
    epx.drop_seed = epi->dropout_seed;
    epx.out_scale = 1.f / (1.f - epi->dropout_p);
  }
  return node_transform_bf16(A1, K1, W1, A2, K2, W2, bias, M, N, out, 0, (epi->epilogue & LPGNN_EPI_RELU) ? 1 : 0, head_w,
                             head_partial, 1, (cudaStream_t)stream, epx);
}

// two partial slices per column tile (the tile's columns are drained by two warps per row)
extern "C" int32_t lpgnn_node_transform_head_parts(int32_t N) { return 2 * (N % 256 == 0 ? N / 256 : (N % 128 == 0 ? N / 128 : N / 64)); }

extern "C" int lpgnn_node_transform_head(const void* A1, int32_t K1, const void* W1, const void* A2, int32_t K2,
                                         const void* W2, const float* bias, int32_t M, int32_t N, void* out,
                                         int epilogue, const float* head_w, float* head_partial,
                                         lpgnn_stream_t stream) {
  return lpgnn_node_transform_head_ex(A1, K1, W1, A2, K2, W2, bias, M, N, out, LPGNN_BF16, epilogue, head_w, head_partial,
                                      stream);
}

extern "C" int lpgnn_node_transform_head_ex(const void* A1, int32_t K1, const void* W1, const void* A2, int32_t K2,
                                            const void* W2, const float* bias, int32_t M, int32_t N, void* out,
                                            int dtype, int epilogue, const float* head_w, float* head_partial,
                                            lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(is_16bit(dtype), "node_transform_head: dtype %d is not a tensor-core operand type", dtype);
  LPGNN_REQUIRE(M >= 0 && N > 0 && K1 > 0 && K2 >= 0, "node_transform_head: bad shape M=%d N=%d K1=%d K2=%d", M, N, K1, K2);
  if (M == 0) return LPGNN_OK;
  LPGNN_REQUIRE(A1 && W1 && head_w && head_partial, "node_transform_head: null pointer");
  LPGNN_REQUIRE(K2 == 0 || (A2 && W2), "node_transform_head: K2=%d but A2/W2 is null", K2);
  if (K2 == 0) { A2 = nullptr; W2 = nullptr; }
  return node_transform_bf16(A1, K1, W1, A2, K2, W2, bias, M, N, out, dtype == LPGNN_F16 ? 2 : 0,
                             (epilogue & LPGNN_EPI_RELU) ? 1 : 0, head_w, head_partial, 1, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------- split-K GEMM
namespace lpgnn {
__global__ void splitk_reduce_kernel(const float4* __restrict__ partial, int splits, int64_t quads, float4* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= quads) return;
  float4 acc = partial[i];
  for (int s = 1; s < splits; ++s) {  // fixed order: deterministic
    const float4 v = partial[(int64_t)s * quads + i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  out[i] = acc;
}
}  // namespace lpgnn

namespace lpgnn {
int gemm_tc_mn(const void* A, const void* B, int64_t Kred, int M, int N, void* out, int ksplit, cudaStream_t st);
}

// dW[N_out, K_in] = dY[Mn, N_out]^T * X[Mn, K_in]: weight gradient straight from the row-major activations
// (MN-major tcgen05 operands: no transposed copies), split-K over the node dimension, deterministic reduce.
extern "C" size_t lpgnn_wgrad_workspace_bytes(int64_t Mn, int32_t N_out, int32_t K_in) {
  return (size_t)lpgnn_gemm_tn_splits(N_out, K_in, (int32_t)((Mn + 63) / 64 * 64)) * (size_t)N_out * (size_t)K_in * sizeof(float);
}

extern "C" int lpgnn_wgrad(const void* dY, const void* X, int64_t Mn, int32_t N_out, int32_t K_in, float* out,
                           void* workspace, size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(Mn > 0 && N_out > 0 && K_in > 0 && dY && X && out, "wgrad: bad arguments");
  const int splits = lpgnn_gemm_tn_splits(N_out, K_in, (int32_t)((Mn + 63) / 64 * 64));
  cudaStream_t st = (cudaStream_t)stream;
  if (splits == 1) return gemm_tc_mn(dY, X, Mn, N_out, K_in, out, 1, st);
  if (!workspace || workspace_bytes < lpgnn_wgrad_workspace_bytes(Mn, N_out, K_in)) {
    set_error("wgrad: workspace too small");
    return LPGNN_EWORKSPACE;
  }
  LPGNN_REQUIRE(((int64_t)N_out * K_in) % 4 == 0 && (uintptr_t)workspace % 16 == 0, "wgrad: N_out*K_in must be a multiple of 4");
  if (int rc = gemm_tc_mn(dY, X, Mn, N_out, K_in, workspace, splits, st)) return rc;
  const int64_t quads = (int64_t)N_out * K_in / 4;
  splitk_reduce_kernel<<<ceil_div(quads, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(workspace), splits, quads,
                                                            reinterpret_cast<float4*>(out));
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int32_t lpgnn_gemm_tn_splits(int32_t M, int32_t N, int32_t K) {
  // enough (tile, slice) work items to fill the machine, at least 8 K-blocks of 64 per slice
  const int BN = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  const int tiles = ((M + 127) / 128) * (N / BN);
  // (tile, K-slice) work items all cost the same, so the GEMM runs in ceil(items / SMs) equal waves: pick the slice
  // count whose last wave is fullest (e.g. 32 tiles on 148 SMs: 9 slices = 288 items = 1.95 waves, not 10 = 2.16),
  // with a small price per slice for the partial tiles it writes and the reduce re-reads; >= 8 K-blocks per slice.
  const int sms = sm_count();
  const int max_by_k = K / 64 / 8;
  int splits = 1;
  double best = -1.0;
  for (int sp = 1; sp <= 32 && sp <= (max_by_k > 1 ? max_by_k : 1); ++sp) {
    const int items = tiles * sp;
    const int waves = (items + sms - 1) / sms;
    const double score = (double)items / ((double)waves * sms) - 0.004 * sp;
    if (score > best) { best = score; splits = sp; }
  }
  // every slice must own at least one K block: re-derive the slice count from the per-slice block count
  const int kblocks = K / 64;
  const int kb_per = (kblocks + splits - 1) / splits;
  return (kblocks + kb_per - 1) / kb_per;
}

extern "C" size_t lpgnn_gemm_tn_workspace_bytes(int32_t M, int32_t N, int32_t K) {
  return (size_t)lpgnn_gemm_tn_splits(M, N, K) * (size_t)M * (size_t)N * sizeof(float);
}

extern "C" int lpgnn_gemm_tn(const void* A, const void* B, int32_t M, int32_t N, int32_t K, float* out, void* workspace,
                             size_t workspace_bytes, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && out, "gemm_tn: bad arguments");
  const int splits = lpgnn_gemm_tn_splits(M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  if (splits == 1) return node_transform_bf16(A, K, B, nullptr, 0, nullptr, nullptr, M, N, out, 1, 0, nullptr, nullptr, 1, st);
  if (workspace_bytes < lpgnn_gemm_tn_workspace_bytes(M, N, K) || !workspace) {
    set_error("gemm_tn: workspace too small");
    return LPGNN_EWORKSPACE;
  }
  LPGNN_REQUIRE(((int64_t)M * N) % 4 == 0 && (uintptr_t)workspace % 16 == 0 && (uintptr_t)out % 16 == 0,
                "gemm_tn: M*N must be a multiple of 4 and buffers 16-byte aligned");
  if (int rc = node_transform_bf16(A, K, B, nullptr, 0, nullptr, nullptr, M, N, workspace, 1, 0, nullptr, nullptr, splits, st))
    return rc;
  const int64_t quads = (int64_t)M * N / 4;
  splitk_reduce_kernel<<<ceil_div(quads, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(workspace), splits, quads,
                                                            reinterpret_cast<float4*>(out));
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

// ---------------------------------------------------------------------------------------------- host->device bursts
// Enqueues `count` independent pinned-host -> device copies with one call (a plain loop of cudaMemcpyAsync):
// the per-copy cost drops from a Python round trip to a native call, which is what bounds sweeps over small LPs.
extern "C" int lpgnn_copy_many_h2d(const uint64_t* dst_ptrs, const uint64_t* src_ptrs, const uint64_t* nbytes,
                                   int32_t count, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(count >= 0 && (count == 0 || (dst_ptrs && src_ptrs && nbytes)), "copy_many_h2d: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < count; ++i) {
    if (nbytes[i] == 0) continue;
    LPGNN_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<void*>(dst_ptrs[i]), reinterpret_cast<const void*>(src_ptrs[i]),
                                  (size_t)nbytes[i], cudaMemcpyHostToDevice, st));
  }
  return LPGNN_OK;
}
