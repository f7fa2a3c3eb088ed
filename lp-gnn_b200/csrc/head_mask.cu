// (a4+a5) Basis-status head + knowledge masking in one pass over the hidden activations.
//
// Replaces torch.nn.Linear(hids,3) (reference arch.py:190: cuBLAS skinny GEMM + bias) and
// add_knowledge (reference arch.py:129-141: F.normalize, *10 and four boolean index_put_) -- ~12
// library kernels -- with one kernel: a warp reads one H-wide row with 128-bit loads, forms the 3
// dot products (weights staged in shared memory), reduces with shuffles, then normalises, scales
// by 10 and applies the +-inf-bound mask read from the node features.
// HBM-bound: rows * (Hdim*sizeof(h) + 4*q + 12) bytes.
#include "common.cuh"

namespace lpgnn {
namespace {

constexpr int kThreads = 256;
constexpr int kRegThreads = 128;  // register-resident variant: smaller blocks for register-file granularity

__device__ __forceinline__ void finish_row_tags(float r0, float r1, float r2, float tag_lo, float tag_up, int64_t row,
                                                float* __restrict__ logits) {
  const float nrm = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
  const float den = fmaxf(nrm, 1e-12f);
  float y0 = (r0 / den) * 10.f, y1 = (r1 / den) * 10.f, y2 = (r2 / den) * 10.f;
  if (tag_lo != 0.f) y0 -= 10.f;
  if (tag_up != 0.f) y2 -= 10.f;
  logits[row * 3 + 0] = y0;
  logits[row * 3 + 1] = y1;
  logits[row * 3 + 2] = y2;
}

__device__ __forceinline__ void finish_row(float r0, float r1, float r2, const float* __restrict__ feas, int q,
                                           int64_t row, float* __restrict__ logits) {
  // F.normalize: x / max(||x||_2, eps), eps = 1e-12 ; then * 10 (arch.py:134-135)
  const float nrm = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
  const float den = fmaxf(nrm, 1e-12f);
  float y0 = (r0 / den) * 10.f, y1 = (r1 / den) * 10.f, y2 = (r2 / den) * 10.f;
  // mask: tag columns q-3 (lower bound) and q-1 (upper bound) are non-zero for +-inf (arch.py:130-140)
  if (__ldg(feas + row * q + (q - 3)) != 0.f) y0 -= 10.f;
  if (__ldg(feas + row * q + (q - 1)) != 0.f) y2 -= 10.f;
  logits[row * 3 + 0] = y0;
  logits[row * 3 + 1] = y1;
  logits[row * 3 + 2] = y2;
}

// eight 16-bit floats of a 16-byte chunk -> fp32 (T = __nv_bfloat16 or __half; never instantiated for float)
template <typename T, int E>
__device__ __forceinline__ void unpack8(const uint4& v, float (&x)[E]) {
  if constexpr (sizeof(T) == 2) {
    const float2 a = Half16<T>::unpack(v.x), b = Half16<T>::unpack(v.y), c = Half16<T>::unpack(v.z), d = Half16<T>::unpack(v.w);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
head_mask_kernel(const T* __restrict__ H, int32_t rows, int32_t Hdim, const float* __restrict__ W,
                 const float* __restrict__ b, const float* __restrict__ feas, int q, float* __restrict__ logits,
                 float* __restrict__ raw_out) {
  extern __shared__ __align__(16) float w_s[];  // [3][Hdim]
  for (int i = threadIdx.x; i < 3 * Hdim; i += kThreads) w_s[i] = __ldg(W + i);
  __syncthreads();
  constexpr int E = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (kThreads / 32);
  const int chunks = Hdim / E;
  const float b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
  for (int64_t row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < rows; row += warps_total) {
    const uint4* hrow = reinterpret_cast<const uint4*>(H + row * Hdim);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int c = lane; c < chunks; c += 32) {
      const uint4 v = __ldg(hrow + c);
      float x[E];
      if constexpr (sizeof(T) == 4) {
        x[0] = __uint_as_float(v.x); x[1] = __uint_as_float(v.y); x[2] = __uint_as_float(v.z); x[3] = __uint_as_float(v.w);
      } else {
        unpack8<T>(v, x);
      }
      const float* w0 = w_s + c * E;
#pragma unroll
      for (int k = 0; k < E; ++k) {
        a0 = fmaf(x[k], w0[k], a0);
        a1 = fmaf(x[k], w0[Hdim + k], a1);
        a2 = fmaf(x[k], w0[2 * Hdim + k], a2);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if (lane == 0) {
      a0 += b0; a1 += b1; a2 += b2;
      if (raw_out) { raw_out[row * 3] = a0; raw_out[row * 3 + 1] = a1; raw_out[row * 3 + 2] = a2; }
      finish_row(a0, a1, a2, feas, q, row, logits);
    }
  }
}

// Fast path.  Weights are staged in shared memory in a lane-major layout
//   ws[((j*CH + c)*E + k)*32 + lane] = W[j][(lane + 32*c)*E + k]
// so every weight read is a conflict-free LDS (the row-major layout of the generic kernel above is
// bank-conflict bound: lanes read 32-byte-strided slices).  A warp handles TWO adjacent rows per
// iteration, sharing each weight read between them and keeping 2*CH 128-bit loads in flight per lane;
// registers stay low enough for 24+ resident warps per SM.
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
head_mask_fast_kernel(const T* __restrict__ H, int32_t rows, int32_t Hdim, const float* __restrict__ W,
                      const float* __restrict__ b, const float* __restrict__ feas, int q, float* __restrict__ logits,
                      float* __restrict__ raw_out) {
  constexpr int E = 16 / sizeof(T);
  __shared__ float ws[3 * CH * E * 32];
  const int lane = threadIdx.x & 31;
  const int chunks = Hdim / E;
  for (int i = threadIdx.x; i < 3 * CH * E * 32; i += kThreads) {
    const int l = i & 31, k = (i >> 5) % E, c = ((i >> 5) / E) % CH, j = (i >> 5) / (E * CH);
    const int ch = l + 32 * c;
    ws[i] = (ch < chunks) ? __ldg(W + (int64_t)j * Hdim + ch * E + k) : 0.f;
  }
  __syncthreads();
  const float b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
  const int warps_total = gridDim.x * (kThreads / 32);
  for (int64_t row = 2 * (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)); row < rows; row += 2 * warps_total) {
    const bool has1 = row + 1 < rows;
    const uint4* hrow0 = reinterpret_cast<const uint4*>(H + row * Hdim);
    const uint4* hrow1 = reinterpret_cast<const uint4*>(H + (row + (has1 ? 1 : 0)) * Hdim);
    uint4 v0[CH], v1[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int ch = lane + 32 * c;
      v0[c] = (ch < chunks) ? __ldg(hrow0 + ch) : make_uint4(0, 0, 0, 0);
      v1[c] = (ch < chunks) ? __ldg(hrow1 + ch) : make_uint4(0, 0, 0, 0);
    }
    // the mask tags are fetched by lanes 0..3 alongside the rows, not after the reduction
    float tag = 0.f;
    if (lane < 4) {
      const int64_t r = row + (((lane >> 1) && has1) ? 1 : 0);
      tag = __ldg(feas + r * q + ((lane & 1) ? (q - 1) : (q - 3)));
    }
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float x[E], y[E];
      if constexpr (sizeof(T) == 4) {
        x[0] = __uint_as_float(v0[c].x); x[1] = __uint_as_float(v0[c].y);
        x[2] = __uint_as_float(v0[c].z); x[3] = __uint_as_float(v0[c].w);
        y[0] = __uint_as_float(v1[c].x); y[1] = __uint_as_float(v1[c].y);
        y[2] = __uint_as_float(v1[c].z); y[3] = __uint_as_float(v1[c].w);
      } else {
        unpack8<T>(v0[c], x);
        unpack8<T>(v1[c], y);
      }
#pragma unroll
      for (int k = 0; k < E; ++k) {
        const float w0 = ws[((0 * CH + c) * E + k) * 32 + lane];
        const float w1 = ws[((1 * CH + c) * E + k) * 32 + lane];
        const float w2 = ws[((2 * CH + c) * E + k) * 32 + lane];
        a[0] = fmaf(x[k], w0, a[0]); a[1] = fmaf(x[k], w1, a[1]); a[2] = fmaf(x[k], w2, a[2]);
        a[3] = fmaf(y[k], w0, a[3]); a[4] = fmaf(y[k], w1, a[4]); a[5] = fmaf(y[k], w2, a[5]);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int j = 0; j < 6; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], off);
    const float t_up0 = __shfl_sync(0xffffffffu, tag, 1);
    const float t_lo1 = __shfl_sync(0xffffffffu, tag, 2);
    const float t_up1 = __shfl_sync(0xffffffffu, tag, 3);
    if (lane == 0) {
      const float r0 = a[0] + b0, r1 = a[1] + b1, r2 = a[2] + b2;
      if (raw_out) { raw_out[row * 3] = r0; raw_out[row * 3 + 1] = r1; raw_out[row * 3 + 2] = r2; }
      finish_row_tags(r0, r1, r2, tag, t_up0, row, logits);
    }
    if (lane == 1 && has1) {
      const float r0 = a[3] + b0, r1 = a[4] + b1, r2 = a[5] + b2;
      if (raw_out) { raw_out[(row + 1) * 3] = r0; raw_out[(row + 1) * 3 + 1] = r1; raw_out[(row + 1) * 3 + 2] = r2; }
      finish_row_tags(r0, r1, r2, t_lo1, t_up1, row + 1, logits);
    }
  }
}

template <typename T, int CH>
void launch_reg(const void* H, int32_t rows, int32_t Hdim, const float* W, const float* b, const float* feas, int q,
                float* logits, float* raw_out, cudaStream_t st) {
  const int grid = min(ceil_div(rows, 2 * (kThreads / 32)), sm_count() * 4);
  head_mask_fast_kernel<T, CH><<<grid, kThreads, 0, st>>>(reinterpret_cast<const T*>(H), rows, Hdim, W, b, feas, q,
                                                          logits, raw_out);
}

__global__ void add_knowledge_kernel(const float* __restrict__ in, int32_t rows, const float* __restrict__ feas,
                                     int q, float* __restrict__ out) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= rows) return;
  finish_row(in[row * 3], in[row * 3 + 1], in[row * 3 + 2], feas, q, row, out);
}

__global__ void head_finish_kernel(const float* __restrict__ partial, int nparts, int32_t rows,
                                   const float* __restrict__ b, const float* __restrict__ feas, int q,
                                   float* __restrict__ logits, float* __restrict__ raw_out) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float r0 = __ldg(b), r1 = __ldg(b + 1), r2 = __ldg(b + 2);
  for (int p = 0; p < nparts; ++p) {  // fixed order: deterministic
    const float* pp = partial + ((int64_t)p * rows + row) * 3;
    r0 += pp[0]; r1 += pp[1]; r2 += pp[2];
  }
  if (raw_out) { raw_out[row * 3] = r0; raw_out[row * 3 + 1] = r1; raw_out[row * 3 + 2] = r2; }
  finish_row(r0, r1, r2, feas, q, row, logits);
}

}  // namespace
}  // namespace lpgnn

using namespace lpgnn;

extern "C" int lpgnn_head_finish_ex(const float* partial, int32_t nparts, int32_t rows, const float* b, const float* feas,
                                    int32_t q, float* logits, float* raw_out, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && nparts >= 1 && q >= 3, "head_finish: bad shape rows=%d nparts=%d q=%d", rows, nparts, q);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(partial && b && feas && logits, "head_finish: null pointer");
  head_finish_kernel<<<ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(partial, nparts, rows, b, feas, q, logits, raw_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_head_finish(const float* partial, int32_t nparts, int32_t rows, const float* b, const float* feas,
                                 int32_t q, float* logits, lpgnn_stream_t stream) {
  return lpgnn_head_finish_ex(partial, nparts, rows, b, feas, q, logits, nullptr, stream);
}

extern "C" int lpgnn_head_mask(const void* H, int h_dtype, int32_t rows, int32_t Hdim, const float* W, const float* b,
                               const float* feas, int32_t q, float* logits, float* raw_out, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && Hdim > 0 && q >= 3, "head_mask: bad shape rows=%d Hdim=%d q=%d", rows, Hdim, q);
  LPGNN_REQUIRE(dtype_ok(h_dtype), "head_mask: bad dtype %d", h_dtype);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(H && W && b && feas && logits, "head_mask: null pointer");
  const int esz = h_dtype == LPGNN_F32 ? 4 : 2;
  LPGNN_REQUIRE((Hdim * esz) % 16 == 0 && (uintptr_t)H % 16 == 0, "head_mask: H rows must be 16-byte multiples/aligned");
  LPGNN_REQUIRE(3 * Hdim * 4 <= 96 * 1024, "head_mask: Hdim=%d too large for the shared-memory weight stage", Hdim);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = Hdim * esz / 16;
  const int ch = (chunks + 31) / 32;
  const bool f32 = h_dtype == LPGNN_F32, f16 = h_dtype == LPGNN_F16;
  if (ch <= 8) {
#define LPGNN_HEAD(CHV)                                                                                   \
  (f32 ? launch_reg<float, CHV>(H, rows, Hdim, W, b, feas, q, logits, raw_out, st)                        \
       : f16 ? launch_reg<__half, CHV>(H, rows, Hdim, W, b, feas, q, logits, raw_out, st)                 \
             : launch_reg<__nv_bfloat16, CHV>(H, rows, Hdim, W, b, feas, q, logits, raw_out, st))
    if (ch <= 1) LPGNN_HEAD(1); else if (ch <= 2) LPGNN_HEAD(2); else if (ch <= 4) LPGNN_HEAD(4); else LPGNN_HEAD(8);
#undef LPGNN_HEAD
  } else {
    const int grid = min(ceil_div(rows, kThreads / 32), sm_count() * 8);
    const size_t smem = (size_t)3 * Hdim * sizeof(float);
    if (f32) {
      LPGNN_CUDA_OK(cudaFuncSetAttribute(head_mask_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      head_mask_kernel<float><<<grid, kThreads, smem, st>>>(reinterpret_cast<const float*>(H), rows, Hdim, W, b, feas,
                                                            q, logits, raw_out);
    } else if (f16) {
      LPGNN_CUDA_OK(cudaFuncSetAttribute(head_mask_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      head_mask_kernel<__half><<<grid, kThreads, smem, st>>>(reinterpret_cast<const __half*>(H), rows, Hdim, W, b, feas, q,
                                                             logits, raw_out);
    } else {
      LPGNN_CUDA_OK(cudaFuncSetAttribute(head_mask_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         96 * 1024));
      head_mask_kernel<__nv_bfloat16><<<grid, kThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(H), rows,
                                                                    Hdim, W, b, feas, q, logits, raw_out);
    }
  }
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}

extern "C" int lpgnn_add_knowledge(const float* logits_in, int32_t rows, const float* feas, int32_t q,
                                   float* logits_out, lpgnn_stream_t stream) {
  if (int rc = check_device()) return rc;
  LPGNN_REQUIRE(rows >= 0 && q >= 3, "add_knowledge: bad shape rows=%d q=%d", rows, q);
  if (rows == 0) return LPGNN_OK;
  LPGNN_REQUIRE(logits_in && feas && logits_out, "add_knowledge: null pointer");
  add_knowledge_kernel<<<ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(logits_in, rows, feas, q, logits_out);
  LPGNN_LAUNCH_OK();
  count_launches(1);
  return LPGNN_OK;
}
