// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so
// liblpgnn.so does not link libcuda directly).  Shared by the tensor-core node transforms (gemm_tc.cu, gemm_x2.cu).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace lpgnn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, K] 16-bit row-major (leading dimension ld elements), box = [box_rows, 64 elements = 128 bytes], 128-byte
// swizzle, out-of-bounds rows read as zero.
inline int make_map_16bit(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows, bool f16,
                          const char* who) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("%s: cuTensorMapEncodeTiled unavailable", who); return LPGNN_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r); return LPGNN_ECUDA; }
  return LPGNN_OK;
}

}  // namespace lpgnn
