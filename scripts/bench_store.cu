// Write-bandwidth microbenchmark (context for the write-dominated input layer): how fast can 205 MB be WRITTEN by
//   0: st.global.v4 (default), 1: st.global.cs.v4 (streaming), 2: st.global.L1::no_allocate, 3: TMA bulk store from shared
// memory (cp.async.bulk.global.shared::cta, 4 KB per bulk op), 4: cudaMemsetAsync.  Build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o scripts/bin/bench_store scripts/bench_store.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void store_kernel(uint4* out, size_t n16) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    if (MODE == 0) out[i] = v;
    if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (MODE == 2) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}

// The input layer's store patterns: a warp owns 16-row tiles of a [rows, 1024] 16-bit matrix (2 KB per row).
//   HALF = true : per 32-column chunk two stores of 8 rows x 64 B (lane (g,t): row g / g+8, bytes 64c + 16t) -- every
//                 128-byte line is written by two different instructions (what mma.sync's accumulator layout gives)
//   HALF = false: per 64-column chunk four stores of 4 rows x 128 B (full lines per instruction)
template <bool HALF>
__global__ void tile_store_kernel(uint8_t* out, int rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
  const long warps = (long)gridDim.x * (blockDim.x >> 5);
  for (long row0 = ((long)blockIdx.x * (blockDim.x >> 5) + warp) * 16; row0 < rows; row0 += warps * 16) {
    if (HALF) {
      uint8_t* ra = out + (row0 + g) * 2048 + 16 * t;
      uint8_t* rb = ra + 8 * 2048;
#pragma unroll 2
      for (int c = 0; c < 32; ++c) {
        if (row0 + g < rows) *reinterpret_cast<uint4*>(ra + 64 * c) = v;
        if (row0 + g + 8 < rows) *reinterpret_cast<uint4*>(rb + 64 * c) = v;
      }
    } else {
      uint8_t* r0 = out + (row0 + (lane >> 3)) * 2048 + 16 * (lane & 7);
#pragma unroll 2
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (row0 + (lane >> 3) + 4 * k < rows) *reinterpret_cast<uint4*>(r0 + (long)k * 4 * 2048 + 128 * c) = v;
      }
    }
  }
}

// each CTA fills a 16 KB shared tile once, then streams it to consecutive 16 KB chunks of the output with bulk stores
__global__ void tma_store_kernel(uint8_t* out, size_t bytes) {
  extern __shared__ __align__(128) uint8_t tile[];
  constexpr int kTile = 16384;
  for (int i = threadIdx.x; i < kTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(tile)[i] = make_uint4(i, blockIdx.x, 1, 2);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
    int pending = 0;
    for (size_t off = (size_t)blockIdx.x * kTile; off + kTile <= bytes; off += (size_t)gridDim.x * kTile) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + off), "r"(src), "r"(kTile) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++pending >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); pending = 4; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const size_t bytes = 100000ull * 1024 * 2;   // the vars-side conv1 output of BASELINE C2 in 16-bit
  uint8_t* buf; cudaMalloc(&buf, bytes);
  uint8_t* flush; cudaMalloc(&flush, 256u << 20);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int mode = 0; mode < 7; ++mode) {
    for (int grid_mult : {2, 8, 32}) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        cudaMemsetAsync(flush, rep, 256u << 20);
        cudaEventRecord(a);
        const int grid = sms * grid_mult;
        if (mode == 0) store_kernel<0><<<grid, 256>>>((uint4*)buf, bytes / 16);
        if (mode == 1) store_kernel<1><<<grid, 256>>>((uint4*)buf, bytes / 16);
        if (mode == 2) store_kernel<2><<<grid, 256>>>((uint4*)buf, bytes / 16);
        if (mode == 3) tma_store_kernel<<<grid, 128, 16384>>>(buf, bytes);
        if (mode == 4) cudaMemsetAsync(buf, 7, bytes);
        if (mode == 5) tile_store_kernel<true><<<sms * 4, 256>>>(buf, 100000);
        if (mode == 6) tile_store_kernel<false><<<sms * 4, 256>>>(buf, 100000);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
      }
      const char* names[] = {"st.global.v4", "st.global.cs.v4", "st.global.L1::no_allocate.v4", "TMA bulk store 16 KB", "cudaMemsetAsync",
                             "16-row tiles, 8 rows x 64 B", "16-row tiles, 4 rows x 128 B"};
      printf("%-30s grid %4d x SMs: %7.1f us  %6.0f GB/s\n", names[mode], grid_mult, best * 1e3f, bytes / best / 1e6f);
      if (mode >= 4) break;
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
