// common.cuh — small device helpers shared by the ALS / RMSE kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define YCNR_MAX_DST 8  // local replica + up to 7 NVLink peers

// Destinations of a solved row: the local replica first, then peer-mapped replicas
// (the fused all-gather: EmfMaster.js:711-723 streamed rows to every node instead).
struct DstList {
  float* p[YCNR_MAX_DST];
  int n;
};

// A step's rows as the worker sees them (EmfWorker.js:214-219): row r covers
// indx/vals[start[r] .. start[r]+len[r]).
struct RowsView {
  const int32_t* __restrict__ row_ids;
  const int64_t* __restrict__ row_start;
  const int32_t* __restrict__ row_len;
  const int32_t* __restrict__ indx;
  const float* __restrict__ vals;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte async copy global->shared; src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
