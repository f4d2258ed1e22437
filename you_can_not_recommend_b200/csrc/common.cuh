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
  // per-portion path only: device flag raised by validate_cols_kernel when a column id of a queued portion
  // lies outside the fixed matrix; the compute kernels of the step then leave without touching memory
  const int32_t* __restrict__ guard;
};

__device__ __forceinline__ bool rows_poisoned(const RowsView& r) {
  return r.guard != nullptr && *reinterpret_cast<const volatile int32_t*>(r.guard) != 0;
}

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

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP): one instruction moves one whole factor row global -> shared
// through the copy engine and reports its bytes to an mbarrier; no per-16-byte cp.async issue slots.
__device__ __forceinline__ void tma_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(bar)
               : "memory");
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

// Column ids of a portion in the upstream wire format are trusted by the reference worker (EmfWorker.js:217-228
// indexes the fixed matrix with them unchecked and degrades to NaN); here an id outside [0, limit) would be an
// out-of-bounds gather, so the per-portion path checks them on the device right after the upload.
__global__ void __launch_bounds__(256) validate_cols_kernel(const int32_t* __restrict__ indx, int64_t n, int32_t limit,
                                                            int32_t* __restrict__ flag) {
  bool bad = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    bad |= (uint32_t)__ldg(indx + e) >= (uint32_t)limit;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
