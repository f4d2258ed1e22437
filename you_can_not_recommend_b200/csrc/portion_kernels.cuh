// portion_kernels.cuh — device-side unpacking of the upstream portion header (sm_100a).
//
// A portion arrives in the reference's wire format (EmfMaster.js:571-614 producer, EmfWorker.js:176-219
// consumer): rows = [R, rowId_0, n_0, rowId_1, n_1, ...], ratings concatenated in row order.  The worker
// walks it with a running offset (EmfWorker.js:217-219); here the raw header is DMA'd as it is and three
// small kernels turn it into the row arrays the compute kernels read (row_ids, row_len, row_start =
// exclusive prefix sum of n), so the host does no per-row work on the RMSE path.
#pragma once
#include "common.cuh"

namespace ycnr {

constexpr int kUnpackThreads = 256;
constexpr int kUnpackRowsPerBlock = 2048;   // 8 passes of 256 rows

// block_sums[b] = sum of n over rows [b*2048, (b+1)*2048)
__global__ void __launch_bounds__(kUnpackThreads) header_block_sums_kernel(const int32_t* __restrict__ hdr, int R,
                                                                           int64_t* __restrict__ block_sums) {
  __shared__ int64_t ws[kUnpackThreads / 32];
  const int base = blockIdx.x * kUnpackRowsPerBlock;
  int64_t s = 0;
  for (int r = base + threadIdx.x; r < min(R, base + kUnpackRowsPerBlock); r += kUnpackThreads)
    s += __ldg(hdr + 2 + 2 * (size_t)r);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < kUnpackThreads / 32; ++w) t += ws[w];
    block_sums[blockIdx.x] = t;
  }
}

// in-place exclusive scan of block_sums[0..nb) by one CTA of 1024 threads
__global__ void __launch_bounds__(1024) header_scan_blocks_kernel(int64_t* __restrict__ block_sums, int nb) {
  __shared__ int64_t ws[32];
  const int per = (nb + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(nb, lo + per);
  int64_t s = 0;
  for (int i = lo; i < hi; ++i) s += block_sums[i];
  int64_t inc = s;   // inclusive scan over threads
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += v;
  }
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    int64_t w = ws[threadIdx.x], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t v = __shfl_up_sync(0xffffffffu, winc, o);
      if (threadIdx.x >= o) winc += v;
    }
    ws[threadIdx.x] = winc - w;   // exclusive warp offsets
  }
  __syncthreads();
  int64_t run = ws[threadIdx.x >> 5] + inc - s;
  for (int i = lo; i < hi; ++i) {
    const int64_t v = block_sums[i];
    block_sums[i] = run;
    run += v;
  }
}

// row_ids / row_len / row_start of rows [b*2048, (b+1)*2048), row_start continuing from block_offs[b]
__global__ void __launch_bounds__(kUnpackThreads) header_unpack_kernel(const int32_t* __restrict__ hdr, int R,
                                                                       const int64_t* __restrict__ block_offs,
                                                                       int32_t* __restrict__ row_ids,
                                                                       int32_t* __restrict__ row_len,
                                                                       int64_t* __restrict__ row_start) {
  __shared__ int64_t ws[kUnpackThreads / 32];
  __shared__ int64_t carry;
  const int base = blockIdx.x * kUnpackRowsPerBlock;
  if (threadIdx.x == 0) carry = block_offs[blockIdx.x];
  __syncthreads();
  for (int p0 = 0; p0 < kUnpackRowsPerBlock; p0 += kUnpackThreads) {
    const int r = base + p0 + threadIdx.x;
    int32_t id = 0, n = 0;
    if (r < R) {
      // the driver places the header so that hdr + 1 (the first pair) is 8-byte aligned
      const int2 v = __ldg(reinterpret_cast<const int2*>(hdr + 1) + r);
      id = v.x;
      n = v.y;
    }
    int64_t inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += v;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    int64_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kUnpackThreads / 32; ++w) {
      const int64_t v = ws[w];
      if (w < (threadIdx.x >> 5)) woff += v;
      total += v;
    }
    const int64_t c = carry;
    if (r < R) {
      row_ids[r] = id;
      row_len[r] = n;
      row_start[r] = c + woff + inc - n;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
}

}  // namespace ycnr
