// rmse_kernels.cuh — train/validation/test RMSE reduction for sm_100a.
//
// Restates EmfWorker.mw_calcRmsePortion (lib/emf/EmfWorker.js:266-315): for every rating
// of every user row  pred = fp32 dot(U[u], V[i]) + globalAvgShift  (EmfBase.js:815-827),
// then in double (JS numbers): rSumDiff2 += (r - pred)^2, rSum += pred, rCnt += 1.
//
// Kernel 1: one 8-lane group per user row (four rows per warp — the validate / test rows of the BASELINE shapes
//           hold 3 to 7 ratings, a whole warp per row idled three quarters of its lanes); the 8 lanes split the
//           factor row in float4 chunks (coalesced 128-byte segments), the user's row stays in registers, two
//           ratings are in flight per group, the fp32 dot is shuffle-reduced inside the group and the group
//           leader accumulates in fp64 in rating order; per-row sums go to row_sums[R][3].
// Kernel 2: one CTA per portion sums its rows in a fixed order (deterministic) into
//           portion_sums[P][4] — per-portion partials are what quirk Q7 needs
//           (EmfMaster.js:777-783 uses the LAST portion's rSum/rCnt).  The fourth sum, the ratings themselves,
//           lets the host derive a pass with another globalAvgShift without touching the device again:
//           sum (r - p - d)^2 = sum (r - p)^2 - 2 d (sum r - sum p) + n d^2  (the third RMSE pass of an iteration).
#pragma once
#include "common.cuh"

namespace ycnr {

struct RmseArgs {
  RowsView rows;
  const float* __restrict__ U;
  const float* __restrict__ V;
  int k;
  int n_rows;
  double shift;
  double* __restrict__ row_sums;  // [n_rows][3] = {sum diff^2, sum pred, sum rating}
  // optional: rows in processing order, longest first, so that the four groups of a warp walk rows of (nearly)
  // the same length — validate / test rows are power-law distributed and a warp lives as long as its longest row
  const int32_t* __restrict__ order;
};

constexpr int kRmseRowSums = 3;
constexpr int kRmsePortionSums = 4;   // rSumDiff2, rCnt, rSum, sum of ratings

// NQ = float4 chunks of the factor row per lane (k <= 32 NQ); NQ = 0: any k, scalar loads
template <int NQ>
__global__ void __launch_bounds__(256) rmse_rows_kernel(const RmseArgs a) {
  const int g0 = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 3);   // this 8-lane group's position
  const int lane = threadIdx.x & 31;
  const int gl = lane & 7;
  const uint32_t gmask = 0xFFu << (lane & 24);
  if (g0 >= a.n_rows) return;                // whole groups leave together
  const int g = a.order ? __ldg(a.order + g0) : g0;
  if (rows_poisoned(a.rows)) {               // a column id of the portion is out of range: no gather
    if (gl == 0) { a.row_sums[3 * (size_t)g] = 0.0; a.row_sums[3 * (size_t)g + 1] = 0.0; a.row_sums[3 * (size_t)g + 2] = 0.0; }
    return;
  }
  const int k = a.k;
  const int u = a.rows.row_ids[g];
  const int64_t beg = a.rows.row_start[g];
  const int n = a.rows.row_len[g];
  const float* uf = a.U + (size_t)u * k;
  double sd2 = 0.0, sp = 0.0, sr = 0.0;
  float4 uc[NQ > 0 ? NQ : 1];
  if (NQ > 0) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const int c = 4 * (gl + 8 * j);
      uc[j] = c < k ? __ldg(reinterpret_cast<const float4*>(uf + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  auto dot_with = [&](int it) {
    const float* vf = a.V + (size_t)it * k;
    float dot = 0.f;
    if (NQ > 0) {
      float4 y[NQ > 0 ? NQ : 1];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = 4 * (gl + 8 * q);
        y[q] = c < k ? __ldg(reinterpret_cast<const float4*>(vf + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        dot = fmaf(uc[q].x, y[q].x, dot);
        dot = fmaf(uc[q].y, y[q].y, dot);
        dot = fmaf(uc[q].z, y[q].z, dot);
        dot = fmaf(uc[q].w, y[q].w, dot);
      }
    } else {
      for (int c = gl; c < k; c += 8) dot = fmaf(__ldg(uf + c), __ldg(vf + c), dot);
    }
    return dot;
  };
  auto group_sum = [&](float v) {
    v += __shfl_xor_sync(gmask, v, 1);
    v += __shfl_xor_sync(gmask, v, 2);
    v += __shfl_xor_sync(gmask, v, 4);
    return v;
  };
  for (int j = 0; j < n; j += 2) {            // two ratings in flight; trip count uniform inside the group
    const bool two = j + 1 < n;
    const int it0 = __ldg(a.rows.indx + beg + j);
    const int it1 = two ? __ldg(a.rows.indx + beg + j + 1) : it0;
    float d0 = dot_with(it0);
    float d1 = two ? dot_with(it1) : 0.f;
    d0 = group_sum(d0);
    d1 = group_sum(d1);
    if (gl == 0) {
      const double r0 = (double)__ldg(a.rows.vals + beg + j);
      const double p0 = (double)d0 + a.shift;
      sd2 += (r0 - p0) * (r0 - p0);
      sp += p0;
      sr += r0;
      if (two) {
        const double r1 = (double)__ldg(a.rows.vals + beg + j + 1);
        const double p1 = (double)d1 + a.shift;
        sd2 += (r1 - p1) * (r1 - p1);
        sp += p1;
        sr += r1;
      }
    }
  }
  if (gl == 0) {
    a.row_sums[3 * (size_t)g] = sd2;
    a.row_sums[3 * (size_t)g + 1] = sp;
    a.row_sums[3 * (size_t)g + 2] = sr;
  }
}

// fixed-order block reduction of four running sums
__device__ __forceinline__ void rmse_block_reduce4(double (&v)[4], double (*s)[256]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i][threadIdx.x] = v[i];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i][threadIdx.x] += s[i][threadIdx.x + o];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = s[i][0];
}

// portion p covers rows [portion_first[p], portion_first[p+1])
__global__ void __launch_bounds__(256) rmse_portion_reduce_kernel(const double* __restrict__ row_sums,
                                                                  const int32_t* __restrict__ row_len,
                                                                  const int32_t* __restrict__ portion_first,
                                                                  double* __restrict__ portion_sums) {
  __shared__ double s[4][256];
  const int p = blockIdx.x;
  const int lo = portion_first[p], hi = portion_first[p + 1];
  double v[4] = {0.0, 0.0, 0.0, 0.0};   // d2, cnt, pred, rating
  for (int r = lo + threadIdx.x; r < hi; r += 256) {
    v[0] += row_sums[3 * (size_t)r];
    v[1] += (double)row_len[r];
    v[2] += row_sums[3 * (size_t)r + 1];
    v[3] += row_sums[3 * (size_t)r + 2];
  }
  rmse_block_reduce4(v, s);
  if (threadIdx.x == 0) {
    portion_sums[4 * (size_t)p] = v[0];      // rSumDiff2
    portion_sums[4 * (size_t)p + 1] = v[1];  // rCnt
    portion_sums[4 * (size_t)p + 2] = v[2];  // rSum
    portion_sums[4 * (size_t)p + 3] = v[3];  // sum of the ratings
  }
}

// Two-level variant for ONE big portion (the per-portion path with multi-million-rating portions): a single
// CTA walking millions of rows is latency-bound (measured: ~5 ms for 1.7 M rows).  Level 1: one CTA per chunk
// of kRmseChunkRows rows -> chunk_sums[c][4]; level 2: rmse_portion_reduce_chunks_kernel adds the chunks.
// Both levels add in a fixed order, so the result does not depend on scheduling.
constexpr int kRmseChunkRows = 4096;

__global__ void __launch_bounds__(256) rmse_chunk_reduce_kernel(const double* __restrict__ row_sums,
                                                                const int32_t* __restrict__ row_len, int n_rows,
                                                                double* __restrict__ chunk_sums) {
  __shared__ double s[4][256];
  const int lo = blockIdx.x * kRmseChunkRows, hi = min(n_rows, lo + kRmseChunkRows);
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = lo + threadIdx.x; r < hi; r += 256) {
    v[0] += row_sums[3 * (size_t)r];
    v[1] += (double)row_len[r];
    v[2] += row_sums[3 * (size_t)r + 1];
    v[3] += row_sums[3 * (size_t)r + 2];
  }
  rmse_block_reduce4(v, s);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) chunk_sums[4 * (size_t)blockIdx.x + i] = v[i];
  }
}

__global__ void __launch_bounds__(256) rmse_portion_reduce_chunks_kernel(const double* __restrict__ chunk_sums,
                                                                         int n_chunks,
                                                                         double* __restrict__ portion_sums) {
  __shared__ double s[4][256];
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int c = threadIdx.x; c < n_chunks; c += 256) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += chunk_sums[4 * (size_t)c + i];
  }
  rmse_block_reduce4(v, s);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) portion_sums[i] = v[i];
  }
}

}  // namespace ycnr
