// rmse_kernels.cuh — train/validation/test RMSE reduction for sm_100a.
//
// Restates EmfWorker.mw_calcRmsePortion (lib/emf/EmfWorker.js:266-315): for every rating
// of every user row  pred = fp32 dot(U[u], V[i]) + globalAvgShift  (EmfBase.js:815-827),
// then in double (JS numbers): rSumDiff2 += (r - pred)^2, rSum += pred, rCnt += 1.
//
// Kernel 1: one warp per user row, four 8-lane groups each taking every 4th rating;
//           the 8 lanes of a group split the factor row in float4 chunks (coalesced
//           128-byte segments), shuffle-reduce the fp32 dot, the group leader
//           accumulates in fp64; per-row sums go to row_sums[R][2].
// Kernel 2: one CTA per portion sums its rows in a fixed order (deterministic) into
//           portion_sums[P][3] — per-portion partials are what quirk Q7 needs
//           (EmfMaster.js:777-783 uses the LAST portion's rSum/rCnt).
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
  double* __restrict__ row_sums;  // [n_rows][2] = {sum diff^2, sum pred}
};

__global__ void __launch_bounds__(256) rmse_rows_kernel(const RmseArgs a) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= a.n_rows) return;
  const int k = a.k;
  const int u = a.rows.row_ids[warp];
  const int64_t beg = a.rows.row_start[warp];
  const int n = a.rows.row_len[warp];
  const float* uf = a.U + (size_t)u * k;
  const int grp = lane >> 3, gl = lane & 7;
  double sd2 = 0.0, sp = 0.0;
  const bool vec = (k & 3) == 0 && k <= 128;
  if (vec) {
    // the user's row stays in registers: lane gl of every group owns the float4 chunks gl, gl+8, gl+16, gl+24
    float4 uc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * (gl + 8 * j);
      uc[j] = c < k ? __ldg(reinterpret_cast<const float4*>(uf + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int j0 = 0; j0 < n; j0 += 4) {     // uniform trip count for the whole warp
      const int j = j0 + grp;
      const bool ok = j < n;
      float dot = 0.f;
      if (ok) {
        const int it = __ldg(a.rows.indx + beg + j);
        const float* vf = a.V + (size_t)it * k;
        float4 y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = 4 * (gl + 8 * q);
          y[q] = c < k ? __ldg(reinterpret_cast<const float4*>(vf + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dot = fmaf(uc[q].x, y[q].x, dot);
          dot = fmaf(uc[q].y, y[q].y, dot);
          dot = fmaf(uc[q].z, y[q].z, dot);
          dot = fmaf(uc[q].w, y[q].w, dot);
        }
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      if (ok && gl == 0) {
        const double pred = (double)dot + a.shift;
        const double diff = (double)__ldg(a.rows.vals + beg + j) - pred;
        sd2 += diff * diff;
        sp += pred;
      }
    }
  } else {
    for (int j0 = 0; j0 < n; j0 += 4) {
      const int j = j0 + grp;
      const bool ok = j < n;
      float dot = 0.f;
      if (ok) {
        const int it = __ldg(a.rows.indx + beg + j);
        const float* vf = a.V + (size_t)it * k;
        for (int c = gl; c < k; c += 8) dot = fmaf(__ldg(uf + c), __ldg(vf + c), dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      if (ok && gl == 0) {
        const double pred = (double)dot + a.shift;
        const double diff = (double)__ldg(a.rows.vals + beg + j) - pred;
        sd2 += diff * diff;
        sp += pred;
      }
    }
  }
  sd2 = warp_sum(sd2);
  sp = warp_sum(sp);
  if (lane == 0) {
    a.row_sums[2 * (size_t)warp] = sd2;
    a.row_sums[2 * (size_t)warp + 1] = sp;
  }
}

// portion p covers rows [portion_first[p], portion_first[p+1])
__global__ void __launch_bounds__(256) rmse_portion_reduce_kernel(const double* __restrict__ row_sums,
                                                                  const int32_t* __restrict__ row_len,
                                                                  const int32_t* __restrict__ portion_first,
                                                                  double* __restrict__ portion_sums) {
  __shared__ double s0[256], s1[256], s2[256];
  const int p = blockIdx.x;
  const int lo = portion_first[p], hi = portion_first[p + 1];
  double d2 = 0.0, sp = 0.0, cnt = 0.0;
  for (int r = lo + threadIdx.x; r < hi; r += 256) {
    d2 += row_sums[2 * (size_t)r];
    sp += row_sums[2 * (size_t)r + 1];
    cnt += (double)row_len[r];
  }
  s0[threadIdx.x] = d2; s1[threadIdx.x] = sp; s2[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s0[threadIdx.x] += s0[threadIdx.x + o];
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s2[threadIdx.x] += s2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    portion_sums[3 * (size_t)p] = s0[0];      // rSumDiff2
    portion_sums[3 * (size_t)p + 1] = s2[0];  // rCnt
    portion_sums[3 * (size_t)p + 2] = s1[0];  // rSum
  }
}

// Two-level variant for ONE big portion (the per-portion path with multi-million-rating portions): a single
// CTA walking millions of rows is latency-bound (measured: ~5 ms for 1.7 M rows).  Level 1: one CTA per chunk
// of kRmseChunkRows rows -> chunk_sums[c][3]; level 2: rmse_portion_reduce_chunks_kernel adds the chunks.
// Both levels add in a fixed order, so the result does not depend on scheduling.
constexpr int kRmseChunkRows = 4096;

__global__ void __launch_bounds__(256) rmse_chunk_reduce_kernel(const double* __restrict__ row_sums,
                                                                const int32_t* __restrict__ row_len, int n_rows,
                                                                double* __restrict__ chunk_sums) {
  __shared__ double s0[256], s1[256], s2[256];
  const int lo = blockIdx.x * kRmseChunkRows, hi = min(n_rows, lo + kRmseChunkRows);
  double d2 = 0.0, sp = 0.0, cnt = 0.0;
  for (int r = lo + threadIdx.x; r < hi; r += 256) {
    d2 += row_sums[2 * (size_t)r];
    sp += row_sums[2 * (size_t)r + 1];
    cnt += (double)row_len[r];
  }
  s0[threadIdx.x] = d2; s1[threadIdx.x] = sp; s2[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s0[threadIdx.x] += s0[threadIdx.x + o];
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s2[threadIdx.x] += s2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    chunk_sums[3 * (size_t)blockIdx.x] = s0[0];
    chunk_sums[3 * (size_t)blockIdx.x + 1] = s2[0];
    chunk_sums[3 * (size_t)blockIdx.x + 2] = s1[0];
  }
}

__global__ void __launch_bounds__(256) rmse_portion_reduce_chunks_kernel(const double* __restrict__ chunk_sums,
                                                                         int n_chunks,
                                                                         double* __restrict__ portion_sums) {
  __shared__ double s[3][256];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int c = threadIdx.x; c < n_chunks; c += 256) {
    a0 += chunk_sums[3 * (size_t)c];
    a1 += chunk_sums[3 * (size_t)c + 1];
    a2 += chunk_sums[3 * (size_t)c + 2];
  }
  s[0][threadIdx.x] = a0; s[1][threadIdx.x] = a1; s[2][threadIdx.x] = a2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] += s[0][threadIdx.x + o];
      s[1][threadIdx.x] += s[1][threadIdx.x + o];
      s[2][threadIdx.x] += s[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    portion_sums[0] = s[0][0];   // rSumDiff2
    portion_sums[1] = s[1][0];   // rCnt
    portion_sums[2] = s[2][0];   // rSum
  }
}

}  // namespace ycnr
