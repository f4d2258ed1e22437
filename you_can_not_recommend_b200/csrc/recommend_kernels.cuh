// recommend_kernels.cuh — top-N item recommendation for a batch of users (sm_100a).
//
// Restates YcnrController.recommendItemsForUser (lib/YcnrController.js:227-284): for every item that is
// not in the user's skip list (rated + explicitly "unrated" items, 244-251)
//     predict = fp32 dot(U[u], V[i]) + globalAvgShift          (EmfBase.predictSync, EmfBase.js:815-827)
// items with predict >= minRecommendRating compete for the list, which is kept sorted by predict
// descending; upstream pops the last entry whenever the list reaches `limit` (281-282), so the
// result holds at most limit - 1 items — reproduced here (keep = limit - 1).
//
// Kernel 1 (recommend_scores_kernel): one warp per (user, item) pair group: fp32 dot products of the user's
//   row (registers) with every item row (L2-resident V), written as double predictions to scratch;
//   skipped / below-threshold items become -inf.
// Kernel 2 (recommend_select_kernel): one CTA per user, `keep` rounds of a block-wide arg-max
//   (ties: lower item id first — the order a stable sort gives upstream), selected entries are retired.
#pragma once
#include "common.cuh"

namespace ycnr {

struct RecommendArgs {
  const float* __restrict__ U;
  const float* __restrict__ V;
  int k;
  int n_items;
  int n_users;                            // users in this batch
  const int32_t* __restrict__ user_ids;   // [n_users] 0-based
  const int64_t* __restrict__ skip_ptr;   // [n_users + 1]
  const int32_t* __restrict__ skip_ids;   // 0-based item ids
  double shift;
  double min_rating;
  int keep;                               // limit - 1
  double* __restrict__ pred;              // scratch [n_users][n_items]
  int32_t* __restrict__ out_ids;          // [n_users][keep]
  double* __restrict__ out_pred;          // [n_users][keep]
  int32_t* __restrict__ out_count;        // [n_users]
};

constexpr int kRecItemsPerCta = 256;   // 8 warps x 32 items

// grid (ceil(items / 256), n_users), 256 threads: warp w scores items [base + 32 w, base + 32 w + 32)
__global__ void __launch_bounds__(256) recommend_scores_kernel(const RecommendArgs a) {
  const int u = a.user_ids[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = a.k;
  const float* uf = a.U + (size_t)u * k;
  const int i0 = blockIdx.x * kRecItemsPerCta + warp * 32;
  double* out = a.pred + (size_t)blockIdx.y * a.n_items;
  // lanes split the factor row; 4 items in flight per warp (8-lane groups)
  const int grp = lane >> 3, gl = lane & 7;
  for (int j0 = 0; j0 < 32; j0 += 4) {
    const int it = i0 + j0 + grp;
    float dot = 0.f;
    if (it < a.n_items) {
      const float* vf = a.V + (size_t)it * k;
      if ((k & 3) == 0) {
        for (int c = 4 * gl; c < k; c += 32) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(uf + c));
          const float4 y = __ldg(reinterpret_cast<const float4*>(vf + c));
          dot = fmaf(x.x, y.x, dot);
          dot = fmaf(x.y, y.y, dot);
          dot = fmaf(x.z, y.z, dot);
          dot = fmaf(x.w, y.w, dot);
        }
      } else {
        for (int c = gl; c < k; c += 8) dot = fmaf(__ldg(uf + c), __ldg(vf + c), dot);
      }
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    if (gl == 0 && it < a.n_items) {
      const double p = (double)dot + a.shift;                 // fp32 dot, then the shift in double (EmfBase.js:825-827)
      out[it] = p >= a.min_rating ? p : -INFINITY;            // YcnrController.js:271
    }
  }
}

// one CTA per user: retire the skip list, then `keep` rounds of arg-max
__global__ void __launch_bounds__(256) recommend_select_kernel(const RecommendArgs a) {
  __shared__ double sv[256];
  __shared__ int si[256];
  const int b = blockIdx.x;
  double* pred = a.pred + (size_t)b * a.n_items;
  for (int64_t e = a.skip_ptr[b] + threadIdx.x; e < a.skip_ptr[b + 1]; e += 256) {
    const int it = a.skip_ids[e];
    if (it >= 0 && it < a.n_items) pred[it] = -INFINITY;       // YcnrController.js:244-251, 268
  }
  __syncthreads();
  int count = 0;
  for (int r = 0; r < a.keep; ++r) {
    double best = -INFINITY;
    int bi = 0x7fffffff;
    for (int it = threadIdx.x; it < a.n_items; it += 256) {
      const double p = pred[it];
      if (p > best) { best = p; bi = it; }                     // ascending scan: ties keep the lower id
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const double q = sv[threadIdx.x + o];
        const int qi = si[threadIdx.x + o];
        if (q > sv[threadIdx.x] || (q == sv[threadIdx.x] && qi < si[threadIdx.x])) {
          sv[threadIdx.x] = q;
          si[threadIdx.x] = qi;
        }
      }
      __syncthreads();
    }
    const double top = sv[0];
    const int ti = si[0];
    __syncthreads();
    if (!(top > -INFINITY)) break;                             // uniform: nothing left above the threshold
    if (threadIdx.x == 0) {
      a.out_ids[(size_t)b * a.keep + r] = ti;
      a.out_pred[(size_t)b * a.keep + r] = top;
      pred[ti] = -INFINITY;
    }
    ++count;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.out_count[b] = count;
}

}  // namespace ycnr
