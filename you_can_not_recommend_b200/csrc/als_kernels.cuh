// als_kernels.cuh — FP32 (FFMA) explicit-ALS row update kernels for sm_100a.
//
// One CTA owns one row (user or item) of the solved side and performs the whole
// reference loop body (lib/emf/EmfWorker.js:214-248) without touching HBM in between:
//
//   gather     Y[c,:] = F[indx[c],:]            EmfBase.js:537-555      cp.async -> smem tiles
//   Gram       A = Y^T Y                        EmfWorker.js:231-232    4x4 register tiles, lower triangle
//   ridge      A += (lambda*n) I                EmfWorker.js:233-235
//   rhs        b = Y^T r                        EmfWorker.js:238-245    extra tile row of the same sweep
//   solve      x = A^-1 b                       EmfWorker.js:246        tile Cholesky in registers
//   write      S[rowId,:] = x                   EmfWorker.js:247        (+ NVLink peer replicas)
//
// Data layout: the k x k system lives in REGISTERS as 4x4 tiles of its lower triangle,
// tile (I,L), L<=I, owned by thread t = I(I+1)/2+L; an additional tile row I = mt holds
// the right-hand side, so the forward substitution happens inside the factorisation
// (Cholesky of the bordered matrix [[A b],[b^T .]]).  Shared memory only carries the
// gathered rows and a one-tile-column panel that is broadcast at every step.
//
// Rows with fewer ratings than factors use the dual system (als_dual_kernel):
//   x = Y^T (Y Y^T + lambda*n I)^-1 r   ==   (Y^T Y + lambda*n I)^-1 Y^T r
// an n x n solve instead of k x k — algebraically identical, ~(k/n)^3 cheaper.
// Long rows are cut into slices (MODE_PARTIAL) whose tile partials are summed in a
// fixed order by MODE_REDUCE, so results do not depend on scheduling.
#pragma once
#include "common.cuh"

#ifndef YCNR_DUAL_FFMA2
#define YCNR_DUAL_FFMA2 1
#endif

namespace ycnr {

enum { MODE_FUSED = 0, MODE_PARTIAL = 1, MODE_REDUCE = 2 };

constexpr int kStageRows = 32;  // ratings staged per smem tile in the primal kernels

// linear tile index -> (I, L).  Column-major enumeration of the bordered lower triangle: column L
// holds the tiles (L,L), (L+1,L), ..., (mt-1,L) and the right-hand-side tile (mt,L), i.e. mt-L+1
// consecutive indices.  A warp therefore owns (parts of) one or two tile columns: in step J of the
// factorisation the warps whose columns are <= J have nothing left to do and skip the phase as a
// whole, and the active warps are fully populated (the row-major order left every warp ~1/3 full).
__device__ __forceinline__ void tile_coords(int t, int mt, int /*ntri*/, int& I, int& L) {
  int l = 0, rem = t;
  while (rem >= mt - l + 1) {
    rem -= mt - l + 1;
    ++l;
  }
  L = l;
  I = l + rem;
}

__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // MUFU.RSQ, <= 2 ulp; no IEEE sqrt/div chains
  return y;
}

// Cholesky of a symmetric 4x4 tile held in full; on return d holds M = L^-1 (lower
// triangle, zeros above).  Only reciprocal square roots are needed: L's diagonal never
// appears on its own.  This is the serial link of every factorisation step, so it uses the
// MUFU approximation (relative error 2^-22, the same order as one fp32 rounding) instead of
// IEEE sqrtf + division (~4x the latency).
__device__ __forceinline__ void chol4_inverse(float (&d)[4][4]) {
  const float i0 = rsqrt_fast(d[0][0]);
  const float l10 = d[1][0] * i0, l20 = d[2][0] * i0, l30 = d[3][0] * i0;
  const float i1 = rsqrt_fast(fmaf(-l10, l10, d[1][1]));
  const float l21 = fmaf(-l20, l10, d[2][1]) * i1;
  const float l31 = fmaf(-l30, l10, d[3][1]) * i1;
  const float i2 = rsqrt_fast(fmaf(-l21, l21, fmaf(-l20, l20, d[2][2])));
  const float l32 = fmaf(-l31, l21, fmaf(-l30, l20, d[3][2])) * i2;
  const float i3 = rsqrt_fast(fmaf(-l32, l32, fmaf(-l31, l31, fmaf(-l30, l30, d[3][3]))));
  const float m10 = -(l10 * i0) * i1;
  const float m21 = -(l21 * i1) * i2;
  const float m32 = -(l32 * i2) * i3;
  const float m20 = -fmaf(l21, m10, l20 * i0) * i2;
  const float m31 = -fmaf(l32, m21, l31 * i1) * i3;
  const float m30 = -fmaf(l32, m20, fmaf(l31, m10, l30 * i0)) * i3;
  d[0][0] = i0;  d[0][1] = 0.f; d[0][2] = 0.f; d[0][3] = 0.f;
  d[1][0] = m10; d[1][1] = i1;  d[1][2] = 0.f; d[1][3] = 0.f;
  d[2][0] = m20; d[2][1] = m21; d[2][2] = i2;  d[2][3] = 0.f;
  d[3][0] = m30; d[3][1] = m31; d[3][2] = m32; d[3][3] = i3;
}

// Blocked right-looking Cholesky + both triangular solves on register tiles.
//   acc[q]  : tile (tI[q], tL[q]) of the bordered system; tI = tL = -1 marks "no tile".
//   panel   : (mt+1)*16 floats ([4][mt+1] float4), minv: mt*16 floats (inverse of every diagonal factor tile),
//             ysm: 4*mt floats of shared memory.
// On return ysm[0 .. 4*mt) holds the solution.  All threads of the CTA must call.
// Two barriers per tile column in the factorisation (the owner of the next diagonal tile
// factors it right after its own trailing update — look-ahead — so nobody waits on a
// dedicated "factor" phase) and one per tile row in the back substitution.  (A single-warp back
// substitution over factor tiles staged in shared memory was measured on B200: k x k solves 7.87 -> 7.65 ms,
// dual rows 20.5 -> 21.2 ms per MAL iteration — its serial latency hurts the small systems; not kept.)
template <int TPT>
__device__ __forceinline__ void tile_cholesky_solve(float (&acc)[TPT][4][4], const int (&tI)[TPT],
                                                    const int (&tL)[TPT], int mt, float* panel, float* minv,
                                                    float* ysm) {
  auto factor_and_publish = [&](float (&t)[4][4], float* dst) {
    chol4_inverse(t);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(t[i][0], t[i][1], t[i][2], t[i][3]);
  };
  const int ps = mt + 1;   // panel is laid out [i][I]: lanes owning consecutive tiles read consecutive 16-byte units
#pragma unroll
  for (int q = 0; q < TPT; ++q)
    if (tI[q] == 0 && tL[q] == 0) factor_and_publish(acc[q], minv);
  __syncthreads();
  for (int J = 0; J < mt; ++J) {
    const float* dcur = minv + 16 * J;
    // (a) panel: X_IJ = A_IJ * L_JJ^-T  (rows I > J, including the rhs row I = mt)
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
      if (tL[q] == J && tI[q] > J) {
        float m[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(dcur + 4 * i);
          m[i][0] = v.x; m[i][1] = v.y; m[i][2] = v.z; m[i][3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a0 = acc[q][i][0], a1 = acc[q][i][1], a2 = acc[q][i][2], a3 = acc[q][i][3];
          const float x0 = a0 * m[0][0];
          const float x1 = fmaf(a1, m[1][1], a0 * m[1][0]);
          const float x2 = fmaf(a2, m[2][2], fmaf(a1, m[2][1], a0 * m[2][0]));
          const float x3 = fmaf(a3, m[3][3], fmaf(a2, m[3][2], fmaf(a1, m[3][1], a0 * m[3][0])));
          acc[q][i][0] = x0; acc[q][i][1] = x1; acc[q][i][2] = x2; acc[q][i][3] = x3;
          *reinterpret_cast<float4*>(panel + 4 * (i * ps + tI[q])) = make_float4(x0, x1, x2, x3);
        }
      }
    }
    __syncthreads();
    // (b) trailing update: A_IL -= X_IJ * X_LJ^T for J < L <= I; the owner of the next
    //     diagonal tile then factors it and publishes its inverse for step J + 1
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
      if (tL[q] > J) {
        float xi[4][4], xl[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(panel + 4 * (i * ps + tI[q]));
          xi[i][0] = v.x; xi[i][1] = v.y; xi[i][2] = v.z; xi[i][3] = v.w;
          const float4 w = *reinterpret_cast<const float4*>(panel + 4 * (i * ps + tL[q]));
          xl[i][0] = w.x; xl[i][1] = w.y; xl[i][2] = w.z; xl[i][3] = w.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float s = acc[q][i][j];
#pragma unroll
            for (int c = 0; c < 4; ++c) s = fmaf(-xi[i][c], xl[j][c], s);
            acc[q][i][j] = s;
          }
        if (tL[q] == J + 1 && tI[q] == J + 1) factor_and_publish(acc[q], minv + 16 * (J + 1));
      }
    }
    __syncthreads();
    // panel is rewritten in (a) of the next step: all its readers passed the barrier above
  }
  // y = L^-1 b sits in row 0 of the rhs tiles; x_I = M_I^T y_I finishes block row I
  auto apply_minv_t = [&](int I, float4 y) {
    const float* m = minv + 16 * I;
    const float4 r0 = *reinterpret_cast<const float4*>(m), r1 = *reinterpret_cast<const float4*>(m + 4);
    const float4 r2 = *reinterpret_cast<const float4*>(m + 8), r3 = *reinterpret_cast<const float4*>(m + 12);
    float4 x;
    x.x = fmaf(r3.x, y.w, fmaf(r2.x, y.z, fmaf(r1.x, y.y, r0.x * y.x)));
    x.y = fmaf(r3.y, y.w, fmaf(r2.y, y.z, r1.y * y.y));
    x.z = fmaf(r3.z, y.w, r2.z * y.z);
    x.w = r3.w * y.w;
    return x;
  };
#pragma unroll
  for (int q = 0; q < TPT; ++q) {
    if (tI[q] == mt && tL[q] >= 0) {
      float4 y = make_float4(acc[q][0][0], acc[q][0][1], acc[q][0][2], acc[q][0][3]);
      if (tL[q] == mt - 1) y = apply_minv_t(mt - 1, y);
      *reinterpret_cast<float4*>(ysm + 4 * tL[q]) = y;
    }
  }
  __syncthreads();
  // back substitution L^T x = y, right-looking over tile rows: in round I the tiles (I, L<I)
  // push x_I into y_L; the tile (I, I-1) makes the last contribution to y_{I-1} and turns it into x_{I-1}
  for (int I = mt - 1; I > 0; --I) {
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
      if (tI[q] == I && tL[q] >= 0 && tL[q] < I) {  // y_L -= X_IL^T x_I
        const float4 x = *reinterpret_cast<const float4*>(ysm + 4 * I);
        float4 y = *reinterpret_cast<const float4*>(ysm + 4 * tL[q]);
        y.x -= acc[q][0][0] * x.x + acc[q][1][0] * x.y + acc[q][2][0] * x.z + acc[q][3][0] * x.w;
        y.y -= acc[q][0][1] * x.x + acc[q][1][1] * x.y + acc[q][2][1] * x.z + acc[q][3][1] * x.w;
        y.z -= acc[q][0][2] * x.x + acc[q][1][2] * x.y + acc[q][2][2] * x.z + acc[q][3][2] * x.w;
        y.w -= acc[q][0][3] * x.x + acc[q][1][3] * x.y + acc[q][2][3] * x.z + acc[q][3][3] * x.w;
        if (tL[q] == I - 1) y = apply_minv_t(I - 1, y);
        *reinterpret_cast<float4*>(ysm + 4 * tL[q]) = y;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// Primal kernels: k x k system, KT = tiles per dimension (4*KT >= k), NT threads,
// TPT tiles per thread.
// ------------------------------------------------------------------------------------
struct PrimalArgs {
  RowsView rows;
  const float* __restrict__ fixed;  // F, row-major, k columns
  size_t fixed_bytes;               // size of F (decides whether the Gram loaders prefetch into L2)
  int k;
  double lambda;
  DstList dst;                      // S replicas
  const int32_t* __restrict__ work; // MODE_FUSED/REDUCE: row index per CTA
  // slices of long rows
  const int32_t* __restrict__ item_row;   // MODE_PARTIAL: row index per work item
  const int32_t* __restrict__ item_off;   // MODE_PARTIAL: first rating of the slice inside the row
  const int32_t* __restrict__ item_order; // tensor-core Gram: processing order of the work items (longest first)
  int n_items_total;
  int split_cols;                         // slice length
  float* __restrict__ partial;            // [items][tiles][16]
  const int32_t* __restrict__ row_first_item;  // MODE_REDUCE: per work entry
  const int32_t* __restrict__ row_n_items;
  const void* tmap;                            // host pointer to the CUtensorMap of `fixed` (tensor-core Gram), or null
  int fixed_rows;
};

// Occupancy hint of the k x k solve (MODE_REDUCE) with three tiles per thread (128 threads at k = 100): at least
// YCNR_REDUCE_MIN_CTAS CTAs per SM (5 -> at most 102 registers): 8.5 ms per MAL iteration against 9.0 ms for two
// tiles per thread (192 threads) at the compiler's own 80 registers.  Other shapes return 0 = unspecified.
// NB: an explicit 1 lets ptxas spend registers freely and costs occupancy (8.9 -> 12.5 ms).
#ifndef YCNR_REDUCE_MIN_CTAS
#define YCNR_REDUCE_MIN_CTAS 5
#endif
constexpr int primal_min_ctas(int kt, int /*nt*/, int tpt, int mode) {
  return (mode == 2 && tpt == 3 && kt <= 25) ? YCNR_REDUCE_MIN_CTAS : 0;
}
template <int KT, int NT, int TPT, int MODE>
__global__ void __launch_bounds__(NT, primal_min_ctas(KT, NT, TPT, MODE)) als_primal_kernel(const PrimalArgs a) {
  constexpr int KP = 4 * KT;          // padded system size
  constexpr int PITCH = KP + 4;       // + (val, 0, 0, 0): the rhs tile row reads its "a" operand here
  constexpr int NTRI = KT * (KT + 1) / 2;
  constexpr int NTILES = NTRI + KT;
  static_assert(NT * TPT >= NTILES, "not enough threads for the tile set");
  __shared__ __align__(16) float Ys[(MODE == MODE_REDUCE) ? 1 : 2 * kStageRows * PITCH];
  __shared__ __align__(16) float panel[(KT + 1) * 16];
  __shared__ __align__(16) float minv[KT * 16];
  __shared__ __align__(16) float ysm[KP];

  if (rows_poisoned(a.rows)) return;
  const int tid = threadIdx.x;
  const int k = a.k;
  int tI[TPT], tL[TPT];
#pragma unroll
  for (int q = 0; q < TPT; ++q) {
    const int t = tid + q * NT;
    if (t < NTILES) tile_coords(t, KT, NTRI, tI[q], tL[q]);
    else { tI[q] = -1; tL[q] = -1; }
  }
  float acc[TPT][4][4];
#pragma unroll
  for (int q = 0; q < TPT; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][i][j] = 0.f;

  int row;          // index into the row list
  int64_t seg_beg;  // first rating of this CTA's slice
  int seg_len;
  if (MODE == MODE_PARTIAL) {
    row = a.item_row[blockIdx.x];
    const int off = a.item_off[blockIdx.x];
    seg_beg = a.rows.row_start[row] + off;
    seg_len = min(a.split_cols, a.rows.row_len[row] - off);
  } else {
    row = a.work[blockIdx.x];
    seg_beg = a.rows.row_start[row];
    seg_len = a.rows.row_len[row];
  }

  if (MODE != MODE_REDUCE) {
    // ---- gather + Gram + rhs ------------------------------------------------------
    const bool vec = (k & 3) == 0;
    const int CH = k >> 2;
    // pad columns stay zero for the whole kernel (never targeted by cp.async)
    for (int q = tid; q < 2 * kStageRows * (PITCH - k); q += NT) {
      const int r = q / (PITCH - k), c = k + q % (PITCH - k);
      if (c != KP) Ys[r * PITCH + c] = 0.f;
    }
    auto stage = [&](int t0, int buf) {
      float* base = Ys + buf * kStageRows * PITCH;
      if (vec) {
        for (int q = tid; q < kStageRows * CH; q += NT) {
          const int r = q / CH, c = q - r * CH;
          const bool ok = t0 + r < seg_len;
          const int col = ok ? __ldg(a.rows.indx + seg_beg + t0 + r) : 0;
          cp_async16(base + r * PITCH + 4 * c, a.fixed + (size_t)col * k + 4 * c, ok ? 16 : 0);
        }
      } else {
        for (int q = tid; q < kStageRows * k; q += NT) {
          const int r = q / k, c = q - r * k;
          const bool ok = t0 + r < seg_len;
          const int col = ok ? __ldg(a.rows.indx + seg_beg + t0 + r) : 0;
          cp_async4(base + r * PITCH + c, a.fixed + (size_t)col * k + c, ok ? 4 : 0);
        }
      }
      for (int r = tid; r < kStageRows; r += NT) {
        const bool ok = t0 + r < seg_len;
        cp_async4(base + r * PITCH + KP, a.rows.vals + seg_beg + (ok ? t0 + r : 0), ok ? 4 : 0);
      }
    };
    int aoff[TPT], boff[TPT];
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
      aoff[q] = tI[q] < 0 ? 0 : 4 * tI[q];  // rhs row (tI == KT) reads (val,0,0,0) at column KP
      boff[q] = tL[q] < 0 ? 0 : 4 * tL[q];
    }
    const int ntile = (seg_len + kStageRows - 1) / kStageRows;
    if (ntile > 0) stage(0, 0);
    cp_async_commit();
    for (int t = 0; t < ntile; ++t) {
      if (t + 1 < ntile) stage((t + 1) * kStageRows, (t + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const float* yb = Ys + (t & 1) * kStageRows * PITCH;
      const int rcount = min(kStageRows, seg_len - t * kStageRows);
#pragma unroll 4
      for (int r = 0; r < rcount; ++r) {
        const float* yr = yb + r * PITCH;
#pragma unroll
        for (int q = 0; q < TPT; ++q) {
          const float4 av = *reinterpret_cast<const float4*>(yr + aoff[q]);
          const float4 bv = *reinterpret_cast<const float4*>(yr + boff[q]);
          const float ar[4] = {av.x, av.y, av.z, av.w};
          const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[q][i][j] = fmaf(ar[i], br[j], acc[q][i][j]);
        }
      }
      __syncthreads();
    }
  }

  if (MODE == MODE_PARTIAL) {
    float* out = a.partial + (size_t)blockIdx.x * NTILES * 16;
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
      const int t = tid + q * NT;
      if (t < NTILES) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(out + (size_t)t * 16 + 4 * i) =
              make_float4(acc[q][i][0], acc[q][i][1], acc[q][i][2], acc[q][i][3]);
      }
    }
    return;
  }

  if (MODE == MODE_REDUCE) {
    const int first = a.row_first_item[blockIdx.x];
    const int nit = a.row_n_items[blockIdx.x];
    for (int it = 0; it < nit; ++it) {  // fixed order => deterministic sums
      const float* in = a.partial + (size_t)(first + it) * NTILES * 16;
#pragma unroll
      for (int q = 0; q < TPT; ++q) {
        const int t = tid + q * NT;
        if (t < NTILES) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(in + (size_t)t * 16 + 4 * i);
            acc[q][i][0] += v.x; acc[q][i][1] += v.y; acc[q][i][2] += v.z; acc[q][i][3] += v.w;
          }
        }
      }
    }
  }

  // ---- ridge (EmfWorker.js:233-235: lambda*n formed in double, stored as float) -------
  const float lam = (float)(a.lambda * (double)a.rows.row_len[row]);
#pragma unroll
  for (int q = 0; q < TPT; ++q) {
    if (tI[q] >= 0 && tI[q] == tL[q]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (4 * tI[q] + i < k) acc[q][i][i] += lam;
        else acc[q][i][i] = 1.0f;  // padding rows/cols: identity block, solution 0
      }
    }
  }
  tile_cholesky_solve<TPT>(acc, tI, tL, KT, panel, minv, ysm);

  const int rowId = a.rows.row_ids[row];
  for (int c = tid; c < k; c += NT) {
    const float x = ysm[c];
    for (int d = 0; d < a.dst.n; ++d) a.dst.p[d][(size_t)rowId * k + c] = x;
  }
}

// ------------------------------------------------------------------------------------
// Solve of systems wider than one tensor-core pass (k > 124).  The Gram kernel covered the k x k system by one
// pass per PAIR (a < b) of column blocks (wt tile rows each); pass p left, per slice, the tile partials of the
// 2 wt x 2 wt "virtual" system [block a | block b] in the usual column-major tile enumeration.  This kernel
// assembles every tile (I, L) of the full system from the pass that holds it, sums the row's slices in a fixed
// order, adds the ridge and runs the same register-tile Cholesky.
// ------------------------------------------------------------------------------------
struct SolveBlocksArgs {
  RowsView rows;
  int k;
  double lambda;
  DstList dst;
  const int32_t* __restrict__ work;            // row index per CTA
  const int32_t* __restrict__ row_first_item;  // first slice of the row (index into the step's work items)
  const int32_t* __restrict__ row_n_items;
  const float* __restrict__ partial;           // [pass][item - item_base][virtual tiles][16]
  size_t pass_stride;                          // floats between passes
  int item_base;
  int nb, wt;                                  // column blocks, tile rows per block
};

// linear index of tile (Iv, Lv) (Iv == ktv: rhs row) in the column-major enumeration of a ktv-row system
__device__ __forceinline__ int tile_linear(int Iv, int Lv, int ktv) {
  return Lv * (ktv + 1) - (Lv * (Lv - 1)) / 2 + (Iv - Lv);
}

template <int KT, int NT, int TPT>
__global__ void __launch_bounds__(NT) als_solve_blocks_kernel(const SolveBlocksArgs a) {
  constexpr int KP = 4 * KT;
  constexpr int NTRI = KT * (KT + 1) / 2;
  constexpr int NTILES = NTRI + KT;
  static_assert(NT * TPT >= NTILES, "not enough threads for the tile set");
  if (rows_poisoned(a.rows)) return;
  __shared__ __align__(16) float panel[(KT + 1) * 16];
  __shared__ __align__(16) float minv[KT * 16];
  __shared__ __align__(16) float ysm[KP];
  const int tid = threadIdx.x;
  const int k = a.k, nb = a.nb, wt = a.wt;
  const int ktv = 2 * wt, ntv = ktv * (ktv + 1) / 2 + ktv;
  const int row = a.work[blockIdx.x];
  const int first = a.row_first_item[blockIdx.x] - a.item_base;
  const int nit = a.row_n_items[blockIdx.x];
  const float lam = (float)(a.lambda * (double)a.rows.row_len[row]);

  int tI[TPT], tL[TPT];
  float acc[TPT][4][4];
#pragma unroll
  for (int q = 0; q < TPT; ++q) {
    const int t = tid + q * NT;
    if (t < NTILES) tile_coords(t, KT, NTRI, tI[q], tL[q]);
    else { tI[q] = -1; tL[q] = -1; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][i][j] = 0.f;
    if (tI[q] < 0) continue;
    const int I = tI[q], L = tL[q];
    const int sb = L / wt;
    const int sa = I == KT ? sb : I / wt;
    if (sb >= nb || sa >= nb) continue;                     // padding tiles beyond the last block: zero
    int pa, pb;                                             // the pass (pa < pb) that holds the tile
    if (sa != sb) { pa = sb; pb = sa; }
    else if (sb + 1 < nb) { pa = sb; pb = sb + 1; }
    else { pa = sb - 1; pb = sb; }
    const int pass = pa * nb - (pa * (pa + 1)) / 2 + (pb - pa - 1);
    const int Lv = L - sb * wt + (sb == pb ? wt : 0);
    const int Iv = I == KT ? ktv : I - sa * wt + (sa == pb ? wt : 0);
    const float* in = a.partial + (size_t)pass * a.pass_stride + ((size_t)first * ntv + tile_linear(Iv, Lv, ktv)) * 16;
    for (int it = 0; it < nit; ++it, in += (size_t)ntv * 16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(in + 4 * i);
        acc[q][i][0] += v.x; acc[q][i][1] += v.y; acc[q][i][2] += v.z; acc[q][i][3] += v.w;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < TPT; ++q) {
    if (tI[q] >= 0 && tI[q] == tL[q]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (4 * tI[q] + i < k) acc[q][i][i] += lam;
        else acc[q][i][i] = 1.0f;
      }
    }
  }
  tile_cholesky_solve<TPT>(acc, tI, tL, KT, panel, minv, ysm);
  const int rowId = a.rows.row_ids[row];
  for (int c = tid; c < k; c += NT) {
    const float x = ysm[c];
    for (int d = 0; d < a.dst.n; ++d) a.dst.p[d][(size_t)rowId * k + c] = x;
  }
}

// ------------------------------------------------------------------------------------
// Dual kernel: rows with n <= 4*MT_MAX ratings, n x n system  G = Y Y^T + lambda*n I.
// ------------------------------------------------------------------------------------
struct DualArgs {
  RowsView rows;
  const float* __restrict__ fixed;
  int k;
  int pitch;  // floats per staged row, multiple of 4 with (pitch/4) odd => conflict-free LDS.128
  double lambda;
  DstList dst;
  const int32_t* __restrict__ work;
  int bar_off;  // als_dual_kernel: float offset of the gather mbarrier inside its dynamic shared memory
};

constexpr int kDualMaxSplit = 8;   // largest K-split of the Gram sweep (k = 100: 25 chunks of 4 columns)

// tiles of the bordered lower triangle of an mt x mt tile system (Gram tiles + rhs tile row)
__host__ __device__ constexpr int dual_ntl(int mt) { return mt * (mt + 1) / 2 + mt; }
// floats of K-split scratch a CTA of NT threads needs for systems of exactly mt tile rows
__host__ __device__ constexpr int dual_red_floats(int mt, int nt) {
  int g = nt / dual_ntl(mt);
  if (g > kDualMaxSplit) g = kDualMaxSplit;
  return g > 1 ? (g - 1) * dual_ntl(mt) * 16 : 0;
}

template <int MT_MAX, int NT>
__global__ void __launch_bounds__(NT) als_dual_kernel(const DualArgs a) {
  constexpr int NMAX = 4 * MT_MAX;
  static_assert(NT >= MT_MAX * (MT_MAX + 1) / 2 + MT_MAX, "not enough threads for the tile set");
  extern __shared__ __align__(16) float dsm[];
  float* Y = dsm;                              // [NMAX][pitch]
  float* panel = Y + NMAX * a.pitch;           // (MT_MAX+1)*16
  float* minv = panel + (MT_MAX + 1) * 16;     // MT_MAX x 16
  float* ysm = minv + MT_MAX * 16;             // NMAX
  float* vs = ysm + NMAX;                      // NMAX
  float* red = vs + NMAX;                      // K-split partial tiles: [G-1][4][ntl] float4

  if (rows_poisoned(a.rows)) return;
  const int tid = threadIdx.x;
  const int k = a.k, pitch = a.pitch;
  const int K4 = (k + 3) & ~3;
  const int row = a.work[blockIdx.x];
  const int64_t beg = a.rows.row_start[row];
  const int n = a.rows.row_len[row];
  const int mt = (n + 3) >> 2;
  const int np = 4 * mt;

  // ---- gather the n rated rows of F -----------------------------------------------
  // Rating r (= row 4I+i of the system) is staged in slot i*mt + I: the four rows of a tile
  // sit mt slots apart, so lanes owning consecutive tiles read consecutive slots and the
  // LDS.128 of the Gram loop are bank-conflict free (pitch/4 is odd).
  const bool bulk = (k & 3) == 0;
  const uint32_t bar = smem_u32(dsm + a.bar_off);   // 8 bytes behind every other array
  if (bulk) {   // TMA bulk copies, one per rated row (see als_dual_tpt_kernel)
    if (tid == 0) {
      tma_mbar_init(bar, 1);
      tma_mbar_expect_tx(bar, (uint32_t)n * (uint32_t)k * 4u);
    }
    const int CH = k >> 2;
    for (int q = tid; q < (np - n) * CH; q += NT) {
      const int r = n + q / CH, c = q % CH;
      *reinterpret_cast<float4*>(Y + ((r & 3) * mt + (r >> 2)) * pitch + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int r = tid; r < n; r += NT) {
      const int col = __ldg(a.rows.indx + beg + r);
      tma_bulk_g2s(Y + ((r & 3) * mt + (r >> 2)) * pitch, a.fixed + (size_t)col * k, (uint32_t)k * 4u, bar);
    }
  } else {
    for (int q = tid; q < np * K4; q += NT) {
      const int r = q / K4, c = q - r * K4;
      const bool ok = r < n && c < k;
      const int col = ok ? __ldg(a.rows.indx + beg + r) : 0;
      cp_async4(Y + ((r & 3) * mt + (r >> 2)) * pitch + c, a.fixed + (size_t)col * k + (ok ? c : 0), ok ? 4 : 0);
    }
  }
  for (int r = tid; r < np; r += NT) vs[r] = r < n ? __ldg(a.rows.vals + beg + r) : 0.f;
  cp_async_commit();
  cp_async_wait<0>();
  if (bulk) tma_mbar_wait(bar, 0);
  __syncthreads();

  // ---- G tiles ------------------------------------------------------------------------
  // Thread t < ntl owns tile t for the factorisation.  The CTA has NT >= G*ntl threads: the spare ones
  // take a share of the Gram sweep (K-split, thread g*ntl + t sums columns [c_beg, c_end) of tile t) and
  // hand their partial tile over through shared memory; the owner adds them in group order, so the result
  // does not depend on scheduling.
  const int ntri = mt * (mt + 1) / 2;
  const int ntl = ntri + mt;
  int G = NT / ntl;
  if (G > kDualMaxSplit) G = kDualMaxSplit;
  const int g = tid / ntl, t = tid - g * ntl;
  int tI[1], tL[1];
  if (g < G) tile_coords(t, mt, ntri, tI[0], tL[0]);
  else { tI[0] = -1; tL[0] = -1; }
  float acc[1][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][i][j] = 0.f;

  const bool gram_tile = tI[0] >= 0 && tI[0] < mt;
  if (gram_tile) {
    const float* ya = Y + tI[0] * pitch;
    const float* yb = Y + tL[0] * pitch;
    const int tstride = mt * pitch;
    const int CHT = K4 >> 2;
    const int c_beg = 4 * ((g * CHT) / G), c_end = 4 * (((g + 1) * CHT) / G);
#if YCNR_DUAL_FFMA2
    // packed FP32 FMA (fma.rn.f32x2, sm_100): the two lanes of a pair accumulate the even and the odd
    // column pairs of the chunk, so both operands are register pairs exactly as LDS.128 delivered them;
    // half the FMA issue slots of the scalar loop (ncu: these kernels are issue-bound, 65-72 % issue active
    // with FFMA only 1/3-1/2 of the instructions)
    float2 acc2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    for (int c = c_beg; c < c_end; c += 4) {
      float4 av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = *reinterpret_cast<const float4*>(ya + i * tstride + c);
        bv[i] = *reinterpret_cast<const float4*>(yb + i * tstride + c);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 s = acc2[i][j];
          s = __ffma2_rn(make_float2(av[i].x, av[i].y), make_float2(bv[j].x, bv[j].y), s);
          s = __ffma2_rn(make_float2(av[i].z, av[i].w), make_float2(bv[j].z, bv[j].w), s);
          acc2[i][j] = s;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[0][i][j] = acc2[i][j].x + acc2[i][j].y;
#else
    for (int c = c_beg; c < c_end; c += 4) {
      float4 av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = *reinterpret_cast<const float4*>(ya + i * tstride + c);
        bv[i] = *reinterpret_cast<const float4*>(yb + i * tstride + c);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s = acc[0][i][j];
          s = fmaf(av[i].x, bv[j].x, s);
          s = fmaf(av[i].y, bv[j].y, s);
          s = fmaf(av[i].z, bv[j].z, s);
          s = fmaf(av[i].w, bv[j].w, s);
          acc[0][i][j] = s;
        }
    }
#endif
  }
  if (G > 1) {   // uniform over the CTA
    if (g > 0) {
      if (gram_tile) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(red + 4 * (((g - 1) * 4 + i) * ntl + t)) =
              make_float4(acc[0][i][0], acc[0][i][1], acc[0][i][2], acc[0][i][3]);
      }
      tI[0] = -1;   // helpers own no tile of the factorisation
      tL[0] = -1;
    }
    __syncthreads();
    if (g == 0 && gram_tile) {
      for (int gg = 1; gg < G; ++gg) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(red + 4 * (((gg - 1) * 4 + i) * ntl + t));
          acc[0][i][0] += v.x; acc[0][i][1] += v.y; acc[0][i][2] += v.z; acc[0][i][3] += v.w;
        }
      }
    }
  }
  if (tI[0] >= 0 && tI[0] < mt) {
    if (tI[0] == tL[0]) {
      const float lam = (float)(a.lambda * (double)n);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (4 * tI[0] + i < n) acc[0][i][i] += lam;
        else acc[0][i][i] = 1.0f;
      }
    }
  } else if (tI[0] == mt) {
    const float4 v = *reinterpret_cast<const float4*>(vs + 4 * tL[0]);
    acc[0][0][0] = v.x; acc[0][0][1] = v.y; acc[0][0][2] = v.z; acc[0][0][3] = v.w;
  }
  tile_cholesky_solve<1>(acc, tI, tL, mt, panel, minv, ysm);

  // ---- x = Y^T z --------------------------------------------------------------------------
  const int rowId = a.rows.row_ids[row];
  for (int p = tid; p < np; p += NT) vs[(p & 3) * mt + (p >> 2)] = ysm[p];   // z in slot order (pad rows: z = 0)
  __syncthreads();
  for (int c = tid; c < k; c += NT) {
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    const float* yc = Y + c;
    for (int s = 0; s < np; s += 4) {   // four independent chains; np is a multiple of 4
      x0 = fmaf(yc[(s + 0) * pitch], vs[s + 0], x0);
      x1 = fmaf(yc[(s + 1) * pitch], vs[s + 1], x1);
      x2 = fmaf(yc[(s + 2) * pitch], vs[s + 2], x2);
      x3 = fmaf(yc[(s + 3) * pitch], vs[s + 3], x3);
    }
    const float x = (x0 + x1) + (x2 + x3);
    for (int d = 0; d < a.dst.n; ++d) a.dst.p[d][(size_t)rowId * k + c] = x;
  }
}

// ------------------------------------------------------------------------------------
// Dual kernel, several tiles per thread: the same n x n system, but the CTA is only as wide as
// ceil(ntl / TPT) threads (rounded to warps).  Every step of the factorisation costs each WARP of the CTA a
// fixed number of control instructions and two barriers whether or not its lanes hold live tiles, so the
// thread-per-tile kernel above pays that overhead 11 times per step at mt = 24; with three or four tiles per
// thread it is paid 3 times and the CTAs are small enough for more systems to be resident per SM.
// MT is exact (the host bins rows by tile-row count).
// ------------------------------------------------------------------------------------
// (a register cap through __launch_bounds__ min-CTAs was measured for the multi-warp bins — 112 / 100 / 88 registers
//  instead of 129: dual rows 15.9 -> 17.4 / 17.4 / 19.7 ms per MAL iteration; the tiles spill — not kept)
template <int MT, int NT, int TPT>
__global__ void __launch_bounds__(NT) als_dual_tpt_kernel(const DualArgs a) {
  constexpr int NMAX = 4 * MT;
  constexpr int NTRI = MT * (MT + 1) / 2;
  constexpr int NTL = NTRI + MT;
  static_assert(NT * TPT >= NTL, "not enough threads for the tile set");
  if (rows_poisoned(a.rows)) return;
  extern __shared__ __align__(16) float dsm[];
  float* Y = dsm;                          // [NMAX][pitch]
  float* panel = Y + NMAX * a.pitch;       // (MT+1)*16
  float* minv = panel + (MT + 1) * 16;     // MT x 16
  float* ysm = minv + MT * 16;             // NMAX
  float* vs = ysm + NMAX;                  // NMAX

  const int tid = threadIdx.x;
  const int k = a.k, pitch = a.pitch;
  const int K4 = (k + 3) & ~3;
  const int row = a.work[blockIdx.x];
  const int64_t beg = a.rows.row_start[row];
  const int n = a.rows.row_len[row];
  constexpr int mt = MT;
  constexpr int np = 4 * MT;

  const bool bulk = (k & 3) == 0;
  const uint32_t bar = smem_u32(vs + NMAX);       // 8 bytes behind the scratch vectors (16-byte aligned)
  if (bulk) {
    // TMA bulk copies: one cp.async.bulk per rated row (k * 4 bytes, 16-byte aligned on both sides) instead of
    // k / 4 cp.async instructions; their bytes complete the mbarrier
    if (tid == 0) {
      tma_mbar_init(bar, 1);
      tma_mbar_expect_tx(bar, (uint32_t)n * (uint32_t)k * 4u);
    }
    const int CH = k >> 2;
    for (int q = tid; q < (np - n) * CH; q += NT) {            // padding rows of the last tile: zero
      const int r = n + q / CH, c = q % CH;
      *reinterpret_cast<float4*>(Y + ((r & 3) * mt + (r >> 2)) * pitch + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();                                          // barrier initialised and armed
    for (int r = tid; r < n; r += NT) {
      const int col = __ldg(a.rows.indx + beg + r);
      tma_bulk_g2s(Y + ((r & 3) * mt + (r >> 2)) * pitch, a.fixed + (size_t)col * k, (uint32_t)k * 4u, bar);
    }
  } else {
    for (int q = tid; q < np * K4; q += NT) {
      const int r = q / K4, c = q - r * K4;
      const bool ok = r < n && c < k;
      const int col = ok ? __ldg(a.rows.indx + beg + r) : 0;
      cp_async4(Y + ((r & 3) * mt + (r >> 2)) * pitch + c, a.fixed + (size_t)col * k + (ok ? c : 0), ok ? 4 : 0);
    }
  }
  for (int r = tid; r < np; r += NT) vs[r] = r < n ? __ldg(a.rows.vals + beg + r) : 0.f;
  cp_async_commit();

  int tI[TPT], tL[TPT];
#pragma unroll
  for (int q = 0; q < TPT; ++q) {
    const int t = tid + q * NT;
    if (t < NTL) tile_coords(t, mt, NTRI, tI[q], tL[q]);
    else { tI[q] = -1; tL[q] = -1; }
  }
  float acc[TPT][4][4];
  cp_async_wait<0>();
  if (bulk) tma_mbar_wait(bar, 0);
  __syncthreads();

  const float lam = (float)(a.lambda * (double)n);
  const int tstride = mt * pitch;
  // One-warp systems with at most 16 Gram tiles (mt <= 5): the sweep would keep half of the lanes idle (the rhs
  // tile owners and the lanes beyond the tile set), so it is K-split inside the warp instead — lane g * NTRI + r
  // sums share g of the column chunks of Gram tile r (row-major rank), the owner lanes collect the GS partial
  // tiles by shuffle in group order (deterministic) and go on as below.
  constexpr int GS = (TPT == 1 && NT == 32 && 2 * NTRI <= 32) ? 32 / NTRI : 1;
  if constexpr (GS > 1) {
    const int g = tid / NTRI, r = tid - g * NTRI;
    int sI = 0, rem = r;
    while (rem > sI) { rem -= sI + 1; ++sI; }        // r = sI (sI + 1) / 2 + sL
    const int sL = rem;
    float2 acc2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    if (g < GS) {
      const float* ya = Y + sI * pitch;
      const float* yb = Y + sL * pitch;
      const int CHT = K4 >> 2;
      const int c_beg = 4 * ((g * CHT) / GS), c_end = 4 * (((g + 1) * CHT) / GS);
      for (int c = c_beg; c < c_end; c += 4) {
        float4 av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          av[i] = *reinterpret_cast<const float4*>(ya + i * tstride + c);
          bv[i] = *reinterpret_cast<const float4*>(yb + i * tstride + c);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 s2 = acc2[i][j];
            s2 = __ffma2_rn(make_float2(av[i].x, av[i].y), make_float2(bv[j].x, bv[j].y), s2);
            s2 = __ffma2_rn(make_float2(av[i].z, av[i].w), make_float2(bv[j].z, bv[j].w), s2);
            acc2[i][j] = s2;
          }
      }
    }
    // owner lane of Gram tile (I, L) reads the partial tiles of lanes rank, NTRI + rank, ...
    const bool own = tI[0] >= 0 && tI[0] < mt;
    const int rank = own ? tI[0] * (tI[0] + 1) / 2 + tL[0] : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float part = acc2[i][j].x + acc2[i][j].y;
        float sum = 0.f;
#pragma unroll
        for (int gg = 0; gg < GS; ++gg) sum += __shfl_sync(0xffffffffu, part, rank + gg * NTRI);
        acc[0][i][j] = own ? sum : 0.f;
      }
    if (own && tI[0] == tL[0]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (4 * tI[0] + i < n) acc[0][i][i] += lam;
        else acc[0][i][i] = 1.0f;
      }
    } else if (tI[0] == mt) {
      const float4 v = *reinterpret_cast<const float4*>(vs + 4 * tL[0]);
      acc[0][0][0] = v.x; acc[0][0][1] = v.y; acc[0][0][2] = v.z; acc[0][0][3] = v.w;
    }
  } else
#pragma unroll
  for (int q = 0; q < TPT; ++q) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][i][j] = 0.f;
    if (tI[q] >= 0 && tI[q] < mt) {
      const float* ya = Y + tI[q] * pitch;
      const float* yb = Y + tL[q] * pitch;
      float2 acc2[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
#pragma unroll 2
      for (int c = 0; c < K4; c += 4) {
        float4 av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          av[i] = *reinterpret_cast<const float4*>(ya + i * tstride + c);
          bv[i] = *reinterpret_cast<const float4*>(yb + i * tstride + c);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 s = acc2[i][j];
            s = __ffma2_rn(make_float2(av[i].x, av[i].y), make_float2(bv[j].x, bv[j].y), s);
            s = __ffma2_rn(make_float2(av[i].z, av[i].w), make_float2(bv[j].z, bv[j].w), s);
            acc2[i][j] = s;
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[q][i][j] = acc2[i][j].x + acc2[i][j].y;
      if (tI[q] == tL[q]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (4 * tI[q] + i < n) acc[q][i][i] += lam;
          else acc[q][i][i] = 1.0f;
        }
      }
    } else if (tI[q] == mt) {
      const float4 v = *reinterpret_cast<const float4*>(vs + 4 * tL[q]);
      acc[q][0][0] = v.x; acc[q][0][1] = v.y; acc[q][0][2] = v.z; acc[q][0][3] = v.w;
    }
  }
  __syncthreads();   // vs is rewritten below only after the solve; Y stays
  tile_cholesky_solve<TPT>(acc, tI, tL, mt, panel, minv, ysm);

  const int rowId = a.rows.row_ids[row];
  for (int p = tid; p < np; p += NT) vs[(p & 3) * mt + (p >> 2)] = ysm[p];   // z in slot order (pad rows: z = 0)
  __syncthreads();
  if ((k & 3) == 0) {
    // one float4 of the solution per thread: np LDS.128 + broadcast LDS of z, 4 FMA each
    for (int c4 = tid; c4 < (k >> 2); c4 += NT) {
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      const float* yc = Y + 4 * c4;
#pragma unroll 2
      for (int s2 = 0; s2 < np; s2 += 2) {
        const float4 y0 = *reinterpret_cast<const float4*>(yc + (s2 + 0) * pitch);
        const float4 y1 = *reinterpret_cast<const float4*>(yc + (s2 + 1) * pitch);
        const float z0 = vs[s2], z1 = vs[s2 + 1];
        x0.x = fmaf(y0.x, z0, x0.x); x0.y = fmaf(y0.y, z0, x0.y); x0.z = fmaf(y0.z, z0, x0.z); x0.w = fmaf(y0.w, z0, x0.w);
        x1.x = fmaf(y1.x, z1, x1.x); x1.y = fmaf(y1.y, z1, x1.y); x1.z = fmaf(y1.z, z1, x1.z); x1.w = fmaf(y1.w, z1, x1.w);
      }
      const float4 x = make_float4(x0.x + x1.x, x0.y + x1.y, x0.z + x1.z, x0.w + x1.w);
      for (int d = 0; d < a.dst.n; ++d) *reinterpret_cast<float4*>(a.dst.p[d] + (size_t)rowId * k + 4 * c4) = x;
    }
  } else {
    for (int c = tid; c < k; c += NT) {
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
      const float* yc = Y + c;
      for (int s2 = 0; s2 < np; s2 += 4) {
        x0 = fmaf(yc[(s2 + 0) * pitch], vs[s2 + 0], x0);
        x1 = fmaf(yc[(s2 + 1) * pitch], vs[s2 + 1], x1);
        x2 = fmaf(yc[(s2 + 2) * pitch], vs[s2 + 2], x2);
        x3 = fmaf(yc[(s2 + 3) * pitch], vs[s2 + 3], x3);
      }
      const float x = (x0 + x1) + (x2 + x3);
      for (int d = 0; d < a.dst.n; ++d) a.dst.p[d][(size_t)rowId * k + c] = x;
    }
  }
}

// sub[c,:] = fixed[indx[c],:]  (cpp_utils/als_utils.cc:22-38)
__global__ void gather_rows_kernel(float* __restrict__ sub, const float* __restrict__ fixed,
                                   const int32_t* __restrict__ indx, int cols, int k) {
  const int64_t total = (int64_t)cols * k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / k), f = (int)(e - (int64_t)c * k);
    sub[e] = fixed[(size_t)indx[c] * k + f];
  }
}

}  // namespace ycnr
