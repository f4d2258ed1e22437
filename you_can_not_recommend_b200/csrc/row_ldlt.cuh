// row_ldlt.cuh — LDL^T solve of one symmetric positive definite system with the matrix ROWS in registers (sm_100a).
//
// The tile Cholesky of als_kernels.cuh keeps 4x4 tiles in registers and pays, per tile column, two CTA barriers,
// a panel round trip through shared memory and — above all — issue slots of warps most of whose lanes own tiles
// of finished columns (ncu, k x k solve at k = 100: 28 K warp instructions per system for 5 K warp-FMAs of
// algorithmic work, 49 % barrier stalls).  Here a lane owns whole rows: the n rows are cut into blocks of H <= 32
// rows, block q holds rows H q .. H q + H - 1 on lanes 0 .. H - 1 of warp q.  Step j of the
// right-looking LDL^T:
//     every lane publishes entry j of its rows (and its right-hand side) ............ 1 STS per block
//     barrier (__syncwarp, or one bar.sync of the system's warps)
//     pivot 1 / d_j (MUFU.RCP; LDL^T needs no square root), t = a_rj / d_j ........... L[r][j], parked in shared memory
//     a_rc -= t * a_cj for the remaining columns: the column arrives as broadcast LDS.128, the update is packed
//     fma.rn.f32x2 (two columns per instruction) .................................... 2 FFMA2 per LDS.128 and block
// Block q only needs columns <= H q + H - 1 (lower triangle by block); the bound is warp-uniform and checked once
// per 16 columns.  Register indices must be compile-time constants, so the rows are kept RELATIVE to the current
// 2-column panel: the second step of a panel writes its results two registers down (an FMA's destination is free),
// the column is published relative to the panel too (aligned LDS.128 for any j), and the panel loop is an ordinary
// runtime loop around one small unrolled body — small matters: the L1.5 instruction cache holds 32 KB, a first
// version with a 4-step body per pair of blocks (~48 KB of hot code) ran 3x SLOWER than the tile kernel.
// The forward substitution rides along (one FMA per step); the back substitution runs in one warp: x_j by shuffle,
// L[j][r] from shared memory.
#pragma once
#include <utility>
#include "common.cuh"

namespace ycnr {

// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>) — register arrays of more than
// ~128 elements only stay in registers when every index is a constant the front end can see
template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kRowColStride = 144;   // floats per published-column buffer: 128 rows + the overshoot of a 16-column read

// shared memory of a solve, in floats: G rows (pitch n | 1, odd: conflict-free for lanes on consecutive rows; the
// same place then takes L column by column) + the slack of a 16-column read | col[2] | colb[2] | w
__host__ __device__ constexpr int row_ldlt_pitch(int n) { return n | 1; }
__host__ __device__ constexpr int row_ldlt_smem_floats(int n) {
  return ((n * row_ldlt_pitch(n) + 128 + 3) & ~3) + 4 * kRowColStride + 128;
}

// 16 entries [BASE, BASE + 16) of a register row (relative to the current 2-column panel): a_c -= t * col_c for the
// entries > JJ; the second step of a panel (JJ == 1) stores two registers down.  cv: the column, broadcast LDS.128
template <int BASE, int JJ, int W>
__device__ __forceinline__ void row_update16(float (&r)[W], const float4 (&cv)[4], float nt) {
  static_assert(BASE % 16 == 0 && BASE + 16 <= W && (JJ == 0 || JJ == 1), "segment outside the register row");
  const float2 nt2 = make_float2(nt, nt);
  static_for<8>([&](auto ec) {
    constexpr int e = 2 * decltype(ec)::value;
    constexpr int i = BASE + e;          // even; the pair (i, i + 1) comes from one half of an LDS.128
    constexpr int g = e >> 2;
    const float c0 = (e & 2) ? cv[g].z : cv[g].x;
    const float c1 = (e & 2) ? cv[g].w : cv[g].y;
    if constexpr (i > JJ) {
      constexpr int d = JJ == 1 ? i - 2 : i;
      const float2 s = __ffma2_rn(nt2, make_float2(c0, c1), make_float2(r[i], r[i + 1]));
      r[d] = s.x;
      r[d + 1] = s.y;
    } else if constexpr (i == JJ) {      // JJ == 0, pair (0, 1): only the odd partner is still live
      r[1] = fmaf(nt, c1, r[1]);
    }
  });
}

// Forward part (L, D and the forward substitution) for the warp that owns rows r = row0 + lane, lane < H.
//   n: system size; cmax: last column the block needs (min(n, row0 + H) - 1); NSEG: 16-column segments of the
//   register row (16 NSEG >= cmax + 1)
//   ra: the lane's row, G[r][0 .. 16 NSEG) (entries beyond cmax: anything); z: right-hand side of row r
//   col, colb: two buffers of kRowColStride floats each; Ls: L[.][j] is parked at Ls[j * lpitch + row]
//   sync(): barrier over all warps of the system (every warp calls this function for its own block)
// Returns (D^-1 L^-1 b)[r].  Lanes >= H and rows >= n take part in the barriers and return 0.
template <int NSEG, class SyncF>
__device__ __forceinline__ float row_ldlt_forward(const int n, const int H, const int row0, const int lane, float* col,
                                                  float* colb, float* Ls, const int lpitch, float (&ra)[16 * NSEG],
                                                  float z, SyncF&& sync) {
  const int r = row0 + lane;
  const bool own = lane < H && r < n;
  const int cmax = min(n, row0 + H) - 1;
  float dinv = 0.f;
  if (!own) z = 0.f;
  for (int P2 = 0; P2 < n; P2 += 2) {
    static_for<2>([&](auto jc) {
      constexpr int jj = decltype(jc)::value;
      const int j = P2 + jj;
      if (j < n) {
        float* cb = col + jj * kRowColStride;
        float* cbb = colb + jj * kRowColStride;
        if (own && r >= j) {   // published relative to the panel: the reads below are aligned LDS.128
          cb[r - P2] = ra[jj];
          cbb[r - P2] = z;
        }
        sync();
        if (j <= cmax) {       // warp-uniform: blocks above the pivot row are finished
          const float inv = rcp_fast(cb[jj]);
          const float bj = cbb[jj];
          const float t = ra[jj] * inv;
          if (own && r > j) {
            z = fmaf(-t, bj, z);
            Ls[j * lpitch + r] = t;
          }
          if (own && r == j) dinv = inv;
          static_for<NSEG>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            if (P2 + 16 * s <= cmax) {
              float4 cv[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) cv[g] = *reinterpret_cast<const float4*>(cb + 16 * s + 4 * g);
              row_update16<16 * s, jj, 16 * NSEG>(ra, cv, -t);
            }
          });
        }
      }
    });
  }
  return z * dinv;
}

// Back substitution L^T x = w by ONE warp that holds w of all NW blocks (wv[q]: row H q + lane).  On return wv = x.
template <int NW>
__device__ __forceinline__ void row_ldlt_backward(const int n, const int H, const int lane, const float* Ls,
                                                  const int lpitch, float (&wv)[NW]) {
  int qj = (n - 1) / H, lj = (n - 1) - qj * H;
  for (int j = n - 1; j > 0; --j) {
    float ws = wv[0];
#pragma unroll
    for (int q = 1; q < NW; ++q)
      if (qj == q) ws = wv[q];
    const float xj = __shfl_sync(0xffffffffu, ws, lj);
#pragma unroll
    for (int q = 0; q < NW; ++q) {
      const int r = H * q + lane;
      if (lane < H && r < j && H * q < j) wv[q] = fmaf(-Ls[r * lpitch + j], xj, wv[q]);   // L[j][r]
    }
    if (--lj < 0) { --qj; lj = H - 1; }
  }
}

}  // namespace ycnr
