// gram_tc.cuh — tcgen05 (5th-gen tensor core) 3xTF32 Gram build for long rows, sm_100a.
//
// Computes, per work item (a slice of <= split_cols ratings of one row), the same tile
// partials as als_primal_kernel<MODE_PARTIAL>:  A = Y^T Y (k x k) and b = Y^T r, as 4x4 tiles
// of the lower triangle plus the rhs tile row (EmfWorker.js:231-232, 238-245).
//
// Split-precision scheme (error-compensated TF32, "3xTF32" with the symmetric halves merged):
//   every fp32 element y is split into  h = tf32(y)  (low 13 mantissa bits cleared) and
//   l = y - h.  One MMA per 8 ratings computes
//       D[128 x 256] += H^T [ H | 2L ]           (M = 128, N = 256, K = 8, kind::tf32)
//   i.e. columns 0..127 accumulate H^T H and columns 128..255 accumulate H^T (2L).  With
//   X = D[:, :128] + D[:, 128:], the symmetrised (X + X^T)/2 = H^T H + H^T L + L^T H, the
//   three significant terms of (H+L)^T (H+L); only L^T L (~2^-22 relative) is dropped.
//   The ratings r ride along as one extra operand column (index KP): row/column KP of the
//   symmetrised X is b.  FP32 accumulation in TMEM throughout.
//
// Operand layout in shared memory (both A and B are "MN-major": the reduction index — the
// rating — is the slow index of the gathered rows).  For MN-major 32-bit operands the only
// legal canonical layout is SWIZZLE_128B_BASE32B: atoms of 4 ratings x 32 factors (512 B),
//   byte = panel*PANEL + (r/4)*512 + (r%4)*128 + ((chunk32 ^ (r%4)) * 32) + (byte % 32)
// (Swizzle<2,5,2>: address bits [5,7) ^= bits [7,9)); LBO = PANEL (next 32 factors),
// SBO = 512 (next 4 ratings); one K = 8 MMA consumes two atoms.  Panels 0..3 hold H (factor
// columns 0..127, zero padded beyond KP), panels 4..7 hold 2L.  A reads M = 128 (panels
// 0..3), B reads N = 256 (panels 0..7) from the same descriptor.
//
// CTA = 1 per SM (all 512 TMEM columns: two 256-column accumulators), warp-specialised:
//   warps 0-3   epilogue: tcgen05.ld -> smem X -> symmetrise -> tile partials to HBM
//   warp  4     TMEM alloc + single-thread MMA issue (tcgen05.mma / tcgen05.commit)
//   warps 5-20  producers: indexed gather (LDG.128) -> split -> swizzled STS, two groups of
//               8 warps filling alternate stages of a 5-deep mbarrier ring
#pragma once
#include "als_kernels.cuh"
#include "common.cuh"

namespace ycnr {

constexpr int kTcStageRows = 32;                              // ratings per stage (4 MMAs of K = 8)
constexpr int kTcStages = 5;
constexpr int kTcPanelBytes = (kTcStageRows / 8) * 1024;      // one 32-column panel of one stage
constexpr int kTcStageBytes = 8 * kTcPanelBytes;              // 4 H panels + 4 2L panels = 32 KB
constexpr int kTcEpiThreads = 128;
constexpr int kTcProdGroup = 256;                             // threads per producer group
constexpr int kTcThreads = kTcEpiThreads + 32 + 2 * kTcProdGroup;  // 672

struct GramTcArgs {
  RowsView rows;
  const float* __restrict__ fixed;
  int k;
  const int32_t* __restrict__ item_row;
  const int32_t* __restrict__ item_off;
  int n_items;
  int split_cols;
  float* __restrict__ partial;   // [items][tiles][16]
  uint32_t variant;              // diagnostics: 1 swaps LBO/SBO, 2/4 raw H^T H / H^T 2L, 8 plain SWIZZLE_128B
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// MN-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout);
// layout_type 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                 uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading byte offset: next 32-column panel
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset: next group of 8 ratings
  d |= 1ull << 46;                                    // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

// kind::tf32, fp32 accumulate, A and B MN-major, M = 128, N = 256 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((256u >> 3) << 17) |
                              ((128u >> 4) << 24);

template <int KT>
__global__ void __launch_bounds__(kTcThreads, 1) gram_tc_kernel(const GramTcArgs a) {
  constexpr int KP = 4 * KT;           // padded system size; the ratings column sits at index KP
  constexpr int NCH = KT + 1;          // 16-byte chunks written per rating: KT data chunks + (val,0,0,0)
  constexpr int NC = KP + 4;           // rows/columns of X that are consumed
  constexpr int NTRI = KT * (KT + 1) / 2;
  constexpr int NTILES = NTRI + KT;
  static_assert(KP + 4 <= 128, "rhs column must fit the M = 128 accumulator");
  constexpr int kTcXsPitch = NC | 1;   // odd: conflict-free row-per-thread stores

  extern __shared__ __align__(1024) uint8_t tc_smem[];
  // carve: [stages][32 KB] | Xs[NC][pitch] | barriers | tmem base
  // SWIZZLE_128B atoms are addressed by absolute shared-memory bits [7,10): align the ring to 1024 B
  uint8_t* stage_base = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
  float* Xs = reinterpret_cast<float*>(stage_base + kTcStages * kTcStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xs + 128 * kTcXsPitch);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kTcStages);
  const uint32_t accf0 = smem_u32(bars + 2 * kTcStages), acce0 = smem_u32(bars + 2 * kTcStages + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // zero every stage once: pad chunks (>= NCH) are never written again
  for (int i = tid; i < kTcStages * kTcStageBytes / 16; i += kTcThreads)
    reinterpret_cast<float4*>(stage_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(full0 + 8 * s, kTcProdGroup);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf0 + 8 * b, 1);
      mbar_init(acce0 + 8 * b, kTcEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  // generic-proxy zero fill must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 5) {
    // ============================ producers ============================
    // Software pipeline per thread, three stages deep, so that ~2 stages of row loads per
    // producer group (4 per SM, ~50 KB) are in flight against HBM latency:
    //   iteration i:  load column ids of stage i+2 | issue row loads of stage i+1 | store stage i
    const int p = tid - (kTcEpiThreads + 32);
    const int group = p / kTcProdGroup;           // fills stages with (global stage index % 2) == group
    const int pt = p - group * kTcProdGroup;
    const int k = a.k;
    constexpr int TASKS = kTcStageRows * NCH;
    constexpr int PER = (TASKS + kTcProdGroup - 1) / kTcProdGroup;
    // a thread's tasks (rating row r, 16-byte chunk q of that row) are the same in every stage
    int r_[PER], q_[PER];
    uint32_t o_[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int task = pt + u * kTcProdGroup;
      r_[u] = -1;
      q_[u] = 0;
      o_[u] = 0;
      if (task < TASKS) {
        const int r = task / NCH, q = task - r * NCH;
        r_[u] = r;
        q_[u] = q;
        if (a.variant & 8u) {   // diagnostics: plain SWIZZLE_128B atoms (8 ratings x 128 B, 16-byte chunks)
          const int r8 = r & 7;
          o_[u] = (uint32_t)(q >> 3) * kTcPanelBytes + (uint32_t)(r >> 3) * 1024u + (uint32_t)r8 * 128u +
                  (uint32_t)(((q & 7) ^ r8) << 4);
        } else {                // SWIZZLE_128B_BASE32B atoms (4 ratings x 128 B, 32-byte chunks)
          const int r4 = r & 3;
          o_[u] = (uint32_t)(q >> 3) * kTcPanelBytes + (uint32_t)(r >> 2) * 512u + (uint32_t)r4 * 128u +
                  (uint32_t)(((((q & 7) >> 1) ^ r4) << 5) | ((q & 1) << 4));
        }
      }
    }
    // iterator over the stages this group fills: (item, stage in item, global stage index)
    struct StageIt {
      int it, st, nst, seg_len;
      int64_t seg_beg;
      uint32_t gs;
      bool valid;
    };
    auto load_item = [&](StageIt& x) {
      if (x.it < a.n_items) {
        const int row = a.item_row[x.it];
        const int off = a.item_off[x.it];
        x.seg_beg = a.rows.row_start[row] + off;
        x.seg_len = min(a.split_cols, a.rows.row_len[row] - off);
        x.nst = (x.seg_len + kTcStageRows - 1) / kTcStageRows;
      }
    };
    auto step = [&](StageIt& x) {   // advance by one global stage
      ++x.st;
      ++x.gs;
      while (x.it < a.n_items && x.st >= x.nst) {
        x.it += gridDim.x;
        x.st = 0;
        load_item(x);
      }
      x.valid = x.it < a.n_items;
    };
    auto next_mine = [&](StageIt& x) {   // advance to the next stage of this group
      do { step(x); } while (x.valid && (int)(x.gs & 1u) != group);
    };
    auto load_idx = [&](const StageIt& x, int (&col)[PER]) {
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        col[u] = -1;
        if (x.valid && r_[u] >= 0 && q_[u] < KT) {
          const int e = x.st * kTcStageRows + r_[u];
          if (e < x.seg_len) col[u] = __ldg(a.rows.indx + x.seg_beg + e);
        }
      }
    };
    auto load_rows = [&](const StageIt& x, const int (&col)[PER], float4 (&v)[PER]) {
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!x.valid || r_[u] < 0) continue;
        if (q_[u] < KT) {
          if (col[u] >= 0 && 4 * q_[u] < k)
            v[u] = __ldg(reinterpret_cast<const float4*>(a.fixed + (size_t)col[u] * k + 4 * q_[u]));
        } else {
          const int e = x.st * kTcStageRows + r_[u];
          if (e < x.seg_len) v[u].x = __ldg(a.rows.vals + x.seg_beg + e);
        }
      }
    };
    StageIt A;
    A.it = blockIdx.x; A.st = -1; A.nst = 0; A.seg_len = 0; A.seg_beg = 0; A.gs = 0xFFFFFFFFu; A.valid = true;
    load_item(A);
    if (A.it >= a.n_items) A.nst = 0;
    next_mine(A);                 // first stage of this group (gs wraps to 0 on the first step)
    StageIt B = A;
    next_mine(B);
    StageIt Cn = B;
    next_mine(Cn);
    int colA[PER], colB[PER], colC[PER];
    float4 vA[PER], vB[PER];
    load_idx(A, colA);
    load_idx(B, colB);
    load_rows(A, colA, vA);
    while (A.valid) {
      load_idx(Cn, colC);
      load_rows(B, colB, vB);
      const uint32_t s = A.gs % kTcStages, ph = (A.gs / kTcStages) & 1u;
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      uint8_t* sb = stage_base + s * kTcStageBytes;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        if (r_[u] < 0) continue;
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(vA[u].x) & 0xFFFFE000u);
        h.y = __uint_as_float(__float_as_uint(vA[u].y) & 0xFFFFE000u);
        h.z = __uint_as_float(__float_as_uint(vA[u].z) & 0xFFFFE000u);
        h.w = __uint_as_float(__float_as_uint(vA[u].w) & 0xFFFFE000u);
        l.x = __uint_as_float(__float_as_uint(2.0f * (vA[u].x - h.x)) & 0xFFFFE000u);
        l.y = __uint_as_float(__float_as_uint(2.0f * (vA[u].y - h.y)) & 0xFFFFE000u);
        l.z = __uint_as_float(__float_as_uint(2.0f * (vA[u].z - h.z)) & 0xFFFFE000u);
        l.w = __uint_as_float(__float_as_uint(2.0f * (vA[u].w - h.w)) & 0xFFFFE000u);
        *reinterpret_cast<float4*>(sb + o_[u]) = h;
        *reinterpret_cast<float4*>(sb + o_[u] + 4 * kTcPanelBytes) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      mbar_arrive(full0 + 8 * s);
      A = B;
      B = Cn;
      next_mine(Cn);
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        vA[u] = vB[u];
        colB[u] = colC[u];
      }
    }
  } else if (warp == 4) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t ltype = (a.variant & 8u) ? 2u : 1u;
      const uint32_t kstride = (a.variant & 8u) ? 1024u : 512u;
      const uint32_t lbo = (a.variant & 1u) ? kstride : (uint32_t)kTcPanelBytes;
      const uint32_t sbo = (a.variant & 1u) ? (uint32_t)kTcPanelBytes : kstride;
      uint32_t gs = 0, itc = 0;
      for (int it = blockIdx.x; it < a.n_items; it += gridDim.x, ++itc) {
        const int row = a.item_row[it];
        const int off = a.item_off[it];
        const int seg_len = min(a.split_cols, a.rows.row_len[row] - off);
        const int nst = (seg_len + kTcStageRows - 1) / kTcStageRows;
        const uint32_t buf = itc & 1u, aph = (itc >> 1) & 1u;
        mbar_wait(acce0 + 8 * buf, aph ^ 1u);       // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t d_tmem = tmem_base + buf * 256u;
        for (int st = 0; st < nst; ++st, ++gs) {
          const uint32_t s = gs % kTcStages, ph = (gs / kTcStages) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const int rows_here = min(kTcStageRows, seg_len - st * kTcStageRows);
          const int ng = (rows_here + 7) >> 3;
          const uint32_t sa = smem_u32(stage_base + s * kTcStageBytes);
          for (int g = 0; g < ng; ++g) {
            const uint64_t desc = tc_smem_desc(sa + (uint32_t)g * 1024u, lbo, sbo, ltype);
            const uint32_t acc = (st > 0 || g > 0) ? 1u : 0u;
            asm volatile(
                "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                "l"(desc), "l"(desc), "r"(kTcIdesc), "r"(acc)
                : "memory");
          }
          // frees the stage for the producers once the MMAs above have read it
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                           empty0 + 8 * s)
                       : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                         accf0 + 8 * buf)
                     : "memory");
      }
    }
    __syncwarp();
  } else {
    // ============================ epilogue ============================
    const int m = tid;  // TMEM lane == row of X; warp w may only touch lanes 32w .. 32w+31
    uint32_t itc = 0;
    for (int it = blockIdx.x; it < a.n_items; it += gridDim.x, ++itc) {
      const uint32_t buf = itc & 1u, aph = (itc >> 1) & 1u;
      mbar_wait(accf0 + 8 * buf, aph);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + buf * 256u;
#pragma unroll 1
      for (int c0 = 0; c0 < NC; c0 += 16) {
        uint32_t h[16], l[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]), "=r"(h[4]), "=r"(h[5]), "=r"(h[6]), "=r"(h[7]),
              "=r"(h[8]), "=r"(h[9]), "=r"(h[10]), "=r"(h[11]), "=r"(h[12]), "=r"(h[13]), "=r"(h[14]), "=r"(h[15])
            : "r"(taddr + (uint32_t)c0));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]), "=r"(l[4]), "=r"(l[5]), "=r"(l[6]), "=r"(l[7]),
              "=r"(l[8]), "=r"(l[9]), "=r"(l[10]), "=r"(l[11]), "=r"(l[12]), "=r"(l[13]), "=r"(l[14]), "=r"(l[15])
            : "r"(taddr + 128u + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (m < NC) {
          const bool use_h = !(a.variant & 4u), use_l = !(a.variant & 2u);   // diagnostics: raw halves
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < NC)
              Xs[m * kTcXsPitch + c0 + j] =
                  (use_h ? __uint_as_float(h[j]) : 0.f) + (use_l ? __uint_as_float(l[j]) : 0.f);
        }
      }
      // accumulator drained: hand it back to the MMA warp before the slower smem -> HBM part
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      mbar_arrive(acce0 + 8 * buf);
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      float* out = a.partial + (size_t)it * NTILES * 16;
      for (int t = tid; t < NTILES; t += kTcEpiThreads) {
        int I, L;
        tile_coords(t, KT, NTRI, I, L);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 o;
          const int ri = 4 * I + i;
          o.x = 0.5f * (Xs[ri * kTcXsPitch + 4 * L + 0] + Xs[(4 * L + 0) * kTcXsPitch + ri]);
          o.y = 0.5f * (Xs[ri * kTcXsPitch + 4 * L + 1] + Xs[(4 * L + 1) * kTcXsPitch + ri]);
          o.z = 0.5f * (Xs[ri * kTcXsPitch + 4 * L + 2] + Xs[(4 * L + 2) * kTcXsPitch + ri]);
          o.w = 0.5f * (Xs[ri * kTcXsPitch + 4 * L + 3] + Xs[(4 * L + 3) * kTcXsPitch + ri]);
          if (a.variant & 6u) {   // diagnostics: unsymmetrised X[4I+i][4L+j]
            o.x = Xs[ri * kTcXsPitch + 4 * L + 0];
            o.y = Xs[ri * kTcXsPitch + 4 * L + 1];
            o.z = Xs[ri * kTcXsPitch + 4 * L + 2];
            o.w = Xs[ri * kTcXsPitch + 4 * L + 3];
          }
          *reinterpret_cast<float4*>(out + (size_t)t * 16 + 4 * i) = o;
        }
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");   // Xs is rewritten by the next item
    }
  }

  // teardown: every tcgen05 op has completed (the epilogue consumed the last accumulator)
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base));
  }
}

template <int KT>
constexpr size_t gram_tc_smem_bytes() {
  return (size_t)kTcStages * kTcStageBytes + (size_t)128 * ((4 * KT + 4) | 1) * sizeof(float) + (2 * kTcStages + 4) * 8 + 16 +
         1024 /* alignment slack */;
}

}  // namespace ycnr
