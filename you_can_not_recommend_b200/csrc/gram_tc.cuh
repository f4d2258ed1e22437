// gram_tc.cuh — tcgen05 (5th-gen tensor core) 3xTF32 Gram build for long rows, sm_100a.
//
// Computes, per work item (a slice of <= split_cols ratings of one row), the same tile
// partials as als_primal_kernel<MODE_PARTIAL>:  A = Y^T Y (k x k) and b = Y^T r, as 4x4 tiles
// of the lower triangle plus the rhs tile row (EmfWorker.js:231-232, 238-245).
//
// Split-precision scheme (error-compensated TF32, "3xTF32" with the symmetric halves merged):
//   every fp32 element y is split into  h = tf32(y)  (low 13 mantissa bits cleared) and
//   l = y - h (exact).  One MMA per 8 ratings computes
//       D[128 x 2*NC] += H^T [ H | L ]          (M = 128, N = 2*NC, K = 8, kind::tf32)
//   i.e. columns 0..NC-1 accumulate H^T H and columns NC..2NC-1 accumulate H^T L.  With
//   X = D[:, :NC] + 2 D[:, NC:], the symmetrised (X + X^T)/2 = H^T H + H^T L + L^T H, the
//   three significant terms of (H+L)^T (H+L); only L^T L (~2^-22 relative) is dropped.
//   The ratings r ride along as one extra operand column (index KP): row/column KP of the
//   symmetrised X is b.  FP32 accumulation in TMEM throughout.  NC = KP + 4 rounded up to 8,
//   so the MMA is exactly as wide as the system needs (N = 208 at k = 100, not 256).
//
// Operand layout in shared memory (both A and B are "MN-major": the reduction index — the
// rating — is the slow index of the gathered rows).  For MN-major 32-bit operands the
// canonical layout is SWIZZLE_128B_BASE32B: atoms of 4 ratings x 32 columns (512 B),
//   byte = panel*4096 + (r/4)*512 + (r%4)*128 + ((chunk32 ^ (r%4)) * 32) + (byte % 32)
// LBO = 4096 (next 32 columns), SBO = 512 (next 4 ratings); one K = 8 MMA consumes two atoms.
// A stage holds 32 ratings x NPAN panels: columns [0, NC) = H, [NC, 2NC) = L.  A reads
// M = 128 columns and B reads N = 2*NC columns from the same descriptor.
//
// Data movement: the gathered factor rows go HBM/L2 -> shared memory with 16-byte cp.async
// straight into their swizzled H position (no register staging); a loader warp requests the
// column ids of its next stage before it issues the current one.  The H columns keep the raw fp32 bits: kind::tf32 ignores
// the low 13 mantissa bits of its operands (verified bit-for-bit on B200 against explicitly
// masked operands), so only the tail l = y - tf32(y) has to be produced by CUDA cores:
// one LDS, one STS and 8 ALU ops per 16 bytes.
//
// CTA = 1 per SM (all 512 TMEM columns: two 256-column accumulators), warp-specialised (29 warps at k = 100):
//   warps 0-3    epilogue: tcgen05.ld -> smem X -> symmetrise -> tile partials to HBM
//   warp  4      TMEM alloc + single-thread MMA issue (tcgen05.mma / tcgen05.commit)
//   next 2*STAGES warps  loaders: two per ring slot, each moving one half (16 ratings) of every stage of
//                its slot; one cp.async instruction moves one rating (lane = 16-byte chunk), column ids are
//                broadcast by shuffle.  Completion is signalled per stage with cp.async.mbarrier.arrive;
//                loaders never execute a fence, so up to STAGES stages of copies stay in flight.  When the
//                fixed matrix does not fit L2 (byItem gathers from U), each loader also issues
//                prefetch.global.L2 for the rows it will gather two ring cycles later, so the bytes in flight
//                against HBM latency are not capped by the ring's shared memory
//   next 2*STAGES warps  splitters: two per ring slot: tail columns (four LDS.128 in flight),
//                fence.proxy.async, hand the stage to the MMA warp
// A warp always works on the same slot: seeing every use of "its" slot in order is what makes waiting on an
// mbarrier phase PARITY sound.
//
// Measured history (B200, MAL k = 100, byUser 69 M + byItem 116 M ratings per iteration):
//   v1  LDG -> registers -> STS producers, one warp per slot ........................ 32.0 ms
//   v2  cp.async into the swizzled layout, 1 loader + 1 splitter warp per slot ....... 27.1 ms
//       same, 12 stages of 16 ratings ............................................... 28.2 ms
//   v3  register-staged producers, 2 warps per slot .................................. 29.8 ms
//   v4  (this file) 2 loaders + 2 splitters per slot, 4-deep LDS, L2 prefetch ........ 21.7 ms
//   v5  register-staged producers, 4 warps per slot + L2 prefetch (ld.global.cg) ..... 25.1 ms
// ncu source view of v2: a slot's cycle was loader issue 1.8k clk -> arrival -> splitter 3k clk (LDS latency
// exposed) -> MMA; loaders spent 77 % and the epilogue 97 % of their samples waiting, i.e. the ring was bound
// by the two producer legs, which v4 halves.  v4: tensor pipe 42-50 % active, LSU data pipe 64 %.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encode function is fetched through cudaGetDriverEntryPoint)

#include "als_kernels.cuh"
#include "common.cuh"

namespace ycnr {

constexpr int kTcStageRows = 32;                              // ratings per stage (4 MMAs of K = 8)
constexpr int kTcPanelBytes = (kTcStageRows / 8) * 1024;      // one 32-column panel of one stage
constexpr int kTcEpiThreads = 128;
constexpr int kTcPrefetchUses = 2;                            // L2 prefetch distance in ring cycles
// One loader warp and one splitter warp per ring slot: a warp sees every use of "its" slot in
// order, which is what makes waiting on an mbarrier phase PARITY sound (a warp that hopped between
// slots could be two phases off and sail through a wait).
constexpr int kTcSmemLimit = 232448;                          // 227 KB per CTA on sm_100
// Measured on B200 (MAL, k = 100): 6 stages x 32 ratings 27.3 ms per iteration, 12 x 16 ratings 28.2 ms —
// the ring is not latency-bound per slot, the shared-memory pipe is the limiter.
constexpr int kTcMaxStages = 6;

template <int KT>
struct TcCfg {
  static constexpr int KP = 4 * KT;            // padded system size; the ratings column sits at index KP
  static constexpr int NCH = KT + 1;           // 16-byte chunks per rating: KT data chunks + (val,0,0,0)
  static constexpr int NC = (KP + 4 + 7) & ~7; // columns of H (and of L) seen by the MMA
  static constexpr int N = 2 * NC;             // MMA N
  static constexpr int NPAN = (N + 31) / 32 < 4 ? 4 : (N + 31) / 32;   // A reads 4 panels (M = 128)
  static constexpr int STAGE_BYTES = NPAN * kTcPanelBytes;
  static constexpr int XP = NC | 1;            // odd pitch: conflict-free row-per-thread stores
  static constexpr int XS_BYTES = (NC * XP * 4 + 15) & ~15;
  static constexpr int STAGES_FIT = (kTcSmemLimit - 2048 - XS_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT < kTcMaxStages ? STAGES_FIT : kTcMaxStages;
  static constexpr int THREADS = kTcEpiThreads + 32 + 4 * 32 * STAGES;   // epilogue | MMA | 2 loaders + 2 splitters per slot
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + XS_BYTES + (3 * STAGES + 4) * 8 + 16 + 1024;
  static_assert(KP + 4 <= 128, "rhs column must fit the M = 128 accumulator");
  static_assert(NCH <= 32, "one lane per 16-byte chunk of a rating");
  static_assert(N <= 256 && N % 16 == 0, "MMA N");
  static_assert(STAGES >= 4, "not enough shared memory for the stage ring");
};

struct GramTcArgs {
  RowsView rows;
  const float* __restrict__ fixed;
  int k;
  const int32_t* __restrict__ item_row;
  const int32_t* __restrict__ item_off;
  const int32_t* __restrict__ item_order;   // j-th work item to process (longest slices first), or nullptr = identity
  int n_items;
  int split_cols;
  float* __restrict__ partial;   // [items][tiles][16]
  int prefetch;                  // 1: the fixed matrix does not fit L2, loaders prefetch far-ahead rows into it
  // Column map of the virtual system the kernel builds: its first chunks_a 16-byte chunks are columns
  // [col_a, col_a + 4 chunks_a) of the gathered rows, the remaining ones start at col_b; columns >= k read as zero.
  // One pass over all columns (k <= 124): col_a = 0, chunks_a = KT.  Systems with k > 124 are covered by one
  // pass per PAIR of column blocks (blocks of w <= 60 columns, 2 w + 4 <= 128), see launch_primal_blocks.
  int col_a, col_b, chunks_a;
  // TMA gather (one-pass systems): tensor map over the fixed matrix [fixed_rows x k] fp32, box {32 columns, 1 row},
  // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  One cp.async.bulk.tensor.2d...tile::gather4 lands four gathered rows x 32
  // columns as exactly one operand atom (verified with scripts/micro/tma_gather4_probe.cu: row r of the tile at
  // byte r * 128, its 32-byte chunks XOR-swizzled by r, columns >= k zero-filled) — without LSU wavefronts.
  CUtensorMap tmap;
  int use_tma;
  int fixed_rows;
  uint32_t variant;              // diagnostics: 16 = splitters also overwrite the H columns with explicitly masked values
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
// layout_type 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                 uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading byte offset: next 32-column panel
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset: next group of 4 ratings
  d |= 1ull << 46;                                    // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

// byte offset of the 16-byte chunk that starts at column c (multiple of 4) of rating r inside a stage
__device__ __forceinline__ uint32_t tc_chunk_offset(int r, int c) {
  return (uint32_t)(c >> 5) * kTcPanelBytes + (uint32_t)(r >> 2) * 512u + (uint32_t)(r & 3) * 128u +
         (uint32_t)(((((c & 31) >> 3) ^ (r & 3)) << 5) | (((c >> 2) & 1) << 4));
}

// Walks the CTA's work items stage by stage (positions blockIdx.x, +gridDim.x, ... of item_order; the
// host never emits an empty slice).  item_order lists the slices longest first, so this static round-robin
// hands every CTA one slice of each length band: the CTAs' totals differ by at most one slice (~0.5 %)
// instead of the ~7-11 % spread of the natural order.  Every thread of a role walks it redundantly: the descriptor loads are
// warp-uniform and cache-resident, and the descriptor of the following item is requested one
// item early so its latency is off the critical path.
struct TcItemIter {
  int it, st, nst, seg_len;
  int64_t seg_beg;
  int n_seg_len;
  int64_t n_seg_beg;
  __device__ __forceinline__ void fetch(const GramTcArgs& a, int item, int64_t& beg, int& len) {
    beg = 0;
    len = 0;
    if (item < a.n_items) {
      const int it0 = a.item_order ? __ldg(a.item_order + item) : item;
      const int row = __ldg(a.item_row + it0);
      const int off = __ldg(a.item_off + it0);
      beg = __ldg(a.rows.row_start + row) + off;
      len = max(0, min(a.split_cols, __ldg(a.rows.row_len + row) - off));
    }
  }
  __device__ __forceinline__ void init(const GramTcArgs& a) {
    it = blockIdx.x;
    st = 0;
    fetch(a, it, seg_beg, seg_len);
    nst = (seg_len + kTcStageRows - 1) / kTcStageRows;
    fetch(a, it + gridDim.x, n_seg_beg, n_seg_len);
  }
  __device__ __forceinline__ bool valid(const GramTcArgs& a) const { return it < a.n_items; }
  __device__ __forceinline__ void next_item(const GramTcArgs& a) {
    it += gridDim.x;
    st = 0;
    seg_beg = n_seg_beg;
    seg_len = n_seg_len;
    nst = (seg_len + kTcStageRows - 1) / kTcStageRows;
    fetch(a, it + gridDim.x, n_seg_beg, n_seg_len);
  }
  __device__ __forceinline__ void next(const GramTcArgs& a) {
    if (++st >= nst) next_item(a);
  }
  __device__ __forceinline__ void advance(const GramTcArgs& a, int stages) {
    st += stages;
    while (it < a.n_items && st >= nst) {
      const int over = st - nst;
      next_item(a);
      st = over;
    }
  }
};

template <int KT>
__global__ void __launch_bounds__(TcCfg<KT>::THREADS, 1) gram_tc_kernel(const __grid_constant__ GramTcArgs a) {
  using Cfg = TcCfg<KT>;
  constexpr int NCH = Cfg::NCH, NC = Cfg::NC, XP = Cfg::XP;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NTRI = KT * (KT + 1) / 2;
  constexpr int NTILES = NTRI + KT;
  // kind::tf32, fp32 accumulate, A and B MN-major, M = 128, N = 2*NC (cute::UMMA::InstrDescriptor)
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(Cfg::N >> 3) << 17) | ((128u >> 4) << 24);

  if (rows_poisoned(a.rows)) return;   // uniform for the whole grid: nothing has been allocated yet
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  // carve: [stages][STAGE_BYTES] | Xs[NC][XP] | barriers | tmem base
  // SWIZZLE_128B atoms are addressed by absolute shared-memory bits [7,10): align the ring to 1024 B
  uint8_t* stage_base = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
  float* Xs = reinterpret_cast<float*>(stage_base + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + STAGES * Cfg::STAGE_BYTES + Cfg::XS_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), raw0 = smem_u32(bars + 2 * STAGES);
  const uint32_t accf0 = smem_u32(bars + 3 * STAGES), acce0 = smem_u32(bars + 3 * STAGES + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // zero every stage once: pad columns are never written again
  for (int i = tid; i < STAGES * Cfg::STAGE_BYTES / 16; i += Cfg::THREADS)
    reinterpret_cast<float4*>(stage_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 2);     // two splitter warps
      mbar_init(empty0 + 8 * s, 1);
      // cp.async path: every lane of the two loader warps arrives when its copies have landed;
      // TMA path: one arrive.expect_tx per loader warp, the copy engine completes the bytes
      mbar_init(raw0 + 8 * s, a.use_tma ? 2 : 64);
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
    // Two loader warps and two splitter warps per ring slot, each owning one half (HR ratings) of every
    // stage of the slot.  ncu (source view, MAL byItem, one warp of each per slot): a slot's cycle was
    // loader issue 1.8k clk -> arrival -> splitter 3k clk (LDS latency exposed) -> MMA; the epilogue and
    // the loaders spent 97 % / 77 % of their samples waiting.  Halving the two producer legs and keeping
    // four LDS.128 in flight in the splitter shortens the cycle; the slots stay in order per warp.
    constexpr int HR = kTcStageRows / 2;
    const int pw = warp - 5;
    const bool loader = pw < 2 * STAGES;
    const int s_own = (loader ? pw : pw - 2 * STAGES) >> 1, half = pw & 1;
    const int k = a.k;
    // lane = 16-byte chunk of a rating (lanes 0..KT-1: factors, lane KT: the rating value).
    // Offsets of that chunk for the four values of r % 4 (the swizzle phase); + (r / 4) * 512.
    uint32_t oh4[4], ol4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      oh4[j] = tc_chunk_offset(j, 4 * min(lane, NCH - 1));
      ol4[j] = tc_chunk_offset(j, NC + 4 * min(lane, NCH - 1));
    }
    uint8_t* const sb = stage_base + s_own * Cfg::STAGE_BYTES + half * (HR / 4) * 512;   // this warp's half of the slot
    TcItemIter ix;
    ix.init(a);
    ix.advance(a, s_own);
    if (loader) {
      // ============================ loaders ============================
      // lane r < HR: column id of rating half*HR + r of the stage under the iterator
      auto load_ids = [&](const TcItemIter& x, int& col, uint32_t& vm, int64_t& e0) {
        col = 0;
        e0 = 0;
        bool ok = false;
        if (x.valid(a)) {
          e0 = x.seg_beg + (int64_t)x.st * kTcStageRows + half * HR;
          ok = lane < HR && x.st * kTcStageRows + half * HR + lane < x.seg_len;
          if (ok) col = __ldg(a.rows.indx + e0 + lane);
        }
        vm = __ballot_sync(0xffffffffu, ok);
      };
      const int src_col = lane < a.chunks_a ? a.col_a + 4 * lane : a.col_b + 4 * (lane - a.chunks_a);
      const float* src_lane = a.fixed + src_col;
      const bool lane_ok = lane < KT && src_col < k;
      const uint32_t val_off = tc_chunk_offset(lane & (HR - 1), 4 * KT);
      auto issue = [&](uint32_t use, int col, uint32_t vm, int64_t e0) {
        mbar_wait(empty0 + 8 * s_own, (use & 1u) ^ 1u);
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const int c = __shfl_sync(0xffffffffu, col, r);
          const bool ok = ((vm >> r) & 1u) && lane_ok;
          if (lane < KT) cp_async16(sb + oh4[r & 3] + (r >> 2) * 512, src_lane + (size_t)c * k, ok ? 16 : 0);
        }
        if (lane < HR) {   // the rating values: lane r -> column KP of rating r, (val, 0, 0, 0)
          const bool ok = (vm >> lane) & 1u;
          cp_async4(sb + val_off, a.rows.vals + (ok ? e0 + lane : 0), ok ? 4 : 0);
        }
        // this lane's arrival on the stage's raw barrier fires when its copies have landed
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(raw0 + 8 * s_own) : "memory");
      };
      // TMA variant of issue(): lane l < 16 lands the atom (group g = l / 4 of four ratings, panel p = l % 4).
      const bool tma = a.use_tma != 0;
      const int hp = (k + 31) >> 5;                     // panels that hold real columns
      const uint64_t tmap_addr = reinterpret_cast<uint64_t>(&a.tmap);
      auto issue_tma = [&](uint32_t use, int col, uint32_t vm) {
        mbar_wait(empty0 + 8 * s_own, (use & 1u) ^ 1u);
        const int g = (lane >> 2) & 3, p = lane & 3;
        int ir[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = __shfl_sync(0xffffffffu, col, 4 * g + q);
          ir[q] = ((vm >> (4 * g + q)) & 1u) ? c : a.fixed_rows;      // out-of-range row: zero-filled by the copy engine
        }
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(raw0 + 8 * s_own),
                       "r"((uint32_t)(4 * hp * 512))
                       : "memory");
        __syncwarp();
        if (lane < 16 && p < hp)
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
                  smem_u32(sb + p * kTcPanelBytes + g * 512)),
              "l"(tmap_addr), "r"(raw0 + 8 * s_own), "r"(32 * p), "r"(ir[0]), "r"(ir[1]), "r"(ir[2]), "r"(ir[3])
              : "memory");
      };
      // L2 prefetch of the rows this warp will gather kTcPrefetchUses ring cycles from now: the bytes in
      // flight against HBM latency are no longer capped by the ring's shared memory (6 x 32 ratings x 400 B).
      // Lane l asks for 128-byte line (l >> 4) and (l >> 4) + 2 of rating l & 15: a 400-byte row that starts
      // on a 16-byte boundary touches exactly the 4 lines at +0, +128, +256, +384.
      const bool do_pf = a.prefetch != 0;
      TcItemIter px = ix;
      if (do_pf) px.advance(a, kTcPrefetchUses * STAGES);
      auto load_pf_ids = [&](const TcItemIter& x, int& col) {
        col = -1;
        if (x.valid(a)) {
          const int rr = half * HR + (lane & (HR - 1));
          if (x.st * kTcStageRows + rr < x.seg_len)
            col = __ldg(a.rows.indx + x.seg_beg + (int64_t)x.st * kTcStageRows + rr);
        }
      };
      const bool one_range = a.chunks_a >= KT;
      auto prefetch_rows = [&](int col) {
        if (col >= 0) {
          if (one_range) {
            const char* p = reinterpret_cast<const char*>(a.fixed + (size_t)col * k) + (lane >> 4) * 128;
            if ((lane >> 4) * 128 < k * 4) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
            if ((lane >> 4) * 128 + 256 < k * 4) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p + 256));
          } else {   // two column ranges of at most 240 bytes: two lines each (the lines they start in and the next)
            const char* row = reinterpret_cast<const char*>(a.fixed + (size_t)col * k);
            const int oa = a.col_a * 4 + (lane >> 4) * 128, ob = a.col_b * 4 + (lane >> 4) * 128;
            if (oa < k * 4) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(row + oa));
            if (ob < k * 4) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(row + ob));
          }
        }
      };
      int pcol = -1;
      if (do_pf) load_pf_ids(px, pcol);
      // two register sets (A/B): the ids of the warp's next stage are requested before the current
      // stage is issued and are never moved while the load is outstanding
      int colA, colB;
      uint32_t vmA, vmB;
      int64_t eA, eB;
      uint32_t use = 0;
      load_ids(ix, colA, vmA, eA);
      while (ix.valid(a)) {
        ix.advance(a, STAGES);
        load_ids(ix, colB, vmB, eB);
        if (do_pf) {
          prefetch_rows(pcol);
          px.advance(a, STAGES);
          load_pf_ids(px, pcol);
        }
        if (tma) issue_tma(use++, colA, vmA);
        else issue(use++, colA, vmA, eA);
        if (!ix.valid(a)) break;
        ix.advance(a, STAGES);
        load_ids(ix, colA, vmA, eA);
        if (do_pf) {
          prefetch_rows(pcol);
          px.advance(a, STAGES);
          load_pf_ids(px, pcol);
        }
        if (tma) issue_tma(use++, colB, vmB);
        else issue(use++, colB, vmB, eB);
      }
      cp_async_wait<0>();
    } else {
      // ============================ splitters ============================
      const bool mask_head = (a.variant & 16u) != 0;
      const bool tma = a.use_tma != 0;
      for (uint32_t use = 0; ix.valid(a); ++use, ix.advance(a, STAGES)) {
        // TMA path: the copy engine zero-fills column KP, so the ratings are fetched here (lane r: rating r of this
        // half, requested before the wait) and written as the (val, 0, 0, 0) chunk by the lane that owns it
        float myval = 0.f;
        if (tma && lane < HR) {
          const int idx = ix.st * kTcStageRows + half * HR + lane;
          if (idx < ix.seg_len) myval = __ldg(a.rows.vals + ix.seg_beg + idx);
        }
        mbar_wait(raw0 + 8 * s_own, use & 1u);
        if (a.variant & 32u) {   // diagnostics only (wrong numbers): no tail split — what the kernel costs without it
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(full0 + 8 * s_own);
          continue;
        }
#pragma unroll
        for (int r0 = 0; r0 < HR; r0 += 4) {
          float vj[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) vj[j] = __shfl_sync(0xffffffffu, myval, r0 + j);
          if (lane < NCH) {
            float4 v[4];
            const bool val_lane = tma && lane == NCH - 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (val_lane) v[j] = make_float4(vj[j], 0.f, 0.f, 0.f);
              else v[j] = *reinterpret_cast<const float4*>(sb + oh4[j] + (r0 >> 2) * 512);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 h, l;
              h.x = __uint_as_float(__float_as_uint(v[j].x) & 0xFFFFE000u);
              h.y = __uint_as_float(__float_as_uint(v[j].y) & 0xFFFFE000u);
              h.z = __uint_as_float(__float_as_uint(v[j].z) & 0xFFFFE000u);
              h.w = __uint_as_float(__float_as_uint(v[j].w) & 0xFFFFE000u);
              l.x = v[j].x - h.x;
              l.y = v[j].y - h.y;
              l.z = v[j].z - h.z;
              l.w = v[j].w - h.w;
              if (mask_head) *reinterpret_cast<float4*>(sb + oh4[j] + (r0 >> 2) * 512) = h;
              else if (val_lane) *reinterpret_cast<float4*>(sb + oh4[j] + (r0 >> 2) * 512) = v[j];
              *reinterpret_cast<float4*>(sb + ol4[j] + (r0 >> 2) * 512) = l;
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * s_own);
      }
    }
  } else if (warp == 4) {
    // ============================ MMA issuer ============================
    // The whole warp walks the loop convergently and one elected lane issues (elect.sync inside the asm
    // block, as CUTLASS's SM100 atoms do).  An earlier version ran the loop under `if (lane == 0)`: in a
    // divergent region the compiler cannot keep the descriptors in uniform registers and wrapped every
    // tcgen05.mma / commit in an ELECT + R2UR + BRA.U.ANY loop — ~130 dependent scalar instructions
    // (~800 clk) per stage; ncu's source view showed this thread busy 75 % of the time and waiting on `full`
    // only 4 %: the issue loop, not the producers or the tensor pipe, bounded the kernel.
    // The issue loop is a serial chain of dependent scalar instructions (~5 clk each) next to a tensor pipe
    // that needs 4 x ~114 clk per stage, so it is kept short: ring position and phase are counters (no
    // divisions), the descriptor is one constant high word plus a low word that advances by a constant per
    // stage / per K-group, and a full stage (the common case) is ONE asm block: elect, 4 MMAs, commit.
    TcItemIter ix;
    ix.init(a);
    uint32_t itc = 0;
    uint32_t s = 0, ph = 0;                                   // ring slot and its phase
    const uint64_t desc0 = tc_smem_desc(smem_u32(stage_base), kTcPanelBytes, 512u, 1u);
    const uint32_t desc_hi = (uint32_t)(desc0 >> 32);
    const uint32_t desc_lo0 = (uint32_t)desc0;                // + (byte offset >> 4): never carries out of the address field
    constexpr uint32_t kStageLo = Cfg::STAGE_BYTES >> 4, kGroupLo = 1024u >> 4;
    static_assert(kTcStageRows == 32, "the full-stage block below issues exactly 4 MMAs");
    while (ix.valid(a)) {
      const int nst = ix.nst, seg_len = ix.seg_len;
      const uint32_t buf = itc & 1u, aph = (itc >> 1) & 1u;
      ++itc;
      mbar_wait(acce0 + 8 * buf, aph ^ 1u);       // epilogue drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t d_tmem = tmem_base + buf * 256u;
      const int nfull = seg_len / kTcStageRows;   // full stages first, then at most one short stage
      for (int st = 0; st < nst; ++st) {
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t lo = desc_lo0 + s * kStageLo;
        const uint32_t bar = empty0 + 8 * s;
        if (st < nfull) {
          asm volatile(
              "{\n .reg .pred p, q, t;\n .reg .b32 l1, l2, l3;\n .reg .b64 d0, d1, d2, d3;\n"
              " setp.ne.b32 p, %4, 0;\n setp.eq.b32 t, 0, 0;\n"
              " add.u32 l1, %1, %5;\n add.u32 l2, l1, %5;\n add.u32 l3, l2, %5;\n"
              " mov.b64 d0, {%1, %2};\n mov.b64 d1, {l1, %2};\n mov.b64 d2, {l2, %2};\n mov.b64 d3, {l3, %2};\n"
              " elect.sync _|q, 0xffffffff;\n"
              " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d0, d0, %3, p;\n"
              " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d1, d1, %3, t;\n"
              " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d2, d2, %3, t;\n"
              " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d3, d3, %3, t;\n"
              " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n}\n" ::"r"(d_tmem),
              "r"(lo), "r"(desc_hi), "r"(IDESC), "r"(st > 0 ? 1u : 0u), "n"(kGroupLo), "r"(bar)
              : "memory");
        } else {
          const int rows_here = seg_len - st * kTcStageRows;
          const int ng = (rows_here + 7) >> 3;
          for (int g = 0; g < ng; ++g) {
            const uint32_t acc = (st > 0 || g > 0) ? 1u : 0u;
            asm volatile(
                "{\n .reg .pred p, q;\n .reg .b64 d;\n setp.ne.b32 p, %4, 0;\n mov.b64 d, {%1, %2};\n"
                " elect.sync _|q, 0xffffffff;\n"
                " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d, d, %3, p;\n}\n" ::"r"(d_tmem),
                "r"(lo + (uint32_t)g * kGroupLo), "r"(desc_hi), "r"(IDESC), "r"(acc)
                : "memory");
          }
          // frees the stage for the producers once the MMAs above have read it
          asm volatile(
              "{\n .reg .pred q;\n elect.sync _|q, 0xffffffff;\n"
              " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar)
              : "memory");
        }
        if (++s == (uint32_t)STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      asm volatile(
          "{\n .reg .pred q;\n elect.sync _|q, 0xffffffff;\n"
          " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(accf0 + 8 * buf)
          : "memory");
      ix.next_item(a);
    }
    __syncwarp();
  } else {
    // ============================ epilogue ============================
    const int m = tid;  // TMEM lane == row of X; warp w may only touch lanes 32w .. 32w+31
    constexpr int EPI_TPT = (NTILES + kTcEpiThreads - 1) / kTcEpiThreads;
    int eI[EPI_TPT], eL[EPI_TPT];   // this thread's output tiles, the same for every item
#pragma unroll
    for (int j = 0; j < EPI_TPT; ++j) {
      eI[j] = eL[j] = 0;
      if (tid + j * kTcEpiThreads < NTILES) tile_coords(tid + j * kTcEpiThreads, KT, NTRI, eI[j], eL[j]);
    }
    uint32_t itc = 0;
    for (int it = blockIdx.x; it < a.n_items; it += gridDim.x) {
      const uint32_t buf = itc & 1u, aph = (itc >> 1) & 1u;
      ++itc;
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
            : "r"(taddr + (uint32_t)NC + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (m < NC) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < NC) Xs[m * XP + c0 + j] = fmaf(2.0f, __uint_as_float(l[j]), __uint_as_float(h[j]));
        }
      }
      // accumulator drained: hand it back to the MMA warp before the slower smem -> HBM part
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      mbar_arrive(acce0 + 8 * buf);
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      const int it0 = a.item_order ? __ldg(a.item_order + it) : it;   // partials stay indexed by the item itself
      float* out = a.partial + (size_t)it0 * NTILES * 16;
#pragma unroll
      for (int j = 0; j < EPI_TPT; ++j) {
        const int t = tid + j * kTcEpiThreads;
        if (t >= NTILES) break;
        const int I = eI[j], L = eL[j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 o;
          const int ri = 4 * I + i;
          o.x = 0.5f * (Xs[ri * XP + 4 * L + 0] + Xs[(4 * L + 0) * XP + ri]);
          o.y = 0.5f * (Xs[ri * XP + 4 * L + 1] + Xs[(4 * L + 1) * XP + ri]);
          o.z = 0.5f * (Xs[ri * XP + 4 * L + 2] + Xs[(4 * L + 2) * XP + ri]);
          o.w = 0.5f * (Xs[ri * XP + 4 * L + 3] + Xs[(4 * L + 3) * XP + ri]);
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
  return TcCfg<KT>::SMEM;
}

}  // namespace ycnr
