// ycnr_als.cu — C ABI of libycnr_als.so (include/ycnr_als.h): context, factor store,
// per-portion and bulk (row set) drivers, launch dispatch, profiling.
// Kernels: als_kernels.cuh (FFMA primal/dual), gram_tc.cuh (tcgen05 3xTF32 Gram),
// rmse_kernels.cuh.  sm_100a only; no CPU fallback anywhere in this file.
#include "ycnr_als.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <utility>
#include <vector>

#include "als_kernels.cuh"
#include "gram_tc.cuh"
#include "portion_kernels.cuh"
#include "recommend_kernels.cuh"
#include "ingest_kernels.cuh"
#include "rmse_kernels.cuh"

#ifndef YCNR_REDUCE_TPT
#define YCNR_REDUCE_TPT 3   // with YCNR_REDUCE_MIN_CTAS = 5: 8.4 ms against 9.0 ms for 2 tiles / 80 registers (MAL, two runs each)
#endif

namespace {

thread_local char g_err[1024] = "";
ycnr_ctx* g_default_ctx = nullptr;   // the most recently created live context (handle-less cpp_utils export)

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CU(expr)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)
#define OK(expr)          \
  do {                    \
    int rc__ = (expr);    \
    if (rc__) return rc__; \
  } while (0)

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    CU(cudaMalloc(&p, want));
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// ---- row classification -------------------------------------------------------------
constexpr int kDualBins = 24;  // one bin per tile-row count mt = ceil(n / 4), n <= 96
constexpr int kMaxChunks = 8;
constexpr int kSpreadBulkRows = 400000;
constexpr int kMinItemsPerChunk = 4 * 148;   // a persistent Gram launch needs a few items per SM

struct PlanCfg {
  int dual_max, split_cols, fused_max;
};

// Device-resident plan: one packed int32 array (row indices per kernel class) + offsets into it
struct DevPlan {
  int n_dual[kDualBins] = {0};
  size_t off_dual[kDualBins] = {0};
  int n_fused = 0, n_multi = 0, n_items = 0;
  size_t off_fused = 0, off_multi = 0, off_multi_first = 0, off_multi_n = 0, off_item_row = 0, off_item_off = 0;
  size_t off_item_order = 0;
  int64_t ratings_dual[kDualBins] = {0};
  int64_t ratings_fused = 0, ratings_multi = 0;
  size_t words = 0;
  // split rows in up to kMaxChunks groups of equal ratings: the Gram of group c+1 overlaps the
  // reduce+solve of group c on a second stream
  int n_chunks = 1;
  int chunk_row[kMaxChunks + 1] = {0};
  int chunk_item[kMaxChunks + 1] = {0};
  int64_t chunk_ratings[kMaxChunks] = {0};
};

// Two passes over the row lengths (len[r * stride]), no allocation, no per-row container work: this runs
// per portion on the e2e path, where the host has to stay ahead of the GPU.
//   n <= 0 ............ Q2 degenerate row (A = 0 upstream): skipped, see DESIGN.md
//   n <= dual_max ..... dual bin (n - 1) / 4
//   n <= fused_max .... fused primal row (FFMA path only)
//   otherwise ......... split row: ceil(n / split_cols) work items for the Gram kernels
void plan_count(const int32_t* len, int stride, int n_rows, const PlanCfg& cfg, DevPlan& p) {
  p = DevPlan();
  for (int r = 0; r < n_rows; ++r) {
    const int n = len[(size_t)r * stride];
    if (n <= 0) continue;
    if (n <= cfg.dual_max) {
      const int b = (n - 1) >> 2;
      p.n_dual[b]++;
      p.ratings_dual[b] += n;
    } else if (n <= cfg.split_cols && n <= cfg.fused_max) {
      p.n_fused++;
      p.ratings_fused += n;
    } else {
      p.n_multi++;
      p.n_items += (n + cfg.split_cols - 1) / cfg.split_cols;
      p.ratings_multi += n;
    }
  }
  size_t o = 0;
  for (int b = 0; b < kDualBins; ++b) { p.off_dual[b] = o; o += p.n_dual[b]; }
  p.off_fused = o;        o += p.n_fused;
  p.off_multi = o;        o += p.n_multi;
  p.off_multi_first = o;  o += p.n_multi;
  p.off_multi_n = o;      o += p.n_multi;
  p.off_item_row = o;     o += p.n_items;
  p.off_item_off = o;     o += p.n_items;
  p.off_item_order = o;   o += p.n_items;
  p.words = o;
}

void plan_fill(const int32_t* len, int stride, int n_rows, const PlanCfg& cfg, DevPlan& p, int32_t* out, int max_chunks) {
  size_t cur[kDualBins];
  for (int b = 0; b < kDualBins; ++b) cur[b] = p.off_dual[b];
  size_t cf = p.off_fused;
  int m = 0, it = 0;
  int32_t* multi = out + p.off_multi;
  int32_t* mfirst = out + p.off_multi_first;
  int32_t* mn = out + p.off_multi_n;
  int32_t* irow = out + p.off_item_row;
  int32_t* ioff = out + p.off_item_off;
  for (int r = 0; r < n_rows; ++r) {
    const int n = len[(size_t)r * stride];
    if (n <= 0) continue;
    if (n <= cfg.dual_max) {
      out[cur[(n - 1) >> 2]++] = r;
    } else if (n <= cfg.split_cols && n <= cfg.fused_max) {
      out[cf++] = r;
    } else {
      multi[m] = r;
      mfirst[m] = it;
      int cnt = 0;
      for (int off = 0; off < n; off += cfg.split_cols, ++cnt) {
        irow[it] = r;
        ioff[it] = off;
        ++it;
      }
      mn[m] = cnt;
      ++m;
    }
  }
  if (p.n_fused > 1) {   // longest rows first inside the launch (ties: list order)
    int32_t* f = out + p.off_fused;
    std::sort(f, f + p.n_fused, [&](int32_t x, int32_t y) {
      const int lx = len[(size_t)x * stride], ly = len[(size_t)y * stride];
      return lx != ly ? lx > ly : x < y;
    });
  }
  // processing order of the work items for the persistent Gram kernel: longest slices first (stable counting
  // sort on length / 32), see TcItemIter
  {
    int32_t* order = out + p.off_item_order;
    const int nb = cfg.split_cols / 32 + 2;
    std::vector<int32_t> cnt(nb + 1, 0);
    auto bucket = [&](int i) {
      const int n = len[(size_t)irow[i] * stride];
      const int l = std::min(cfg.split_cols, n - ioff[i]);
      return nb - 1 - std::min(nb - 1, l / 32);        // descending
    };
    for (int i = 0; i < p.n_items; ++i) cnt[bucket(i) + 1]++;
    for (int b = 0; b < nb; ++b) cnt[b + 1] += cnt[b];
    for (int i = 0; i < p.n_items; ++i) order[cnt[bucket(i)]++] = i;
  }
  // chunk cuts at equal cumulative ratings
  const int nm = p.n_multi;
  int chunks = (int)std::min<int64_t>(std::min(kMaxChunks, max_chunks), (int64_t)p.n_items / kMinItemsPerChunk);
  if (chunks < 2) chunks = 1;
  p.n_chunks = chunks;
  p.chunk_row[0] = 0;
  p.chunk_item[0] = 0;
  int64_t acc = 0, done = 0;
  int c = 1;
  for (int q = 0; q < nm && c < chunks; ++q) {
    acc += len[(size_t)multi[q] * stride];
    if (acc >= p.ratings_multi * c / chunks) {
      p.chunk_row[c] = q + 1;
      p.chunk_item[c] = mfirst[q] + mn[q];
      p.chunk_ratings[c - 1] = acc - done;
      done = acc;
      ++c;
    }
  }
  for (; c <= chunks; ++c) {   // last cut (and any cut the loop did not reach)
    p.chunk_row[c] = nm;
    p.chunk_item[c] = p.n_items;
    if (c - 1 < kMaxChunks) { p.chunk_ratings[c - 1] = p.ratings_multi - done; done = p.ratings_multi; }
  }
}

// Row-class thresholds from the options (shared by ycnr_create and the CPU-only planner diagnostics)
PlanCfg derive_plan_cfg(const ycnr_options& o, bool* use_tc_out) {
  const int k = o.factors_count;
  const bool tc = o.gram_path == YCNR_GRAM_TC3XTF32 ||
                  (o.gram_path == YCNR_GRAM_AUTO && (k & 3) == 0 && k <= 256 && k >= 16);
  // dual / primal crossover.  Measured on B200 at k = 100 (ns per row, whole GPU): dual rows of 22, 23, 24 tile rows
  // cost 56, 78, 82; a primal row just above the crossover 59 (tensor-core Gram 0.107 per rating + 46 for the k x k
  // solve) — so rows above 88 ratings go primal there.  Wider systems keep 96 (their solve grows with k^3: 650 ns at
  // k = 256 against 150 for the largest dual row), FFMA-path systems too (their primal Gram is 4x slower).
  const int dual_cap = (tc && k <= 112) ? 88 : 96;
  const int dflt_dual = std::min(dual_cap, std::max(0, ((k - 1) / 4) * 4));
  PlanCfg cfg;
  cfg.dual_max = o.dual_max_cols < 0 ? dflt_dual : std::min(96, o.dual_max_cols);
  cfg.split_cols = o.split_cols > 0 ? std::max(o.split_cols, ycnr::kStageRows) : 4096;
  if (cfg.dual_max > cfg.split_cols) cfg.dual_max = cfg.split_cols;
  // AUTO: tensor cores whenever the rhs column fits the M = 128 accumulator and rows are 16-byte aligned
  const bool use_tc = o.gram_path == YCNR_GRAM_TC3XTF32 ||
                      (o.gram_path == YCNR_GRAM_AUTO && (k & 3) == 0 && k <= 256 && k >= 16);
  cfg.fused_max = use_tc ? (o.tc_min_cols > 0 ? o.tc_min_cols - 1 : 0) : cfg.split_cols;
  if (use_tc_out) *use_tc_out = use_tc;
  return cfg;
}

struct RowSet {
  bool used = false;
  int step_type = 0;
  int n_rows = 0;
  int64_t span = 0, nnz = 0;
  int n_portions = 0;
  DevBuf rows;      // row_ids | row_len | portion_first (int32) then row_start (int64), packed
  DevBuf ratings;   // indx (int32) | vals (float)
  DevBuf plan;      // DevPlan words
  DevBuf sums;      // rmse: row_sums [R][3] | portion_sums [P][4]
  DevBuf order;     // rmse: entries longest first
  DevBuf erows;     // rmse: work entries (rows cut into <= kRmseChunk ratings): start i64 | ids | len | efirst[P+1]
  RowsView eview{};
  const int32_t* d_efirst = nullptr;
  int n_entries = 0;
  RowsView view{};
  const int32_t* d_portion_first = nullptr;
  DevPlan dplan;
  // RMSE sets: per-portion sums {rSumDiff2, rCnt, rSum, sum of ratings} of the last pass, its shift and the
  // versions of the two factor matrices it saw.  While the factors have not changed, a pass with another shift
  // is derived on the host (the third pass of an iteration, EmfLord.js:898, costs nothing).
  double* h_sums = nullptr;          // page-locked [P][4]
  cudaEvent_t sums_ready = nullptr;
  bool pending = false, cached = false;
  double cache_shift = 0.0;
  uint64_t cache_ver[2] = {0, 0};
};

struct Slot {  // staging for the per-portion path
  void* host = nullptr;
  size_t host_cap = 0;
  void* host2 = nullptr;      // batches of small portions: row arrays start | ids | len
  size_t host2_cap = 0;
  void* host3 = nullptr;      // batches: portion firsts + launch plan (host -> device) / RMSE sums (device -> host)
  size_t host3_cap = 0;
  int inflight = -1;          // index into ycnr_ctx::rmse_inflight of the RMSE batch whose sums land in host3
  DevBuf dev;
  cudaEvent_t done = nullptr;
  cudaEvent_t solved = nullptr;   // kernels of the portion finished: its rows may be copied back
  bool pending = false;
};
constexpr int kSlots = 4;

// ---- batches of small portions ------------------------------------------------------------------------
// The reference hands the worker ~10 000 ratings per message (ratingsInPortionForAls, EmfBase.js:97-103): about 150
// user rows, microseconds of GPU work.  Portions below kBatchDirectRatings are therefore QUEUED: the call checks the
// header, appends the rows to the open batch (ratings of page-locked, address-adjacent cache buffers are only
// noted as DMA segments, others are copied into the slot) and returns; the batch goes to the device as ONE launch
// group when it holds kBatchFlushRatings ratings, when a large portion arrives, or when the step ends.
constexpr int64_t kBatchFlushRatings = (int64_t)4 << 20;
constexpr int64_t kBatchDirectRatings = (int64_t)1 << 20;   // portions in unregistered buffers: staged in the slot up to this size
constexpr int64_t kBatchMaxRatings = (int64_t)32 << 20;     // portions in page-locked (cached) buffers: DMA'd in place up to this size

struct BatchSeg {
  const int32_t* indx;
  const float* vals;
  int64_t n;
  int64_t dev_off;     // first rating of the segment in the batch's device arrays
  int64_t stage_off;   // staged segments: offset inside the slot's staging arrays
  bool direct;         // page-locked caller memory: DMA'd from where it lies
};

struct Batch {
  int kind = 0;        // 0 = closed, 1 = ALS, 2 = RMSE
  Slot* slot = nullptr;
  int32_t* ids = nullptr;                            // row arrays, in the slot's page-locked host2
  int32_t* len = nullptr;
  int64_t* start = nullptr;
  size_t n_rows = 0, cap_rows = 0;
  std::vector<int32_t> pfirst;                       // RMSE: first entry of every portion
  std::vector<int64_t> tags;                         // RMSE: caller's tag per portion
  std::vector<ycnr_portion_info> infos;              // RMSE: rows_from / rows_cnt / ratings per portion
  std::vector<BatchSeg> segs;
  std::vector<std::pair<int32_t, int32_t>> ranges;   // ALS: [first, last] row id per portion
  int64_t ratings = 0, staged = 0;
};

struct RmseInflight {
  Slot* slot = nullptr;
  cudaEvent_t done = nullptr;
  std::vector<int64_t> tags;
  std::vector<ycnr_portion_info> infos;
  const double* h_sums = nullptr;                    // [P][4] in the slot's host3
  bool live = false;
};

// ---- host worker pool for the multi-portion entry points -------------------------------------------------
// ycnr_als_portions / ycnr_rmse_portions_async see all portions of a half-step (an RMSE pass) at once: their
// headers (1.75 M rows per pass on MAL: 4.6 ms single-threaded, more than the pass's kernels and DMA together)
// are scanned and written into the batch arrays by a few threads.  The caller takes part; workers sleep on a
// condition variable between regions.
struct WorkPool {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  const std::function<void(int)>* fn = nullptr;
  int n_tasks = 0, active = 0;
  std::atomic<int> next{0};
  uint64_t gen = 0;
  bool stop = false;
  void start(int workers) {
    for (int i = 0; i < workers; ++i)
      th.emplace_back([this] {
        uint64_t seen = 0;
        for (;;) {
          const std::function<void(int)>* f;
          int n;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv_go.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            f = fn;
            n = n_tasks;
          }
          for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) (*f)(i);
          std::lock_guard<std::mutex> lk(mu);
          if (--active == 0) cv_done.notify_one();
        }
      });
  }
  void run(int n, const std::function<void(int)>& f) {
    if (th.empty() || n < 2) {
      for (int i = 0; i < n; ++i) f(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      fn = &f;
      n_tasks = n;
      next.store(0);
      active = (int)th.size();
      ++gen;
    }
    cv_go.notify_all();
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i);
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return active == 0; });
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv_go.notify_all();
    for (auto& t : th) t.join();
    th.clear();
  }
};

// What a scan of a portion header yields without writing anything (multi-portion entry points)
struct PortionPre {
  int64_t entries = 0;    // rows (ALS) / work entries of at most kRmseChunk ratings (RMSE)
  int64_t ratings = 0;
  int bad = 0;
};
// A header whose rows still have to be written into the open batch's arrays
struct FillTask {
  int kind;
  const int32_t* rows;
  size_t r0;              // first entry in the batch arrays
  int64_t base;           // first rating in the batch's device arrays
};

struct ProfRec {
  int cls;
  cudaEvent_t a, b;
  int sub;   // dual bin (tile-row count - 1) or -1
};

}  // namespace

struct ycnr_ctx {
  ycnr_options opts{};
  int k = 0;
  int dual_max = 0;
  int split_cols = 0;
  int fused_max = 0;       // rows longer than this go through the partial (+reduce) kernels
  bool use_tc = false;     // tcgen05 3xTF32 Gram for the partial kernels
  bool tc_tested = false;  // tc_self_test ran on this context's device
  CUtensorMap tmap[2];     // TMA descriptors of the two replicas (box 32 columns x 1 row, SWIZZLE_128B_ATOM_32B)
  bool tma_ok[2] = {false, false};
  bool no_tma = false;     // YCNR_NO_TMA=1: keep the cp.async gather (A/B measurements)
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // H2D of portion inputs, overlaps the previous portion's kernels
  cudaEvent_t copied = nullptr;
  // per-portion path: the dual bins of a portion are small launches (about one wave each at 8 M ratings);
  // spread over these streams they run concurrently and fill each other's tail waves
#ifndef YCNR_BIN_STREAMS
#define YCNR_BIN_STREAMS 3
#endif
  static constexpr int kBinStreams = YCNR_BIN_STREAMS;
  cudaStream_t bin_stream[kBinStreams] = {nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[kBinStreams] = {nullptr};
  cudaStream_t d2h_stream = nullptr;    // solved rows -> host factor segment, overlaps the next portion
  bool d2h_pending = false;
  cudaStream_t aux_stream = nullptr;    // reduce+solve of split rows, overlapping the next Gram chunk
  cudaEvent_t chunk_ev[kMaxChunks + 2] = {nullptr};
  bool aux_pending = false;
  std::vector<std::pair<const char*, size_t>> pinned;  // ycnr_host_register regions
  float* d_fac[2] = {nullptr, nullptr};
  float* h_fac[2] = {nullptr, nullptr};
  bool h_registered[2] = {false, false};
  bool device_current[2] = {false, false};
  int64_t fac_rows[2] = {0, 0};
  uint64_t fac_version[2] = {1, 1};   // bumped whenever a replica may have changed (cached RMSE sums go stale)
  std::vector<void*> peers[2];
  std::vector<RowSet> rowsets;
  // device-resident ratings table (ycnr_table_upload): user_ptr i64[users+1] | item i32 | rating f32 | dt i8 | elem_user i32
  struct {
    DevBuf buf;
    int64_t nnz = 0;
    const int64_t* user_ptr = nullptr;
    const int32_t* item = nullptr;
    const float* rating = nullptr;
    const int8_t* dt = nullptr;
    const int32_t* elem_user = nullptr;
    bool loaded = false;
  } table;
  DevBuf ingest_tmp;
  DevBuf partial;      // split-row tile partials
  DevBuf gather_tmp;   // ycnr_s_als_build_sub_fixed_facts
  // per-portion path: column ids are checked on the device (validate_cols_kernel); a raised flag makes the
  // compute kernels of the step leave early and the step end with an error
  int32_t* d_bad = nullptr;
  int32_t* h_bad = nullptr;   // page-locked
  Slot slots[kSlots];
  int next_slot = 0;
  Batch batch;
  size_t max_host3 = 0;   // largest plan block any slot has needed: every slot is grown to it at its next use
  size_t max_slot_dev = 0;
  size_t max_batch_rows = 0;
  int flushes_in_pass = 0;   // batches flushed since the step / RMSE pass began: the first ones are smaller (see batch_add)
  int64_t batch_flush_ratings = kBatchFlushRatings;   // YCNR_BATCH_FLUSH overrides (tests exercise many flushes)
  std::vector<RmseInflight> rmse_inflight;
  std::vector<int> rmse_order;                                    // live entries of rmse_inflight, oldest first
  std::vector<std::pair<int64_t, ycnr_portion_info>> rmse_done;   // completed, not yet polled (FIFO)
  size_t rmse_done_head = 0;
  // Per-portion sums {rSumDiff2, rCnt, rSum, sum of ratings} of the previous RMSE pass, by tag (portionNo): the
  // third pass of an iteration (EmfLord.js:898: same test portions, shift applied, factors untouched) is answered
  // from them — sum (r-p-d)^2 = sum (r-p)^2 - 2d (sum r - sum p) + n d^2 — without touching the device again.
  struct PortionSums {
    int32_t rows_from, rows_cnt;
    int64_t ratings;
    double d2, cnt, pred, rat;
  };
  struct RmsePass {
    int step_type = -1;
    double shift = 0.0;
    uint64_t ver[2] = {0, 0};
    std::unordered_map<int64_t, PortionSums> sums;   // by tag
  } rmse_prev, rmse_cur;
  // per-portion step state
  int step_type = -1;
  double rmse_shift = 0.0;
  std::vector<std::pair<int32_t, int32_t>> solved_ranges;  // [first,last] row ids per portion
  // YCNR_TRACE=1: host-side time of the per-portion path, printed per step to stderr
  bool trace = false;
  // Bulk half-steps: dual bins on the bin streams too?  -1 = auto (row sets below kSpreadBulkRows rows: the bins
  // are then a few waves each; measured on B200: ML-1M shape 1.30 -> 1.04 ms per iteration, MAL 53.2 -> 52.3 ms),
  // YCNR_SPREAD_BULK=0/1 forces it.  Overlapped launches make the per-class event times overlap as well.
  int spread_bulk = -1;
  bool spread_batch = true;   // batches of small portions: dual bins on the bin streams (YCNR_SPREAD_BATCH=0: in stream order)
  WorkPool pool;
  std::vector<FillTask> fill_tasks;   // deferred header writes of the open batch (multi-portion entry points)
  double t_parse = 0, t_slot_wait = 0, t_copy_issue = 0, t_launch = 0, t_add = 0, t_rmse_calls = 0, t_fill = 0, t_scan = 0;
  int t_portions = 0;
  // kernels whose dynamic shared-memory limit was raised on this context's device (function attributes are
  // per device: a process-wide flag would skip the second device of a process)
  std::vector<std::pair<const void*, size_t>> smem_configured;
  // profiling
  std::vector<ProfRec> prof_open;
  std::vector<cudaEvent_t> ev_pool;
  ycnr_profile prof{};
  double dual_bin_ms[kDualBins] = {0};
  int64_t dual_bin_rows[kDualBins] = {0};
};

namespace {

int set_device(ycnr_ctx* c) {
  CU(cudaSetDevice(c->opts.device));
  return 0;
}

DstList make_dst(ycnr_ctx* c, int which) {
  DstList d{};
  d.p[0] = c->d_fac[which];
  d.n = 1;
  for (void* p : c->peers[which])
    if (d.n < YCNR_MAX_DST) d.p[d.n++] = (float*)p;
  return d;
}

int ensure_dynamic_smem(ycnr_ctx* c, const void* func, size_t bytes) {
  for (auto& e : c->smem_configured)
    if (e.first == func) {
      if (e.second >= bytes) return 0;
      CU(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      e.second = bytes;
      return 0;
    }
  CU(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  c->smem_configured.emplace_back(func, bytes);
  return 0;
}

struct ProfScope {
  ycnr_ctx* c;
  int cls;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  int sub = -1;
  ProfScope(ycnr_ctx* ctx, int cls_, int64_t rows, int64_t ratings, cudaStream_t stream = nullptr, int sub_ = -1)
      : c(ctx), cls(cls_), st(stream ? stream : ctx->stream), sub(sub_) {
    if (sub >= 0) c->dual_bin_rows[sub] += rows;
    c->prof.launches[cls]++;
    c->prof.rows[cls] += rows;
    c->prof.ratings[cls] += ratings;
    c->prof.total_launches++;
    if (!c->opts.profile) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    c->prof_open.push_back({cls, a, b, sub});
  }
};

// ---- kernel dispatch -----------------------------------------------------------------
// gram_tc_kernel feeds the RAW fp32 bits of the gathered rows as the "head" operand and relies on kind::tf32
// ignoring the low 13 mantissa bits (the PTX text only says the operand "is" tf32).  Once per context: one
// 64-rating slice with full-mantissa values through the kernel as it is and with explicitly masked heads
// (variant 16); if the partials differ by a single bit the context keeps the explicit masking on.
template <int KT>
int tc_self_test(ycnr_ctx* c) {
  using namespace ycnr;
  constexpr int NTILES = KT * (KT + 1) / 2 + KT;
  constexpr int NR = 64;
  const int k = c->k;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_fix = 0, o_idx = al((size_t)NR * k * 4), o_val = al(o_idx + NR * 4), o_rs = al(o_val + NR * 4);
  const size_t o_ri = al(o_rs + 8), o_rl = al(o_ri + 4), o_ir = al(o_rl + 4), o_io = al(o_ir + 4);
  const size_t o_p0 = al(o_io + 4), o_p1 = al(o_p0 + (size_t)NTILES * 64), total = al(o_p1 + (size_t)NTILES * 64);
  std::vector<char> h(o_p0, 0);
  float* fx = (float*)(h.data() + o_fix);
  uint32_t lcg = 12345u;
  for (int i = 0; i < NR * k; ++i) {
    lcg = lcg * 1664525u + 1013904223u;
    fx[i] = ((int)(lcg >> 8) - (1 << 23)) * (1.0f / (1 << 23)) * 0.7f;     // all 24 mantissa bits in use
  }
  for (int i = 0; i < NR; ++i) {
    ((int32_t*)(h.data() + o_idx))[i] = i;
    ((float*)(h.data() + o_val))[i] = (float)(1 + i % 10);
  }
  *(int64_t*)(h.data() + o_rs) = 0;
  *(int32_t*)(h.data() + o_ri) = 0;
  *(int32_t*)(h.data() + o_rl) = NR;
  *(int32_t*)(h.data() + o_ir) = 0;
  *(int32_t*)(h.data() + o_io) = 0;
  DevBuf buf;
  OK(buf.ensure(total));
  char* d = (char*)buf.p;
  CU(cudaMemcpyAsync(d, h.data(), o_p0, cudaMemcpyHostToDevice, c->stream));
  GramTcArgs t{};
  t.rows.row_start = (const int64_t*)(d + o_rs);
  t.rows.row_ids = (const int32_t*)(d + o_ri);
  t.rows.row_len = (const int32_t*)(d + o_rl);
  t.rows.indx = (const int32_t*)(d + o_idx);
  t.rows.vals = (const float*)(d + o_val);
  t.fixed = (const float*)(d + o_fix);
  t.k = k;
  t.item_row = (const int32_t*)(d + o_ir);
  t.item_off = (const int32_t*)(d + o_io);
  t.n_items = 1;
  t.split_cols = 4096;
  t.chunks_a = 1 << 20;
  t.use_tma = 0;
  const size_t smem = gram_tc_smem_bytes<KT>();
  OK(ensure_dynamic_smem(c, reinterpret_cast<const void*>(&gram_tc_kernel<KT>), smem));
  for (int v = 0; v < 2; ++v) {
    t.partial = (float*)(d + (v ? o_p1 : o_p0));
    t.variant = v ? 16u : 0u;
    gram_tc_kernel<KT><<<1, TcCfg<KT>::THREADS, smem, c->stream>>>(t);
    CU(cudaGetLastError());
  }
  c->prof.total_launches += 2;
  std::vector<uint32_t> out((size_t)NTILES * 32);
  CU(cudaMemcpyAsync(out.data(), d + o_p0, (size_t)NTILES * 64, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(out.data() + (size_t)NTILES * 16, d + o_p1, (size_t)NTILES * 64, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  buf.release();
  if (memcmp(out.data(), out.data() + (size_t)NTILES * 16, (size_t)NTILES * 64) != 0) {
    c->opts.tc_variant |= 16;
    fprintf(stderr, "[ycnr] kind::tf32 does not truncate raw fp32 operands on this device: explicit head masking enabled\n");
  }
  return 0;
}

struct TcColMap {
  int col_a = 0, col_b = 0, chunks_a = 1 << 20;   // default: one pass over columns [0, k)
};

template <int KT>
int launch_gram_tc(ycnr_ctx* c, const ycnr::PrimalArgs& pa, int item_from, int n_items, int64_t ratings,
                   float* partial_base = nullptr, TcColMap cm = TcColMap()) {
  using namespace ycnr;
  if constexpr (4 * KT + 4 > 128) {
    return fail("tcgen05 Gram path supports factorsCount <= 124");
  } else {
    if (!c->tc_tested) {
      c->tc_tested = true;
      OK(tc_self_test<KT>(c));
    }
    GramTcArgs t{};
    t.rows = pa.rows;
    t.fixed = pa.fixed;
    t.k = pa.k;
    constexpr int NTILES = KT * (KT + 1) / 2 + KT;
    t.item_row = pa.item_row + item_from;
    t.item_off = pa.item_off + item_from;
    t.item_order = (item_from == 0 && n_items == pa.n_items_total) ? pa.item_order : nullptr;   // chunked launches: natural order
    t.n_items = n_items;
    t.split_cols = pa.split_cols;
    t.partial = partial_base ? partial_base : pa.partial + (size_t)item_from * NTILES * 16;
    t.col_a = cm.col_a;
    t.col_b = cm.col_b;
    t.chunks_a = cm.chunks_a;
    t.use_tma = 0;
    t.fixed_rows = pa.fixed_rows;
    if (pa.tmap && cm.chunks_a >= KT && !c->no_tma) {   // one-pass systems gather with the copy engine
      memcpy(&t.tmap, pa.tmap, sizeof(CUtensorMap));
      t.use_tma = 1;
    }
    t.variant = (uint32_t)c->opts.tc_variant;
    t.prefetch = pa.fixed_bytes > ((size_t)48 << 20) ? 1 : 0;
    const size_t smem = gram_tc_smem_bytes<KT>();
    OK(ensure_dynamic_smem(c, reinterpret_cast<const void*>(&gram_tc_kernel<KT>), smem));
    const int grid = std::min(n_items, c->num_sms);
    ProfScope ps(c, YCNR_K_GRAM_TC, n_items, ratings);
    gram_tc_kernel<KT><<<grid, TcCfg<KT>::THREADS, smem, c->stream>>>(t);
    CU(cudaGetLastError());
    return 0;
  }
}

template <int KT, int NT>
int launch_primal_kt(ycnr_ctx* c, const ycnr::PrimalArgs& base, const DevPlan& p, const int32_t* plan_base) {
  using namespace ycnr;
  constexpr int NTILES = KT * (KT + 1) / 2 + KT;
  if (p.n_fused > 0) {
    PrimalArgs a = base;
    a.work = plan_base + p.off_fused;
    ProfScope ps(c, YCNR_K_PRIMAL_FUSED, p.n_fused, p.ratings_fused);
    als_primal_kernel<KT, NT, 1, MODE_FUSED><<<p.n_fused, NT, 0, c->stream>>>(a);
  }
  if (p.n_items > 0) {
    OK(c->partial.ensure((size_t)p.n_items * NTILES * 16 * sizeof(float)));
    PrimalArgs a = base;
    a.item_row = plan_base + p.off_item_row;
    a.item_off = plan_base + p.off_item_off;
    a.item_order = plan_base + p.off_item_order;
    a.n_items_total = p.n_items;
    a.partial = (float*)c->partial.p;
    a.work = plan_base + p.off_multi;
    a.row_first_item = plan_base + p.off_multi_first;
    a.row_n_items = plan_base + p.off_multi_n;
    if (c->use_tc && p.n_chunks > 1) {
      // optional (solve_chunks > 1): reduce+solve of chunk i on the aux stream under the Gram of chunk i+1.
      // Measured on B200 (MAL): no gain — 56.5 ms per iteration with the solve overlapped against 55.7 ms
      // in stream order; both kernels compete for the same issue slots — so it is off by default.
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        const int i0 = p.chunk_item[ch], i1 = p.chunk_item[ch + 1];
        const int m0 = p.chunk_row[ch], m1 = p.chunk_row[ch + 1];
        if (i1 > i0) OK((launch_gram_tc<KT>(c, a, i0, i1 - i0, p.chunk_ratings[ch])));
        CU(cudaEventRecord(c->chunk_ev[ch], c->stream));
        CU(cudaStreamWaitEvent(c->aux_stream, c->chunk_ev[ch], 0));
        if (m1 > m0) {
          PrimalArgs r = a;
          r.work = a.work + m0;
          r.row_first_item = a.row_first_item + m0;
          r.row_n_items = a.row_n_items + m0;
          ProfScope ps(c, YCNR_K_REDUCE_SOLVE, m1 - m0, 0, c->aux_stream);
          als_primal_kernel<KT, NT, 1, MODE_REDUCE><<<m1 - m0, NT, 0, c->aux_stream>>>(r);
        }
      }
      c->aux_pending = true;
    } else {
      if (c->use_tc) {
        OK((launch_gram_tc<KT>(c, a, 0, p.n_items, p.ratings_multi)));
      } else {
        ProfScope ps(c, YCNR_K_GRAM_PARTIAL, p.n_items, p.ratings_multi);
        als_primal_kernel<KT, NT, 1, MODE_PARTIAL><<<p.n_items, NT, 0, c->stream>>>(a);
      }
      ProfScope ps(c, YCNR_K_REDUCE_SOLVE, p.n_multi, 0);
      constexpr int RT = YCNR_REDUCE_TPT;                                   // tiles per thread of the solve
      constexpr int RNT = (((NTILES + RT - 1) / RT) + 31) & ~31;
      als_primal_kernel<KT, RNT, RT, MODE_REDUCE><<<p.n_multi, RNT, 0, c->stream>>>(a);
    }
  }
  CU(cudaGetLastError());
  return 0;
}

// ---- k > 124: one tensor-core pass per pair of column blocks ----------------------------------------
// The tcgen05 Gram holds a system of at most 124 columns (+ the rhs) in its M = 128 accumulator.  Wider systems
// are cut into nb blocks of w <= 60 columns; the pass for the pair (a < b) gathers only the columns of the two
// blocks and builds the Gram of the 2 w-column system [a | b] with the kernel as it is — its diagonal quarters
// are blocks (a,a) and (b,b) of the full matrix, the off-diagonal quarter is block (b,a), its rhs row holds
// b_a and b_b.  nb (nb - 1) / 2 passes; als_solve_blocks_kernel picks every tile from the pass that holds it.
struct BlockCfg {
  int nb, w;
};
BlockCfg block_cfg(int k) {
  const int nb = (k + 59) / 60;
  for (int w : {44, 52, 60})
    if (nb * w >= k) return {nb, w};
  return {nb, 60};
}
constexpr size_t kMaxBlockPartialBytes = (size_t)12 << 30;

template <int KT, int NT, int TPT>
int launch_solve_blocks(ycnr_ctx* c, const ycnr::SolveBlocksArgs& a, int rows) {
  using namespace ycnr;
  ProfScope ps(c, YCNR_K_REDUCE_SOLVE, rows, 0);
  als_solve_blocks_kernel<KT, NT, TPT><<<rows, NT, 0, c->stream>>>(a);
  CU(cudaGetLastError());
  return 0;
}

template <int KTV>
int launch_primal_blocks(ycnr_ctx* c, const ycnr::PrimalArgs& base, const DevPlan& p, const int32_t* plan_base,
                         BlockCfg bc) {
  using namespace ycnr;
  if (p.n_items <= 0) return 0;
  constexpr int NTV = KTV * (KTV + 1) / 2 + KTV;
  const int n_pass = bc.nb * (bc.nb - 1) / 2;
  int max_items = 0;
  for (int ch = 0; ch < p.n_chunks; ++ch) max_items = std::max(max_items, p.chunk_item[ch + 1] - p.chunk_item[ch]);
  const size_t pass_stride = (size_t)max_items * NTV * 16;
  const size_t bytes = pass_stride * n_pass * sizeof(float);
  if (bytes > kMaxBlockPartialBytes + (kMaxBlockPartialBytes >> 1))
    return fail("factorsCount %d: %zu bytes of tile partials for %d rows in one group (internal chunking limit)", c->k, bytes, p.n_multi);
  OK(c->partial.ensure(bytes));
  PrimalArgs a = base;
  a.item_row = plan_base + p.off_item_row;
  a.item_off = plan_base + p.off_item_off;
  a.item_order = nullptr;
  a.n_items_total = p.n_items;
  a.partial = (float*)c->partial.p;
  for (int ch = 0; ch < p.n_chunks; ++ch) {
    const int i0 = p.chunk_item[ch], i1 = p.chunk_item[ch + 1];
    const int m0 = p.chunk_row[ch], m1 = p.chunk_row[ch + 1];
    if (i1 <= i0 || m1 <= m0) continue;
    int pass = 0;
    for (int ba = 0; ba < bc.nb; ++ba)
      for (int bb = ba + 1; bb < bc.nb; ++bb, ++pass) {
        TcColMap cm;
        cm.col_a = ba * bc.w;
        cm.col_b = bb * bc.w;
        cm.chunks_a = bc.w / 4;
        OK((launch_gram_tc<KTV>(c, a, i0, i1 - i0, p.chunk_ratings[ch], (float*)c->partial.p + pass * pass_stride, cm)));
      }
    SolveBlocksArgs sa{};
    sa.rows = base.rows;
    sa.k = c->k;
    sa.lambda = base.lambda;
    sa.dst = base.dst;
    sa.work = plan_base + p.off_multi + m0;
    sa.row_first_item = plan_base + p.off_multi_first + m0;
    sa.row_n_items = plan_base + p.off_multi_n + m0;
    sa.partial = (const float*)c->partial.p;
    sa.pass_stride = pass_stride;
    sa.item_base = i0;
    sa.nb = bc.nb;
    sa.wt = bc.w / 4;
    const int kt = (c->k + 3) / 4;
    if (kt <= 32) OK((launch_solve_blocks<32, 192, 3>(c, sa, m1 - m0)));
    else if (kt <= 40) OK((launch_solve_blocks<40, 288, 3>(c, sa, m1 - m0)));
    else if (kt <= 48) OK((launch_solve_blocks<48, 320, 4>(c, sa, m1 - m0)));
    else if (kt <= 56) OK((launch_solve_blocks<56, 416, 4>(c, sa, m1 - m0)));
    else OK((launch_solve_blocks<64, 480, 5>(c, sa, m1 - m0)));
  }
  return 0;
}

int launch_primal(ycnr_ctx* c, const ycnr::PrimalArgs& base, const DevPlan& p, const int32_t* plan_base) {
  const int k = c->k;
  if (c->use_tc && k > 124) {
    const BlockCfg bc = block_cfg(k);
    if (bc.w == 44) return launch_primal_blocks<22>(c, base, p, plan_base, bc);
    if (bc.w == 52) return launch_primal_blocks<26>(c, base, p, plan_base, bc);
    return launch_primal_blocks<30>(c, base, p, plan_base, bc);
  }
  if (k <= 20) return launch_primal_kt<5, 32>(c, base, p, plan_base);
  if (k <= 32) return launch_primal_kt<8, 64>(c, base, p, plan_base);
  if (k <= 64) return launch_primal_kt<16, 160>(c, base, p, plan_base);
  if (k <= 100) return launch_primal_kt<25, 352>(c, base, p, plan_base);
  if (k <= 124) return launch_primal_kt<31, 544>(c, base, p, plan_base);  // largest KT whose rhs column fits M = 128
  if (k <= 128) return launch_primal_kt<32, 576>(c, base, p, plan_base);
  return fail("factorsCount %d > 128 needs the tensor-core Gram (factorsCount %% 4 == 0, gramPath auto or tc)", k);
}

// threads per CTA for systems of mt tile rows: room for a 2..8-way K-split of the Gram sweep while
// the systems are small (mt <= 8), exactly the tile set (rounded to warps) above that
// (measured on B200, MAL byUser: the K-split pays for mt <= 8 — 7.2 ms against 8.5 ms for those rows —
//  and costs occupancy above that)
constexpr int dual_nt(int mt) {
  if (mt <= 4) return 32;
  if (mt <= 6) return 64;
  if (mt <= 8) return 96;
  return (ycnr::dual_ntl(mt) + 31) & ~31;
}

template <int MT_MAX, int NT>
int launch_dual_bin(ycnr_ctx* c, const ycnr::DualArgs& base, int count, int64_t ratings, const int32_t* work,
                    cudaStream_t st) {
  using namespace ycnr;
  if (count <= 0) return 0;
  DualArgs a = base;
  a.work = work;
  int red = 0;
  for (int mt = 1; mt <= MT_MAX; ++mt) red = std::max(red, dual_red_floats(mt, NT));
  a.bar_off = 4 * MT_MAX * a.pitch + (MT_MAX + 1) * 16 + 16 * MT_MAX + 8 * MT_MAX + red;   // multiple of 4 floats
  const size_t smem = (size_t)a.bar_off * sizeof(float) + 16;
  if (smem > 48 * 1024) OK(ensure_dynamic_smem(c, reinterpret_cast<const void*>(&als_dual_kernel<MT_MAX, NT>), smem));
  ProfScope ps(c, YCNR_K_DUAL_FUSED, count, ratings, st, MT_MAX - 1);
  als_dual_kernel<MT_MAX, NT><<<count, NT, smem, st>>>(a);
  CU(cudaGetLastError());
  return 0;
}

// Bins of mt >= YCNR_DUAL2_MIN_MT tile rows run the several-tiles-per-thread kernel: CTA width = the tile set
// divided by YCNR_DUAL_TPT, rounded up to whole warps.
#ifndef YCNR_DUAL_TPT
#define YCNR_DUAL_TPT 3
#endif
#ifndef YCNR_DUAL2_MIN_MT
#define YCNR_DUAL2_MIN_MT 1   // measured on B200 (MAL, k = 100): with its in-warp K-split the one-warp kernel wins everywhere
#endif
constexpr int dual2_nt(int mt) { return ((ycnr::dual_ntl(mt) + YCNR_DUAL_TPT - 1) / YCNR_DUAL_TPT + 31) & ~31; }
constexpr int dual2_tpt(int mt) { return (ycnr::dual_ntl(mt) + dual2_nt(mt) - 1) / dual2_nt(mt); }

template <int MT>
int launch_dual_bin2(ycnr_ctx* c, const ycnr::DualArgs& base, int count, int64_t ratings, const int32_t* work,
                     cudaStream_t st) {
  using namespace ycnr;
  if (count <= 0) return 0;
  constexpr int NT = dual2_nt(MT), TPT = dual2_tpt(MT);
  DualArgs a = base;
  a.work = work;
  const size_t smem = ((size_t)4 * MT * a.pitch + (MT + 1) * 16 + 16 * MT + 8 * MT) * sizeof(float) + 16;   // + gather mbarrier
  if (smem > 48 * 1024) OK(ensure_dynamic_smem(c, reinterpret_cast<const void*>(&als_dual_tpt_kernel<MT, NT, TPT>), smem));
  ProfScope ps(c, YCNR_K_DUAL_FUSED, count, ratings, st, MT - 1);
  als_dual_tpt_kernel<MT, NT, TPT><<<count, NT, smem, st>>>(a);
  CU(cudaGetLastError());
  return 0;
}

template <int MT>
int launch_dual_bin_any(ycnr_ctx* c, const ycnr::DualArgs& d, int count, int64_t ratings, const int32_t* work,
                        cudaStream_t st) {
  if constexpr (MT >= YCNR_DUAL2_MIN_MT) return launch_dual_bin2<MT>(c, d, count, ratings, work, st);
  else return launch_dual_bin<MT, dual_nt(MT)>(c, d, count, ratings, work, st);
}

template <int... B>
int launch_dual_bins(ycnr_ctx* c, const ycnr::DualArgs& d, const DevPlan& p, const int32_t* plan_base, bool spread,
                     std::integer_sequence<int, B...>) {
  int rc = 0;
  ((rc = rc ? rc
            : launch_dual_bin_any<B + 1>(c, d, p.n_dual[B], p.ratings_dual[B], plan_base + p.off_dual[B],
                                         spread ? c->bin_stream[B % ycnr_ctx::kBinStreams] : c->stream)),
   ...);
  return rc;
}

int dual_pitch(int k) {
  int p = (k + 3) & ~3;
  if (((p >> 2) & 1) == 0) p += 4;
  return p;
}

// One ALS half-step over a device-resident row list + plan.
int run_als(ycnr_ctx* c, int step_type, const RowsView& view, const DevPlan& p, const int32_t* plan_base,
            bool spread = false) {
  const int solved = step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
  const int fixed = 1 - solved;
  const double lambda = step_type == YCNR_BY_USER ? c->opts.user_fact_reg : c->opts.item_fact_reg;
  ycnr::DualArgs d{};
  d.rows = view;
  d.fixed = c->d_fac[fixed];
  d.k = c->k;
  d.pitch = dual_pitch(c->k);
  d.lambda = lambda;
  d.dst = make_dst(c, solved);
  ycnr::PrimalArgs a{};
  a.rows = view;
  a.fixed = c->d_fac[fixed];
  a.fixed_bytes = (size_t)c->fac_rows[fixed] * c->k * sizeof(float);
  a.tmap = c->tma_ok[fixed] ? &c->tmap[fixed] : nullptr;
  a.fixed_rows = (int)c->fac_rows[fixed];
  a.k = c->k;
  a.lambda = lambda;
  a.dst = d.dst;
  a.split_cols = c->split_cols;
  // long rows first: their reduce+solve chunks run on the aux stream under the dual kernels too
  CU(cudaEventRecord(c->chunk_ev[kMaxChunks], c->stream));
  CU(cudaStreamWaitEvent(c->aux_stream, c->chunk_ev[kMaxChunks], 0));   // aux starts after everything queued so far
  if (spread) {   // everything queued on the main stream so far (the portion's H2D wait included) precedes the bins
    CU(cudaEventRecord(c->fork_ev, c->stream));
    for (int i = 0; i < ycnr_ctx::kBinStreams; ++i) CU(cudaStreamWaitEvent(c->bin_stream[i], c->fork_ev, 0));
  }
  OK(launch_primal(c, a, p, plan_base));
  OK(launch_dual_bins(c, d, p, plan_base, spread, std::make_integer_sequence<int, kDualBins>{}));
  if (spread) {
    for (int i = 0; i < ycnr_ctx::kBinStreams; ++i) {
      CU(cudaEventRecord(c->join_ev[i], c->bin_stream[i]));
      CU(cudaStreamWaitEvent(c->stream, c->join_ev[i], 0));
    }
  }
  if (c->aux_pending) {
    CU(cudaEventRecord(c->chunk_ev[kMaxChunks + 1], c->aux_stream));
    CU(cudaStreamWaitEvent(c->stream, c->chunk_ev[kMaxChunks + 1], 0));
    c->aux_pending = false;
  }
  c->device_current[solved] = true;
  c->fac_version[solved]++;
  return 0;
}

// RMSE work entries: one 8-lane group walks an entry sequentially, so a user row with thousands of validate ratings
// would be a serial tail that no amount of GPUs shortens.  Rows longer than kRmseChunk ratings are cut into
// entries of at most kRmseChunk (same row id, consecutive ratings); the per-portion reduction adds the entries'
// sums in order, so the result stays deterministic.  efirst[p] = first entry of portion p.
constexpr int kRmseChunk = 64;

void expand_rmse_entries(const int32_t* ids, const int64_t* start, const int32_t* len, int n_rows, const int32_t* pfirst,
                         int P, std::vector<int32_t>& eids, std::vector<int64_t>& estart, std::vector<int32_t>& elen,
                         std::vector<int32_t>& efirst) {
  eids.clear(); estart.clear(); elen.clear();
  efirst.assign((size_t)P + 1, 0);
  int p = 0;
  for (int r = 0; r < n_rows; ++r) {
    while (p < P && pfirst[p] <= r) efirst[p++] = (int32_t)eids.size();
    int n = len[r];
    int64_t st = start[r];
    do {
      const int m = std::min(n, kRmseChunk);
      eids.push_back(ids[r]);
      estart.push_back(st);
      elen.push_back(m);
      st += m;
      n -= m;
    } while (n > 0);
  }
  while (p <= P) efirst[p++] = (int32_t)eids.size();
}

// rows by length, longest first (stable counting sort; ties keep the row order)
void rows_longest_first(const int32_t* len, int n_rows, int32_t* order) {
  int mx = 0;
  for (int r = 0; r < n_rows; ++r) mx = std::max(mx, len[r]);
  const int nb = std::min(mx, 4095) + 1;
  std::vector<int32_t> cnt(nb + 1, 0);
  auto bucket = [&](int r) { return nb - 1 - std::min(std::max(len[r], 0), nb - 1); };
  for (int r = 0; r < n_rows; ++r) cnt[bucket(r) + 1]++;
  for (int b = 0; b < nb; ++b) cnt[b + 1] += cnt[b];
  for (int r = 0; r < n_rows; ++r) order[cnt[bucket(r)]++] = r;
}

int run_rmse(ycnr_ctx* c, const RowsView& view, int n_rows, int64_t nnz, double shift, double* d_row_sums,
             const int32_t* d_portion_first, int n_portions, double* d_portion_sums, double* d_chunk_sums = nullptr,
             const int32_t* d_order = nullptr) {
  if (n_rows <= 0 || n_portions <= 0) return 0;
  ycnr::RmseArgs a{};
  a.rows = view;
  a.U = c->d_fac[YCNR_USER_FACTORS];
  a.V = c->d_fac[YCNR_ITEM_FACTORS];
  a.k = c->k;
  a.n_rows = n_rows;
  a.shift = shift;
  a.row_sums = d_row_sums;
  a.order = d_order;
  {
    ProfScope ps(c, YCNR_K_RMSE_ROWS, n_rows, nnz);
    const int rows_per_cta = 32;   // one 8-lane group per row
    const int grid = (n_rows + rows_per_cta - 1) / rows_per_cta;
    if ((c->k & 3) == 0 && c->k <= 128) ycnr::rmse_rows_kernel<4><<<grid, 256, 0, c->stream>>>(a);
    else if ((c->k & 3) == 0 && c->k <= 256) ycnr::rmse_rows_kernel<8><<<grid, 256, 0, c->stream>>>(a);
    else ycnr::rmse_rows_kernel<0><<<grid, 256, 0, c->stream>>>(a);
  }
  if (d_chunk_sums && n_portions == 1 && n_rows > 2 * ycnr::kRmseChunkRows) {
    // one big portion (per-portion path): two-level fixed-order reduction instead of a single CTA
    const int nch = (n_rows + ycnr::kRmseChunkRows - 1) / ycnr::kRmseChunkRows;
    ProfScope ps(c, YCNR_K_RMSE_REDUCE, n_portions, 0);
    ycnr::rmse_chunk_reduce_kernel<<<nch, 256, 0, c->stream>>>(d_row_sums, view.row_len, n_rows, d_chunk_sums);
    ycnr::rmse_portion_reduce_chunks_kernel<<<1, 256, 0, c->stream>>>(d_chunk_sums, nch, d_portion_sums);
    c->prof.launches[YCNR_K_RMSE_REDUCE] += 1;
    c->prof.total_launches += 1;
  } else {
    ProfScope ps(c, YCNR_K_RMSE_REDUCE, n_portions, 0);
    ycnr::rmse_portion_reduce_kernel<<<n_portions, 256, 0, c->stream>>>(d_row_sums, view.row_len, d_portion_first,
                                                                         d_portion_sums);
  }
  CU(cudaGetLastError());
  return 0;
}

// Device copy of the RMSE work entries of a row set (host arrays of its rows given) + their processing order.
int build_rmse_entries(ycnr_ctx* c, RowSet& rs, const int32_t* ids, const int64_t* start, const int32_t* len,
                       const int32_t* pfirst) {
  std::vector<int32_t> eids, elen, efirst, order;
  std::vector<int64_t> estart;
  expand_rmse_entries(ids, start, len, rs.n_rows, pfirst, rs.n_portions, eids, estart, elen, efirst);
  const int E = (int)eids.size();
  rs.n_entries = E;
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_ids = al((size_t)E * 8), o_len = al(o_ids + (size_t)E * 4), o_pf = al(o_len + (size_t)E * 4);
  OK(rs.erows.ensure(al(o_pf + (size_t)(rs.n_portions + 1) * 4) + 16));
  char* d = (char*)rs.erows.p;
  if (E) {
    CU(cudaMemcpyAsync(d, estart.data(), (size_t)E * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + o_ids, eids.data(), (size_t)E * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + o_len, elen.data(), (size_t)E * 4, cudaMemcpyHostToDevice, c->stream));
    order.resize((size_t)E);
    rows_longest_first(elen.data(), E, order.data());
    OK(rs.order.ensure((size_t)E * 4));
    CU(cudaMemcpyAsync(rs.order.p, order.data(), (size_t)E * 4, cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaMemcpyAsync(d + o_pf, efirst.data(), (size_t)(rs.n_portions + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  rs.eview = rs.view;
  rs.eview.row_start = (const int64_t*)d;
  rs.eview.row_ids = (const int32_t*)(d + o_ids);
  rs.eview.row_len = (const int32_t*)(d + o_len);
  rs.d_efirst = (const int32_t*)(d + o_pf);
  OK(rs.sums.ensure(((size_t)3 * E + 4 * (size_t)rs.n_portions + 4) * sizeof(double)));
  CU(cudaStreamSynchronize(c->stream));   // the host vectors go out of scope
  return 0;
}

int ensure_fixed_current(ycnr_ctx* c, int which) {
  if (c->device_current[which]) return 0;
  if (!c->h_fac[which]) return fail("factor matrix %d has no host copy attached and no device content", which);
  CU(cudaMemcpyAsync(c->d_fac[which], c->h_fac[which], (size_t)c->fac_rows[which] * c->k * sizeof(float),
                     cudaMemcpyHostToDevice, c->stream));
  c->device_current[which] = true;
  c->fac_version[which]++;
  return 0;
}

// Parse a portion header (EmfWorker.js:176-219) and stage header + plan + ratings in one slot.
struct StagedPortion {
  RowsView view{};
  DevPlan plan;
  const int32_t* plan_base = nullptr;
  int n_rows = 0;
  int64_t ratings = 0;
  int32_t first_row = -1, last_row = -1;
  double* d_sums = nullptr;        // rmse scratch: row_sums[R][3] | portion_sums[4] | chunk sums
  const int32_t* d_pfirst = nullptr;
  Slot* slot = nullptr;
};

// The upstream worker trusts the header (EmfWorker.js:176-219); here a stale or corrupt one would turn into
// out-of-bounds device writes (also into every NVLink peer replica), so row ids are checked against the solved
// matrix and must ascend strictly (the master emits them in id order, EmfMaster.js:582-609).
int check_header(const int32_t* rows, int R, int64_t lim_rows, int64_t* ratings_out) {
  int64_t off = 0;
  int64_t prev = -1;
  for (int r = 0; r < R; ++r) {
    const int64_t id = rows[1 + 2 * (size_t)r];
    const int n = rows[2 + 2 * (size_t)r];
    if (n < 0) return fail("portion header: negative cols in row %d", r);
    if (id < 0 || id >= lim_rows) return fail("portion header: row id %lld outside the factor matrix (0..%lld)", (long long)id, (long long)lim_rows - 1);
    if (id <= prev) return fail("portion header: row ids must ascend strictly (row %d: %lld after %lld)", r, (long long)id, (long long)prev);
    prev = id;
    off += n;
  }
  *ratings_out = off;
  return 0;
}

int launch_validate_cols(ycnr_ctx* c, const int32_t* d_indx, int64_t n, int64_t lim_cols) {
  if (n <= 0) return 0;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->num_sms * 8);
  validate_cols_kernel<<<grid, 256, 0, c->stream>>>(d_indx, n, (int32_t)lim_cols, c->d_bad);
  CU(cudaGetLastError());
  c->prof.launches[YCNR_K_GATHER] += 1;
  c->prof.total_launches += 1;
  return 0;
}

int stage_portion(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals, StagedPortion& s) {
  const int R = rows[0];
  if (R < 0) return fail("portion header: negative row count");
  s.n_rows = R;
  const int32_t* hdr_len = rows + 2;   // (rowId, n) pairs: lengths at stride 2
  int64_t off = 0;
  const int solved_w = c->step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
  OK(check_header(rows, R, c->fac_rows[solved_w], &off));
  s.ratings = off;
  if (R > 0) { s.first_row = rows[1]; s.last_row = rows[1 + 2 * (size_t)(R - 1)]; }
  const PlanCfg cfg{c->dual_max, c->split_cols, c->fused_max};
  plan_count(hdr_len, 2, R, cfg, s.plan);
  const size_t pw = s.plan.words;
  // layout (16-byte aligned sections): start[R] i64 | ids[R] | len[R] | pfirst[2] | plan | indx | vals
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t o_start = 0;
  size_t o_ids = al(o_start + (size_t)R * 8);
  size_t o_len = al(o_ids + (size_t)R * 4);
  size_t o_pf = al(o_len + (size_t)R * 4);
  size_t o_plan = al(o_pf + 8);
  size_t o_indx = al(o_plan + pw * 4);
  size_t o_vals = al(o_indx + (size_t)off * 4);
  size_t total = al(o_vals + (size_t)off * 4);

  Slot& sl = c->slots[c->next_slot];
  c->next_slot = (c->next_slot + 1) % kSlots;
  const double tw0 = now_ms();
  if (sl.pending) {
    CU(cudaEventSynchronize(sl.done));
    sl.pending = false;
  }
  c->t_slot_wait += now_ms() - tw0;
  if (!sl.done) CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  if (!sl.solved) CU(cudaEventCreateWithFlags(&sl.solved, cudaEventDisableTiming));
  // ratings that already sit in page-locked caller memory are DMA'd from there
  auto is_pinned = [&](const void* p, size_t bytes) {
    const char* q = (const char*)p;
    for (auto& r : c->pinned)
      if (q >= r.first && q + bytes <= r.first + r.second) return true;
    return false;
  };
  const bool direct = off > 0 && is_pinned(indx, (size_t)off * 4) && is_pinned(vals, (size_t)off * 4);
  const size_t host_need = direct ? o_indx : total;
  if (host_need > sl.host_cap) {
    if (sl.host) cudaFreeHost(sl.host);
    sl.host = nullptr;
    sl.host_cap = 0;
    size_t want = host_need + host_need / 4 + 4096;
    CU(cudaMallocHost(&sl.host, want));
    sl.host_cap = want;
  }
  OK(sl.dev.ensure(total));
  char* h = (char*)sl.host;
  const double tp0 = now_ms();
  {   // header -> row arrays, written in place in the page-locked slot
    int64_t* h_start = (int64_t*)(h + o_start);
    int32_t* h_ids = (int32_t*)(h + o_ids);
    int32_t* h_len = (int32_t*)(h + o_len);
    int64_t run = 0;
    for (int r = 0; r < R; ++r) {
      const int32_t id = rows[1 + 2 * (size_t)r], n = rows[2 + 2 * (size_t)r];
      h_ids[r] = id;
      h_len[r] = n;
      h_start[r] = run;
      run += n;
    }
    plan_fill(h_len, 1, R, cfg, s.plan, (int32_t*)(h + o_plan), c->opts.solve_chunks);
  }
  int32_t pf[2] = {0, R};
  memcpy(h + o_pf, pf, 8);
  const double tp1 = now_ms();
  c->t_parse += tp1 - tp0;
  c->t_portions++;
  char* d = (char*)sl.dev.p;
  if (direct) {
    CU(cudaMemcpyAsync(d, h, o_indx, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d + o_indx, indx, (size_t)off * 4, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d + o_vals, vals, (size_t)off * 4, cudaMemcpyHostToDevice, c->copy_stream));
  } else {
    if (off) {
      memcpy(h + o_indx, indx, (size_t)off * 4);
      memcpy(h + o_vals, vals, (size_t)off * 4);
    }
    CU(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, c->copy_stream));
  }
  CU(cudaEventRecord(c->copied, c->copy_stream));
  CU(cudaStreamWaitEvent(c->stream, c->copied, 0));
  c->t_copy_issue += now_ms() - tp1;
  // Unregistered caller buffers were copied into the slot above and may be refilled once we return.
  // Regions registered with ycnr_host_register are a portion CACHE (usePortionsCache): they are DMA'd
  // asynchronously and must stay unmodified until the step ends (ycnr_end_train_step / the return of
  // ycnr_rmse_portion) — no host wait here, so the copy engine stays busy while the host plans ahead.
  s.view.row_start = (const int64_t*)(d + o_start);
  s.view.row_ids = (const int32_t*)(d + o_ids);
  s.view.row_len = (const int32_t*)(d + o_len);
  s.view.indx = (const int32_t*)(d + o_indx);
  s.view.vals = (const float*)(d + o_vals);
  s.view.guard = c->d_bad;
  s.d_pfirst = (const int32_t*)(d + o_pf);
  s.plan_base = (const int32_t*)(d + o_plan);
  s.slot = &sl;
  OK(launch_validate_cols(c, s.view.indx, off, c->fac_rows[1 - solved_w]));
  return 0;
}

// RMSE portions: the raw header is DMA'd and unpacked on the device (portion_kernels.cuh); the host only
// sums the row lengths (it needs the ratings count to size the copy of indx/vals).
int stage_rmse_portion(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals, StagedPortion& s) {
  const int R = rows[0];
  if (R < 0) return fail("portion header: negative row count");
  s.n_rows = R;
  int64_t off = 0;
  OK(check_header(rows, R, c->fac_rows[YCNR_USER_FACTORS], &off));   // RMSE rows are always users (EmfMaster.js:520-529)
  s.ratings = off;
  if (R > 0) { s.first_row = rows[1]; s.last_row = rows[1 + 2 * (size_t)(R - 1)]; }
  const int nb = (R + ycnr::kUnpackRowsPerBlock - 1) / ycnr::kUnpackRowsPerBlock;
  // device layout (16-byte aligned sections): pad(4 B) hdr[2R+1] | pfirst[2] | start[R] i64 | ids[R] | len[R]
  //                                           | block sums i64[nb] | indx | vals | sums f64[2R+3]
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t hdr_bytes = (size_t)(2 * R + 1) * 4;
  size_t o_hdr = 4;   // hdr[1] (the first (rowId, n) pair) lands on an 8-byte boundary
  size_t o_pf = al(o_hdr + hdr_bytes);
  size_t o_start = al(o_pf + 8);
  size_t o_ids = al(o_start + (size_t)R * 8);
  size_t o_len = al(o_ids + (size_t)R * 4);
  size_t o_bs = al(o_len + (size_t)R * 4);
  size_t o_indx = al(o_bs + (size_t)(nb + 1) * 8);
  size_t o_vals = al(o_indx + (size_t)off * 4);
  size_t o_sums = al(o_vals + (size_t)off * 4);
  // sums: row_sums[R][3] | portion_sums[4] | chunk_sums[ceil(R / chunk)][4]
  size_t dev_total = al(o_sums + ((size_t)(3 * R + 4) + 4 * ((size_t)R / ycnr::kRmseChunkRows + 2)) * 8);

  Slot& sl = c->slots[c->next_slot];
  c->next_slot = (c->next_slot + 1) % kSlots;
  if (sl.pending) {
    CU(cudaEventSynchronize(sl.done));
    sl.pending = false;
  }
  if (!sl.done) CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  if (!sl.solved) CU(cudaEventCreateWithFlags(&sl.solved, cudaEventDisableTiming));
  auto is_pinned = [&](const void* p, size_t bytes) {
    const char* q = (const char*)p;
    for (auto& r : c->pinned)
      if (q >= r.first && q + bytes <= r.first + r.second) return true;
    return false;
  };
  const bool hdr_direct = is_pinned(rows, hdr_bytes);
  const bool direct = off > 0 && is_pinned(indx, (size_t)off * 4) && is_pinned(vals, (size_t)off * 4);
  // staging (page-locked slot memory): pfirst | [hdr] | [indx | vals]
  const size_t h_pf = 0, h_hdr = 16, h_indx = al(h_hdr + (hdr_direct ? 0 : hdr_bytes));
  const size_t h_vals = al(h_indx + (direct ? 0 : (size_t)off * 4));
  const size_t host_need = al(h_vals + (direct ? 0 : (size_t)off * 4));
  if (host_need > sl.host_cap) {
    if (sl.host) cudaFreeHost(sl.host);
    sl.host = nullptr;
    sl.host_cap = 0;
    size_t want = host_need + host_need / 4 + 4096;
    CU(cudaMallocHost(&sl.host, want));
    sl.host_cap = want;
  }
  OK(sl.dev.ensure(dev_total));
  char* h = (char*)sl.host;
  char* d = (char*)sl.dev.p;
  int32_t pf[2] = {0, R};
  memcpy(h + h_pf, pf, 8);
  CU(cudaMemcpyAsync(d + o_pf, h + h_pf, 8, cudaMemcpyHostToDevice, c->copy_stream));
  if (hdr_direct) {
    CU(cudaMemcpyAsync(d + o_hdr, rows, hdr_bytes, cudaMemcpyHostToDevice, c->copy_stream));
  } else {
    memcpy(h + h_hdr, rows, hdr_bytes);
    CU(cudaMemcpyAsync(d + o_hdr, h + h_hdr, hdr_bytes, cudaMemcpyHostToDevice, c->copy_stream));
  }
  if (off) {
    if (direct) {
      CU(cudaMemcpyAsync(d + o_indx, indx, (size_t)off * 4, cudaMemcpyHostToDevice, c->copy_stream));
      CU(cudaMemcpyAsync(d + o_vals, vals, (size_t)off * 4, cudaMemcpyHostToDevice, c->copy_stream));
    } else {
      memcpy(h + h_indx, indx, (size_t)off * 4);
      memcpy(h + h_vals, vals, (size_t)off * 4);
      CU(cudaMemcpyAsync(d + o_indx, h + h_indx, (size_t)off * 4, cudaMemcpyHostToDevice, c->copy_stream));
      CU(cudaMemcpyAsync(d + o_vals, h + h_vals, (size_t)off * 4, cudaMemcpyHostToDevice, c->copy_stream));
    }
  }
  CU(cudaEventRecord(c->copied, c->copy_stream));
  CU(cudaStreamWaitEvent(c->stream, c->copied, 0));
  const int32_t* d_hdr = (const int32_t*)(d + o_hdr);
  int64_t* d_bs = (int64_t*)(d + o_bs);
  if (R > 0) {
    ProfScope ps(c, YCNR_K_GATHER, R, 0);   // accounted with the data-movement kernels
    ycnr::header_block_sums_kernel<<<nb, ycnr::kUnpackThreads, 0, c->stream>>>(d_hdr, R, d_bs);
    ycnr::header_scan_blocks_kernel<<<1, 1024, 0, c->stream>>>(d_bs, nb);
    ycnr::header_unpack_kernel<<<nb, ycnr::kUnpackThreads, 0, c->stream>>>(
        d_hdr, R, d_bs, (int32_t*)(d + o_ids), (int32_t*)(d + o_len), (int64_t*)(d + o_start));
    CU(cudaGetLastError());
    c->prof.launches[YCNR_K_GATHER] += 2;   // three launches in the scope above
    c->prof.total_launches += 2;
  }
  s.view.row_start = (const int64_t*)(d + o_start);
  s.view.row_ids = (const int32_t*)(d + o_ids);
  s.view.row_len = (const int32_t*)(d + o_len);
  s.view.indx = (const int32_t*)(d + o_indx);
  s.view.vals = (const float*)(d + o_vals);
  s.view.guard = c->d_bad;
  s.d_sums = (double*)(d + o_sums);
  s.d_pfirst = (const int32_t*)(d + o_pf);
  s.plan_base = nullptr;
  s.slot = &sl;
  OK(launch_validate_cols(c, s.view.indx, off, c->fac_rows[YCNR_ITEM_FACTORS]));
  return 0;
}

int finish_slot(ycnr_ctx* c, Slot* sl) {
  CU(cudaEventRecord(sl->done, c->stream));
  sl->pending = true;
  return 0;
}

bool region_pinned(ycnr_ctx* c, const void* p, size_t bytes) {
  const char* q = (const char*)p;
  for (auto& r : c->pinned)
    if (q >= r.first && q + bytes <= r.first + r.second) return true;
  return false;
}

int ensure_pinned(void** ptr, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  if (*ptr) cudaFreeHost(*ptr);
  *ptr = nullptr;
  *cap = 0;
  const size_t want = bytes + bytes / 4 + 4096;
  CU(cudaMallocHost(ptr, want));
  *cap = want;
  return 0;
}

// Move the sums of a finished RMSE batch into the completed queue (portion order).
int drain_inflight(ycnr_ctx* c, int idx, bool wait) {
  RmseInflight& f = c->rmse_inflight[idx];
  if (!f.live) return 0;
  if (!wait && cudaEventQuery(f.done) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  CU(cudaEventSynchronize(f.done));
  for (size_t p = 0; p < f.tags.size(); ++p) {
    ycnr_portion_info pi = f.infos[p];
    pi.r_sum_diff2 = f.h_sums[4 * p];
    pi.r_cnt = f.h_sums[4 * p + 1];
    pi.r_sum = f.h_sums[4 * p + 2];
    c->rmse_done.emplace_back(f.tags[p], pi);
    c->rmse_cur.sums[f.tags[p]] = {pi.rows_from, pi.rows_cnt, pi.ratings_in_portion, f.h_sums[4 * p], f.h_sums[4 * p + 1],
                                   f.h_sums[4 * p + 2], f.h_sums[4 * p + 3]};
  }
  f.live = false;
  if (f.slot) f.slot->inflight = -1;
  return 0;
}

int acquire_slot(ycnr_ctx* c, Slot** out) {
  Slot& sl = c->slots[c->next_slot];
  c->next_slot = (c->next_slot + 1) % kSlots;
  const double tw0 = now_ms();
  while (sl.inflight >= 0 && !c->rmse_order.empty()) {   // completions stay in flush order: drain the older ones first
    const int idx = c->rmse_order.front();
    OK(drain_inflight(c, idx, true));
    c->rmse_order.erase(c->rmse_order.begin());
  }
  if (sl.pending) {
    CU(cudaEventSynchronize(sl.done));
    sl.pending = false;
  }
  c->t_slot_wait += now_ms() - tw0;
  if (!sl.done) CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  if (!sl.solved) CU(cudaEventCreateWithFlags(&sl.solved, cudaEventDisableTiming));
  *out = &sl;
  return 0;
}

int batch_flush(ycnr_ctx* c);
int run_fill_tasks(ycnr_ctx* c);

// Row arrays of the open batch live in the slot's page-locked host2 buffer (start i64[cap] | ids i32[cap] |
// len i32[cap]) and are DMA'd from there: the per-row loop below is the only pass over the header.
int batch_rows_reserve(ycnr_ctx* c, Batch& b, size_t need) {
  if (need <= b.cap_rows) return 0;
  const size_t cap = std::max<size_t>(std::max(need + need / 2, 2 * b.cap_rows), 65536);
  void* nh = nullptr;
  CU(cudaMallocHost(&nh, cap * 16));
  int64_t* nstart = (int64_t*)nh;
  int32_t* nids = (int32_t*)((char*)nh + cap * 8);
  int32_t* nlen = (int32_t*)((char*)nh + cap * 12);
  if (b.n_rows) {
    memcpy(nstart, b.start, b.n_rows * 8);
    memcpy(nids, b.ids, b.n_rows * 4);
    memcpy(nlen, b.len, b.n_rows * 4);
  }
  if (b.slot->host2) cudaFreeHost(b.slot->host2);
  b.slot->host2 = nh;
  b.slot->host2_cap = cap * 16;
  b.start = nstart;
  b.ids = nids;
  b.len = nlen;
  b.cap_rows = cap;
  return 0;
}

// Scan of a portion header without writing anything: entries, ratings and the validity flags batch_add checks.
PortionPre scan_header(int kind, const int32_t* rows, int64_t lim_rows) {
  PortionPre p;
  const int R = rows[0];
  if (R < 0 || R >= (1 << 24)) return p;       // not a batch candidate: the single-portion path reports it
  const int32_t* pr = rows + 1;
  int64_t prev = -1, ratings = 0, entries = 0;
  int bad = 0;
  for (int r = 0; r < R; ++r) {
    const int32_t id = pr[2 * (size_t)r], n = pr[2 * (size_t)r + 1];
    bad |= (n < 0) | (id < 0) | (id >= lim_rows) | (id <= prev);
    prev = id;
    if (kind == 1) {
      ratings += n;
    } else {
      ratings += n < 0 ? 0 : n;
      entries += n <= kRmseChunk ? 1 : (n + kRmseChunk - 1) / kRmseChunk;
    }
  }
  p.entries = kind == 1 ? R : entries;
  p.ratings = ratings;
  p.bad = bad;
  return p;
}

// Write the rows (RMSE: work entries) of a scanned, valid header into the batch arrays; run = first rating
void fill_header(int kind, const int32_t* rows, int32_t* pid, int32_t* pln, int64_t* pst, int64_t run) {
  const int R = rows[0];
  const int32_t* pr = rows + 1;
  if (kind == 1) {
    for (int r = 0; r < R; ++r) {
      const int32_t n = pr[2 * (size_t)r + 1];
      pid[r] = pr[2 * (size_t)r];
      pln[r] = n;
      pst[r] = run;
      run += n;
    }
    return;
  }
  size_t e = 0;
  for (int r = 0; r < R; ++r) {
    const int32_t id = pr[2 * (size_t)r];
    int32_t n = pr[2 * (size_t)r + 1];
    do {
      const int32_t m = n < kRmseChunk ? n : kRmseChunk;
      pid[e] = id;
      pln[e] = m;
      pst[e] = run;
      ++e;
      run += m;
      n -= m;
    } while (n > 0);
  }
}

// Deferred header writes of the open batch (queued by the multi-portion entry points), spread over the pool
int run_fill_tasks(ycnr_ctx* c) {
  if (c->fill_tasks.empty()) return 0;
  const double t0 = c->trace ? now_ms() : 0.0;
  Batch& b = c->batch;
  const std::vector<FillTask>& tasks = c->fill_tasks;
  const int n = (int)tasks.size();
  constexpr int kPer = 8;
  const std::function<void(int)> fn = [&](int blk) {
    const int i1 = std::min(n, (blk + 1) * kPer);
    for (int i = blk * kPer; i < i1; ++i) {
      const FillTask& t = tasks[i];
      fill_header(t.kind, t.rows, b.ids + t.r0, b.len + t.r0, b.start + t.r0, t.base);
    }
  };
  c->pool.run((n + kPer - 1) / kPer, fn);
  c->fill_tasks.clear();
  if (c->trace) c->t_fill += now_ms() - t0;
  return 0;
}

// Append one small portion to the open batch of `kind` in ONE pass over its header: the rows are written into the
// batch arrays and checked on the way (row ids inside [0, lim_rows) and strictly ascending, no negative lengths —
// see check_header; 2 ns per row, the separate check + fill loops cost 4.8).  *taken = false (nothing appended)
// when the portion turns out to hold kBatchDirectRatings ratings or more: it takes the single-portion path.
// pre != nullptr (multi-portion entry points): the header was scanned already (scan_header); its rows are written
// later by run_fill_tasks, before the batch is flushed and before the entry point returns.
int batch_add(ycnr_ctx* c, int kind, const int32_t* rows, const int32_t* indx, const float* vals, int64_t lim_rows,
              int64_t tag, ycnr_portion_info* info, bool* taken, const PortionPre* pre = nullptr) {
  *taken = true;
  Batch& b = c->batch;
  struct AddTimer {
    ycnr_ctx* c; double t0;
    ~AddTimer() { c->t_add += now_ms() - t0; }
  } add_timer{c, c->trace ? now_ms() : 0.0};
  if (b.kind != 0 && b.kind != kind) OK(batch_flush(c));
  if (b.kind == 0) {
    OK(acquire_slot(c, &b.slot));
    OK(ensure_pinned(&b.slot->host, &b.slot->host_cap, (size_t)(c->batch_flush_ratings + kBatchDirectRatings) * 8));
    b.kind = kind;
    b.pfirst.clear(); b.tags.clear(); b.infos.clear(); b.segs.clear(); b.ranges.clear();
    b.ratings = b.staged = 0;
    b.n_rows = 0;
    b.cap_rows = b.slot->host2_cap / 16;
    b.start = (int64_t*)b.slot->host2;
    b.ids = (int32_t*)((char*)b.slot->host2 + b.cap_rows * 8);
    b.len = (int32_t*)((char*)b.slot->host2 + b.cap_rows * 12);
    OK(batch_rows_reserve(c, b, c->max_batch_rows + c->max_batch_rows / 4 + 4096));
  }
  const int R = rows[0];
  // entries of this portion: rows, plus (RMSE) one more per kRmseChunk ratings beyond the first chunk of a row;
  // the portion holds fewer than kBatchDirectRatings ratings or it is rolled back below
  OK(batch_rows_reserve(c, b, b.n_rows + (size_t)R + (kind == 2 ? (size_t)(kBatchMaxRatings / kRmseChunk) + 1 : 0)));
  const int64_t base = b.ratings;
  int64_t run = base;
  const size_t r0 = b.n_rows;
  {
    int32_t* pid = b.ids + r0;
    int32_t* pln = b.len + r0;
    int64_t* pst = b.start + r0;
    const int32_t* pr = rows + 1;
    size_t e = 0;
    int64_t prev = -1;
    int bad = 0;
    if (pre) {
      bad = pre->bad;
      e = (size_t)pre->entries;
      run = base + pre->ratings;
    } else if (kind == 1) {
      for (int r = 0; r < R; ++r) {
        const int32_t id = pr[2 * (size_t)r], n = pr[2 * (size_t)r + 1];
        bad |= (n < 0) | (id < 0) | (id >= lim_rows) | (id <= prev);
        prev = id;
        pid[r] = id;
        pln[r] = n;
        pst[r] = run;
        run += n;
      }
      e = (size_t)R;
    } else {   // RMSE: rows cut into entries of at most kRmseChunk ratings (see expand_rmse_entries)
      const size_t e_max = (size_t)R + (size_t)(kBatchMaxRatings / kRmseChunk);
      for (int r = 0; r < R && e < e_max; ++r) {
        const int32_t id = pr[2 * (size_t)r];
        int32_t n = pr[2 * (size_t)r + 1];
        bad |= (n < 0) | (id < 0) | (id >= lim_rows) | (id <= prev);
        prev = id;
        pid[e] = id;
        pst[e] = run;
        if (n <= kRmseChunk) {
          pln[e] = n < 0 ? 0 : n;
          ++e;
          run += n < 0 ? 0 : n;
        } else {
          pln[e] = kRmseChunk;
          ++e;
          run += kRmseChunk;
          n -= kRmseChunk;
          while (n > 0 && e < e_max) {
            const int32_t m = n < kRmseChunk ? n : kRmseChunk;
            pid[e] = id;
            pln[e] = m;
            pst[e] = run;
            ++e;
            run += m;
            n -= m;
          }
        }
      }
    }
    if (bad) {                                                   // rescan for the message
      const int rc = check_header(rows, R, lim_rows, &run);
      return rc ? rc : fail("portion header: invalid");
    }
    const int64_t got = run - base;
    const bool pinned_in = got > 0 && region_pinned(c, indx, (size_t)got * 4) && region_pinned(c, vals, (size_t)got * 4);
    if (got >= (pinned_in ? kBatchMaxRatings : kBatchDirectRatings) ||
        (kind == 2 && e >= (size_t)R + (size_t)(kBatchMaxRatings / kRmseChunk))) {
      // a large portion after all (one that would have to be staged, or a huge one): not for the batch
      *taken = false;
      return 0;
    }
    b.n_rows = r0 + e;
    if (pre) c->fill_tasks.push_back({kind, rows, r0, base});
  }
  const int64_t off = run - base;
  if (kind == 2) b.pfirst.push_back((int32_t)r0);
  const size_t n_added = b.n_rows - r0;
  if (off > 0) {
    const bool direct = region_pinned(c, indx, (size_t)off * 4) && region_pinned(c, vals, (size_t)off * 4);
    BatchSeg* last = b.segs.empty() ? nullptr : &b.segs.back();
    // The master's portion cache is the step's fetch itself (one array): consecutive portions follow each other
    // in memory, separated only by the rating the conversion loop dropped (quirk Q2).  Such a portion extends the
    // previous DMA segment; the few entries in between travel along and are never addressed.
    const int64_t gap = (direct && last && last->direct) ? (int64_t)(indx - (last->indx + last->n)) : -1;
    if (gap >= 0 && gap <= 64 && (int64_t)(vals - (last->vals + last->n)) == gap && last->dev_off + last->n == base &&
        region_pinned(c, last->indx, (size_t)(last->n + gap + off) * 4) &&
        region_pinned(c, last->vals, (size_t)(last->n + gap + off) * 4)) {
      last->n += gap + off;
      if (gap) {
        if (pre) c->fill_tasks.back().base += gap;
        else for (size_t r = 0; r < n_added; ++r) b.start[r0 + r] += gap;
        run += gap;
      }
    } else if (!direct && last && !last->direct) {
      last->n += off;                                           // staged segments are adjacent by construction
    } else {
      b.segs.push_back({indx, vals, off, base, b.staged, direct});
    }
    if (!direct) {
      const size_t cap = (size_t)(c->batch_flush_ratings + kBatchDirectRatings);
      char* h = (char*)b.slot->host;
      memcpy(h + (size_t)b.staged * 4, indx, (size_t)off * 4);
      memcpy(h + cap * 4 + (size_t)b.staged * 4, vals, (size_t)off * 4);
      b.staged += off;
    }
  }
  b.ratings = run;
  const int32_t first = R > 0 ? rows[1] : -1, lastr = R > 0 ? rows[1 + 2 * (size_t)(R - 1)] : -1;
  if (kind == 1 && R > 0) b.ranges.emplace_back(first, lastr);
  if (info) {
    memset(info, 0, sizeof(*info));
    info->rows_from = first;
    info->rows_cnt = R;
    info->ratings_in_portion = off;
  }
  if (kind == 2) {
    ycnr_portion_info pi;
    memset(&pi, 0, sizeof(pi));
    pi.rows_from = first;
    pi.rows_cnt = R;
    pi.ratings_in_portion = off;
    b.tags.push_back(tag);
    b.infos.push_back(pi);
  }
  c->t_portions++;
  // ramp-up: the first batches of a step are an eighth, a quarter, half of the flush size — the GPU starts after
  // ~0.5 M ratings have been queued instead of 4 M (measured on MAL byUser: the step began with ~3 ms of host work
  // and DMA before the first kernel)
  const int64_t flush_at = std::max<int64_t>(1, c->batch_flush_ratings >> std::max(0, 3 - c->flushes_in_pass));
  if (b.ratings >= flush_at) OK(batch_flush(c));
  return 0;
}

int batch_flush(ycnr_ctx* c) {
  Batch& b = c->batch;
  if (b.kind == 0) return 0;
  OK(run_fill_tasks(c));
  c->flushes_in_pass++;
  const int kind = b.kind;
  b.kind = 0;
  Slot& sl = *b.slot;
  const int R = (int)b.n_rows;
  const int P = (int)b.tags.size();
  const int64_t off = b.ratings;
  c->max_batch_rows = std::max(c->max_batch_rows, b.n_rows);
  const double tp0 = now_ms();
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  DevPlan plan;
  const PlanCfg cfg{c->dual_max, c->split_cols, c->fused_max};
  if (kind == 1) plan_count(b.len, 1, R, cfg, plan);
  const size_t pw = kind == 1 ? plan.words : 0;
  // device: start[R] i64 | ids[R] | len[R] | pfirst[P+1] | plan | indx | vals | rmse sums
  // host3 (page-locked): pfirst | plan, then (RMSE) the landing area of the portion sums
  size_t o_start = 0;
  size_t o_ids = al(o_start + (size_t)R * 8);
  size_t o_len = al(o_ids + (size_t)R * 4);
  size_t o_pf = al(o_len + (size_t)R * 4);
  size_t o_plan = al(o_pf + (size_t)(P + 2) * 4);
  size_t o_hdr_end = al(o_plan + pw * 4);
  size_t o_indx = o_hdr_end;
  size_t o_vals = al(o_indx + (size_t)off * 4);
  size_t o_sums = al(o_vals + (size_t)off * 4);
  size_t total = kind == 2 ? al(o_sums + ((size_t)3 * R + 4 * (size_t)P + 8) * 8) : o_sums;
  const size_t h3_bytes = o_hdr_end - o_pf, h3_sums = al(h3_bytes);
  c->max_host3 = std::max(c->max_host3, h3_sums + (size_t)(4 * P + 4) * 8);
  OK(ensure_pinned(&sl.host3, &sl.host3_cap, c->max_host3));   // (page-locked allocations cost milliseconds: grow all slots once)
  c->max_slot_dev = std::max(c->max_slot_dev, total);
  OK(sl.dev.ensure(c->max_slot_dev));
  char* h3 = (char*)sl.host3;
  char* d = (char*)sl.dev.p;
  memset(h3, 0, (size_t)(P + 2) * 4);
  if (kind == 2) {
    memcpy(h3, b.pfirst.data(), (size_t)P * 4);
    ((int32_t*)h3)[P] = R;
    // (entries stay in portion order here: sorting ~1.7 M validate rows per pass on the host costs more than the
    //  kernel gains from it)
  } else {
    plan_fill(b.len, 1, R, cfg, plan, (int32_t*)(h3 + (o_plan - o_pf)), c->opts.solve_chunks);
  }
  const double tp1 = now_ms();
  c->t_parse += tp1 - tp0;
  if (R) {
    CU(cudaMemcpyAsync(d + o_start, b.start, (size_t)R * 8, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d + o_ids, b.ids, (size_t)R * 4, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d + o_len, b.len, (size_t)R * 4, cudaMemcpyHostToDevice, c->copy_stream));
  }
  CU(cudaMemcpyAsync(d + o_pf, h3, h3_bytes, cudaMemcpyHostToDevice, c->copy_stream));
  const size_t cap = (size_t)(c->batch_flush_ratings + kBatchDirectRatings);
  for (const BatchSeg& sg : b.segs) {
    const void* si = sg.direct ? (const void*)sg.indx : (const void*)((char*)sl.host + (size_t)sg.stage_off * 4);
    const void* sv = sg.direct ? (const void*)sg.vals : (const void*)((char*)sl.host + cap * 4 + (size_t)sg.stage_off * 4);
    CU(cudaMemcpyAsync(d + o_indx + (size_t)sg.dev_off * 4, si, (size_t)sg.n * 4, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(d + o_vals + (size_t)sg.dev_off * 4, sv, (size_t)sg.n * 4, cudaMemcpyHostToDevice, c->copy_stream));
  }
  CU(cudaEventRecord(c->copied, c->copy_stream));
  CU(cudaStreamWaitEvent(c->stream, c->copied, 0));
  c->t_copy_issue += now_ms() - tp1;
  RowsView view{};
  view.row_start = (const int64_t*)(d + o_start);
  view.row_ids = (const int32_t*)(d + o_ids);
  view.row_len = (const int32_t*)(d + o_len);
  view.indx = (const int32_t*)(d + o_indx);
  view.vals = (const float*)(d + o_vals);
  view.guard = c->d_bad;
  if (kind == 1) {
    const int solved = c->step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
    OK(launch_validate_cols(c, view.indx, off, c->fac_rows[1 - solved]));
    if (R > 0) {
      const double tl0 = now_ms();
      OK(run_als(c, c->step_type, view, plan, (const int32_t*)(d + o_plan), c->spread_batch));
      c->t_launch += now_ms() - tl0;
      if (c->h_fac[solved] && c->h_registered[solved]) {
        // solved rows back to the host segment under the next batch's kernels; portions whose id ranges touch
        // are copied as one range (rows between two ranges may belong to another worker and are left alone)
        CU(cudaEventRecord(sl.solved, c->stream));
        CU(cudaStreamWaitEvent(c->d2h_stream, sl.solved, 0));
        size_t i = 0;
        while (i < b.ranges.size()) {
          int32_t lo = b.ranges[i].first, hi = b.ranges[i].second;
          size_t j = i + 1;
          while (j < b.ranges.size() && b.ranges[j].first == hi + 1) { hi = b.ranges[j].second; ++j; }
          const size_t fo = (size_t)lo * c->k;
          CU(cudaMemcpyAsync(c->h_fac[solved] + fo, c->d_fac[solved] + fo, (size_t)(hi - lo + 1) * c->k * sizeof(float),
                             cudaMemcpyDeviceToHost, c->d2h_stream));
          i = j;
        }
        c->d2h_pending = true;
      } else {
        for (auto& r : b.ranges) c->solved_ranges.push_back(r);
      }
    }
    OK(finish_slot(c, &sl));
  } else {
    OK(launch_validate_cols(c, view.indx, off, c->fac_rows[YCNR_ITEM_FACTORS]));
    double* d_rows = (double*)(d + o_sums);
    double* d_port = d_rows + 3 * (size_t)R;
    if (R > 0) {
      OK(run_rmse(c, view, R, off, c->rmse_shift, d_rows, (const int32_t*)(d + o_pf), P, d_port));
    } else {
      CU(cudaMemsetAsync(d_port, 0, sizeof(double) * 4 * P, c->stream));
    }
    CU(cudaMemcpyAsync(h3 + h3_sums, d_port, sizeof(double) * 4 * P, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_bad, c->d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    OK(finish_slot(c, &sl));
    int idx = -1;
    for (size_t i = 0; i < c->rmse_inflight.size(); ++i)
      if (!c->rmse_inflight[i].live) { idx = (int)i; break; }
    if (idx < 0) { c->rmse_inflight.emplace_back(); idx = (int)c->rmse_inflight.size() - 1; }
    RmseInflight& f = c->rmse_inflight[idx];
    f.slot = &sl;
    f.done = sl.done;
    f.tags = b.tags;
    f.infos = b.infos;
    f.h_sums = (const double*)(h3 + h3_sums);
    f.live = true;
    sl.inflight = idx;
    c->rmse_order.push_back(idx);
  }
  return 0;
}

// Completed RMSE portions, oldest first.  wait: also flush the open batch and wait for everything in flight.
int rmse_collect(ycnr_ctx* c, bool wait) {
  if (wait) OK(batch_flush(c));
  while (!c->rmse_order.empty()) {
    const int idx = c->rmse_order.front();
    OK(drain_inflight(c, idx, wait));
    if (c->rmse_inflight[idx].live) break;      // not finished yet (and we do not wait): keep the order
    c->rmse_order.erase(c->rmse_order.begin());
  }
  return 0;
}

void collect_profile(ycnr_ctx* c) {
  for (auto& r : c->prof_open) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      c->prof.ms[r.cls] += ms;
      if (r.sub >= 0) c->dual_bin_ms[r.sub] += ms;
    }
    c->ev_pool.push_back(r.a);
    c->ev_pool.push_back(r.b);
  }
  c->prof_open.clear();
}

}  // namespace

extern "C" {

const char* ycnr_last_error(void) { return g_err; }

int ycnr_device_count(int32_t* count_out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count_out = 0;
    return fail("cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count_out = n;
  return 0;
}

int ycnr_create(const ycnr_options* o, ycnr_ctx** out) {
  if (!o || !out) return fail("ycnr_create: null argument");
  *out = nullptr;
  if (o->use_double_precision)
    return fail("useDoublePrecision=true is not supported: the B200 path computes in float32 only (EmfBase.js:112)");
  if (o->lowmem) return fail("lowmem=true (file-backed factors, EmfBase.js:116) is not supported by the GPU path");
  if (o->factors_count <= 0 || o->factors_count > 256)
    return fail("factorsCount %d outside the supported range 1..256", o->factors_count);
  if (o->factors_count > 128 && ((o->factors_count & 3) || o->gram_path == YCNR_GRAM_FFMA))
    return fail("factorsCount %d > 128 runs on the tensor-core Gram only: it needs factorsCount %% 4 == 0 and gramPath auto or tc",
                o->factors_count);
  if (o->total_users <= 0 || o->total_items <= 0) return fail("totalUsersCount/totalItemsCount must be positive");
  if (o->gram_path == YCNR_GRAM_TC3XTF32 && ((o->factors_count & 3) || o->factors_count < 8))
    return fail("gram_path=TC3XTF32 needs factorsCount %% 4 == 0 and >= 8 (got %d)", o->factors_count);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail("no CUDA device available (%s): the ALS hot path has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (o->device < 0 || o->device >= ndev) return fail("device %d out of range (0..%d)", o->device, ndev - 1);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, o->device));
  if (prop.major != 10)
    return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", o->device, prop.major, prop.minor);
  ycnr_ctx* c = new ycnr_ctx();
  c->opts = *o;
  c->k = o->factors_count;
  bool use_tc = false;
  const PlanCfg cfg = derive_plan_cfg(*o, &use_tc);
  c->dual_max = cfg.dual_max;
  c->split_cols = cfg.split_cols;
  c->fused_max = cfg.fused_max;
  c->use_tc = use_tc;
  // systems wider than one tensor-core pass keep nb (nb - 1) / 2 sets of tile partials: bound them by working
  // through the split rows in up to 8 groups (launch_primal_blocks)
  if (use_tc && c->k > 124 && c->opts.solve_chunks < 2) c->opts.solve_chunks = kMaxChunks;
  c->num_sms = prop.multiProcessorCount;
  c->trace = getenv("YCNR_TRACE") != nullptr;
  if (const char* e = getenv("YCNR_BATCH_FLUSH")) c->batch_flush_ratings = std::max<int64_t>(1, atoll(e));
  if (const char* e = getenv("YCNR_SPREAD_BULK")) c->spread_bulk = atoi(e) ? 1 : 0;
  if (const char* e = getenv("YCNR_SPREAD_BATCH")) c->spread_batch = atoi(e) != 0;
  {   // host threads of the multi-portion entry points (the caller is one of them)
    const int hw = (int)std::thread::hardware_concurrency();
    int threads = std::max(1, std::min(6, hw / 4));
    if (const char* e = getenv("YCNR_HOST_THREADS")) threads = std::max(1, std::min(64, atoi(e)));
    c->pool.start(threads - 1);
  }
  c->fac_rows[0] = o->total_users;
  c->fac_rows[1] = o->total_items;
  CU(cudaSetDevice(o->device));
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->copied, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
  for (int i = 0; i < ycnr_ctx::kBinStreams; ++i) {
    CU(cudaStreamCreateWithFlags(&c->bin_stream[i], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
  }
  for (auto& e : c->chunk_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (int w = 0; w < 2; ++w) CU(cudaMalloc(&c->d_fac[w], (size_t)c->fac_rows[w] * c->k * sizeof(float)));
  c->no_tma = getenv("YCNR_NO_TMA") != nullptr;
  if (const char* e = getenv("YCNR_TC_VARIANT")) c->opts.tc_variant = atoi(e);   // diagnostics (gram_tc.cuh: variant bits)
  if (use_tc && (c->k & 3) == 0) {
    // TMA descriptors for the gather of the tensor-core Gram (row stride k * 4 bytes is a multiple of 16)
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres) != cudaSuccess) {
      cudaGetLastError();
      enc = nullptr;
    }
    for (int w = 0; w < 2 && enc; ++w) {
      cuuint64_t gdim[2] = {(cuuint64_t)c->k, (cuuint64_t)c->fac_rows[w]};
      cuuint64_t gstr[1] = {(cuuint64_t)c->k * sizeof(float)};
      cuuint32_t box[2] = {32, 1};
      cuuint32_t estr[2] = {1, 1};
      c->tma_ok[w] = enc(&c->tmap[w], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c->d_fac[w], gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
  }
  CU(cudaMalloc(&c->d_bad, sizeof(int32_t)));
  CU(cudaMemset(c->d_bad, 0, sizeof(int32_t)));
  CU(cudaMallocHost(&c->h_bad, sizeof(int32_t)));
  *c->h_bad = 0;
  g_default_ctx = c;
  *out = c;
  return 0;
}

int ycnr_destroy(ycnr_ctx* c) {
  if (!c) return 0;
  c->pool.shutdown();
  cudaSetDevice(c->opts.device);
  cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->stream);
  if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
  if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
  collect_profile(c);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  for (int w = 0; w < 2; ++w) {
    if (c->h_registered[w]) cudaHostUnregister(c->h_fac[w]);
    if (c->d_fac[w]) cudaFree(c->d_fac[w]);
  }
  for (auto& rs : c->rowsets) {
    rs.rows.release(); rs.ratings.release(); rs.plan.release(); rs.sums.release(); rs.order.release(); rs.erows.release();
    if (rs.h_sums) cudaFreeHost(rs.h_sums);
    if (rs.sums_ready) cudaEventDestroy(rs.sums_ready);
  }
  for (auto& s : c->slots) {
    if (s.host) cudaFreeHost(s.host);
    if (s.host2) cudaFreeHost(s.host2);
    if (s.host3) cudaFreeHost(s.host3);
    s.dev.release();
    if (s.done) cudaEventDestroy(s.done);
    if (s.solved) cudaEventDestroy(s.solved);
  }
  if (g_default_ctx == c) g_default_ctx = nullptr;
  if (c->d_bad) cudaFree(c->d_bad);
  if (c->h_bad) cudaFreeHost(c->h_bad);
  c->partial.release();
  c->gather_tmp.release();
  c->table.buf.release();
  c->ingest_tmp.release();
  for (auto& r : c->pinned) cudaHostUnregister((void*)r.first);
  if (c->copied) cudaEventDestroy(c->copied);
  for (auto e : c->chunk_ev) if (e) cudaEventDestroy(e);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (int i = 0; i < ycnr_ctx::kBinStreams; ++i) {
    if (c->bin_stream[i]) { cudaStreamSynchronize(c->bin_stream[i]); cudaStreamDestroy(c->bin_stream[i]); }
    if (c->join_ev[i]) cudaEventDestroy(c->join_ev[i]);
  }
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

// ---- factor store ---------------------------------------------------------------------
int ycnr_attach_factors(ycnr_ctx* c, float* uf, float* vf) {
  if (!c || !uf || !vf) return fail("ycnr_attach_factors: null argument");
  OK(set_device(c));
  float* hp[2] = {uf, vf};
  for (int w = 0; w < 2; ++w) {
    if (c->h_registered[w]) { cudaHostUnregister(c->h_fac[w]); c->h_registered[w] = false; }
    c->h_fac[w] = hp[w];
    const size_t bytes = (size_t)c->fac_rows[w] * c->k * sizeof(float);
    // page-lock the shm segment so the per-step copies are true DMA; pageable still works
    if (cudaHostRegister(hp[w], bytes, cudaHostRegisterDefault) == cudaSuccess) c->h_registered[w] = true;
    else cudaGetLastError();
    c->device_current[w] = false;
    OK(ensure_fixed_current(c, w));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ycnr_upload_factors(ycnr_ctx* c, int32_t which) {
  if (!c || which < 0 || which > 1) return fail("ycnr_upload_factors: bad argument");
  OK(set_device(c));
  c->device_current[which] = false;
  OK(ensure_fixed_current(c, which));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ycnr_download_factors(ycnr_ctx* c, int32_t which, int32_t row_from, int32_t row_cnt) {
  if (!c || which < 0 || which > 1) return fail("ycnr_download_factors: bad argument");
  if (!c->h_fac[which]) return fail("ycnr_download_factors: no host matrix attached");
  if (row_cnt < 0) { row_from = 0; row_cnt = (int32_t)c->fac_rows[which]; }
  if (row_from < 0 || (int64_t)row_from + row_cnt > c->fac_rows[which]) return fail("ycnr_download_factors: row range");
  OK(set_device(c));
  const size_t off = (size_t)row_from * c->k;
  CU(cudaMemcpyAsync(c->h_fac[which] + off, c->d_fac[which] + off, (size_t)row_cnt * c->k * sizeof(float),
                     cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ycnr_invalidate_device(ycnr_ctx* c, int32_t which) {
  if (!c || which < 0 || which > 1) return fail("ycnr_invalidate_device: bad argument");
  c->device_current[which] = false;
  return 0;
}

int ycnr_device_factors(ycnr_ctx* c, int32_t which, void** out) {
  if (!c || which < 0 || which > 1 || !out) return fail("ycnr_device_factors: bad argument");
  *out = c->d_fac[which];
  return 0;
}

int ycnr_stream(ycnr_ctx* c, void** out) {
  if (!c || !out) return fail("ycnr_stream: bad argument");
  *out = (void*)c->stream;
  return 0;
}

int ycnr_synchronize(ycnr_ctx* c) {
  if (!c) return fail("ycnr_synchronize: null context");
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  if (c->d2h_pending) {
    CU(cudaStreamSynchronize(c->d2h_stream));
    c->d2h_pending = false;
  }
  return 0;
}

int ycnr_host_register(ycnr_ctx* c, void* ptr, size_t bytes) {
  if (!c || !ptr || bytes == 0) return fail("ycnr_host_register: bad argument");
  OK(set_device(c));
  for (auto& r : c->pinned)
    if (r.first == (const char*)ptr) return r.second >= bytes ? 0 : fail("ycnr_host_register: region already registered with a smaller size");
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  c->pinned.emplace_back((const char*)ptr, bytes);
  return 0;
}

int ycnr_host_unregister(ycnr_ctx* c, void* ptr) {
  if (!c || !ptr) return fail("ycnr_host_unregister: bad argument");
  OK(set_device(c));
  for (size_t i = 0; i < c->pinned.size(); ++i)
    if (c->pinned[i].first == (const char*)ptr) {
      CU(cudaStreamSynchronize(c->copy_stream));
      CU(cudaHostUnregister(ptr));
      c->pinned.erase(c->pinned.begin() + i);
      return 0;
    }
  return fail("ycnr_host_unregister: region not registered");
}

// ---- per-portion path -----------------------------------------------------------------
int ycnr_start_train_step(ycnr_ctx* c, int32_t step_type) {
  if (!c || (step_type != YCNR_BY_USER && step_type != YCNR_BY_ITEM)) return fail("ycnr_start_train_step: bad stepType");
  OK(set_device(c));
  OK(rmse_collect(c, true));   // (nothing is left queued from an earlier pass)
  c->step_type = step_type;
  c->flushes_in_pass = 0;
  c->solved_ranges.clear();
  const int solved = step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
  c->fac_version[solved]++;   // peers may store into this replica during the step even if no portion arrives here
  OK(ensure_fixed_current(c, 1 - solved));
  OK(ensure_fixed_current(c, solved));  // rows that are not solved keep their values
  return 0;
}

static int als_portion_impl(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals,
                            ycnr_portion_info* info, const PortionPre* pre);

int ycnr_als_portion(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals,
                     ycnr_portion_info* info) {
  return als_portion_impl(c, rows, indx, vals, info, nullptr);
}

static int als_portion_impl(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals,
                            ycnr_portion_info* info, const PortionPre* pre) {
  if (!c || !rows || !indx || !vals) return fail("ycnr_als_portion: null argument");
  if (c->step_type != YCNR_BY_USER && c->step_type != YCNR_BY_ITEM)
    return fail("ycnr_als_portion: call ycnr_start_train_step first");
  const double t0 = now_ms();
  OK(set_device(c));
  if (rows[0] >= 0 && rows[0] < (1 << 24)) {   // portions are queued and launched in batches
    const int solved_w = c->step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
    bool taken = false;
    OK(batch_add(c, 1, rows, indx, vals, c->fac_rows[solved_w], 0, info, &taken, pre));
    if (taken) {
      if (info) info->time_ms = now_ms() - t0;
      return 0;
    }
  }
  OK(batch_flush(c));
  StagedPortion s;
  OK(stage_portion(c, rows, indx, vals, s));
  if (s.n_rows > 0) {
    const double tl0 = now_ms();
    OK(run_als(c, c->step_type, s.view, s.plan, s.plan_base, true));
    c->t_launch += now_ms() - tl0;
    const int solved = c->step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
    if (c->h_fac[solved] && c->h_registered[solved]) {
      // the portion's row-id range goes back to the host segment on its own stream as soon as its kernels
      // are done, under the next portion's kernels (portions cover disjoint ascending id ranges)
      CU(cudaEventRecord(s.slot->solved, c->stream));
      CU(cudaStreamWaitEvent(c->d2h_stream, s.slot->solved, 0));
      const size_t off = (size_t)s.first_row * c->k;
      CU(cudaMemcpyAsync(c->h_fac[solved] + off, c->d_fac[solved] + off,
                         (size_t)(s.last_row - s.first_row + 1) * c->k * sizeof(float), cudaMemcpyDeviceToHost,
                         c->d2h_stream));
      c->d2h_pending = true;
    } else {
      c->solved_ranges.emplace_back(s.first_row, s.last_row);   // pageable segment: one merged copy at the end
    }
  }
  OK(finish_slot(c, s.slot));
  if (info) {
    memset(info, 0, sizeof(*info));
    info->rows_from = s.first_row;
    info->rows_cnt = s.n_rows;
    info->ratings_in_portion = s.ratings;
    info->time_ms = now_ms() - t0;
  }
  return 0;
}

int ycnr_end_train_step(ycnr_ctx* c) {
  if (!c) return fail("ycnr_end_train_step: null context");
  if (c->step_type != YCNR_BY_USER && c->step_type != YCNR_BY_ITEM) return fail("ycnr_end_train_step: no step open");
  OK(set_device(c));
  OK(batch_flush(c));
  const int solved = c->step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
  if (c->h_fac[solved]) {
    // merge the per-portion [first,last] ranges (ascending in practice) and copy them back
    auto& v = c->solved_ranges;
    std::sort(v.begin(), v.end());
    size_t i = 0;
    while (i < v.size()) {
      int32_t lo = v[i].first, hi = v[i].second;
      size_t j = i + 1;
      while (j < v.size() && v[j].first <= hi + 1) { hi = std::max(hi, v[j].second); ++j; }
      const size_t off = (size_t)lo * c->k;
      CU(cudaMemcpyAsync(c->h_fac[solved] + off, c->d_fac[solved] + off, (size_t)(hi - lo + 1) * c->k * sizeof(float),
                         cudaMemcpyDeviceToHost, c->stream));
      i = j;
    }
  }
  CU(cudaMemcpyAsync(c->h_bad, c->d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  const double ts0 = now_ms();
  CU(cudaStreamSynchronize(c->stream));
  const double ts1 = now_ms();
  if (c->trace) {
    fprintf(stderr, "[ycnr trace] step %d: %d portions, host parse %.2f ms, slot wait %.2f ms, copy issue %.2f ms, "
            "launch %.2f ms, batch_add total %.2f ms, header scan %.2f ms, header fill %.2f ms, final sync %.2f ms\n",
            c->step_type, c->t_portions, c->t_parse, c->t_slot_wait, c->t_copy_issue, c->t_launch, c->t_add, c->t_scan,
            c->t_fill, ts1 - ts0);
  }
  c->t_parse = c->t_slot_wait = c->t_copy_issue = c->t_launch = c->t_add = c->t_fill = c->t_scan = 0;
  c->t_portions = 0;
  if (c->d2h_pending) {
    CU(cudaStreamSynchronize(c->d2h_stream));
    c->d2h_pending = false;
  }
  c->solved_ranges.clear();
  c->step_type = -1;
  if (*c->h_bad) {
    *c->h_bad = 0;
    CU(cudaMemset(c->d_bad, 0, sizeof(int32_t)));
    return fail("ycnr_als_portion: a portion of this step held a column id outside the fixed factor matrix "
                "(0..%lld); its rows and those of the portions queued after it were not solved",
                (long long)c->fac_rows[1 - solved] - 1);
  }
  return 0;
}

int ycnr_start_calc_rmse(ycnr_ctx* c, int32_t step_type, double shift) {
  if (!c || (step_type != YCNR_RMSE_VALIDATE && step_type != YCNR_RMSE_TEST)) return fail("ycnr_start_calc_rmse: bad stepType");
  OK(set_device(c));
  OK(rmse_collect(c, true));   // a queued batch still belongs to the previous pass (its shift)
  c->step_type = step_type;
  c->flushes_in_pass = 0;
  c->rmse_shift = shift;
  OK(ensure_fixed_current(c, YCNR_USER_FACTORS));
  OK(ensure_fixed_current(c, YCNR_ITEM_FACTORS));
  if (!c->rmse_cur.sums.empty()) std::swap(c->rmse_prev, c->rmse_cur);
  c->rmse_cur.sums.clear();
  c->rmse_cur.step_type = step_type;
  c->rmse_cur.shift = shift;
  c->rmse_cur.ver[0] = c->fac_version[0];
  c->rmse_cur.ver[1] = c->fac_version[1];
  return 0;
}

static int rmse_portion_sync(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals, ycnr_portion_info* info,
                             double* ratings_sum_out);

int ycnr_rmse_portion(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals,
                      ycnr_portion_info* info) {
  return rmse_portion_sync(c, rows, indx, vals, info, nullptr);
}

static int rmse_portion_sync(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals, ycnr_portion_info* info,
                             double* ratings_sum_out) {
  if (!c || !rows || !indx || !vals || !info) return fail("ycnr_rmse_portion: null argument");
  if (c->step_type != YCNR_RMSE_VALIDATE && c->step_type != YCNR_RMSE_TEST)
    return fail("ycnr_rmse_portion: call ycnr_start_calc_rmse first");
  const double t0 = now_ms();
  OK(set_device(c));
  OK(rmse_collect(c, true));
  StagedPortion s;
  OK(stage_rmse_portion(c, rows, indx, vals, s));
  double sums[4] = {0, 0, 0, 0};
  if (s.n_rows > 0) {
    double* d_portion = s.d_sums + 3 * (size_t)s.n_rows;
    OK(run_rmse(c, s.view, s.n_rows, s.ratings, c->rmse_shift, s.d_sums, s.d_pfirst, 1, d_portion, d_portion + 4));
    CU(cudaMemcpyAsync(sums, d_portion, sizeof(sums), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_bad, c->d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  OK(finish_slot(c, s.slot));
  CU(cudaStreamSynchronize(c->stream));
  if (*c->h_bad) {
    *c->h_bad = 0;
    CU(cudaMemset(c->d_bad, 0, sizeof(int32_t)));
    return fail("ycnr_rmse_portion: item id outside the item factor matrix (0..%lld)", (long long)c->fac_rows[1] - 1);
  }
  memset(info, 0, sizeof(*info));
  info->rows_from = s.first_row;
  info->rows_cnt = s.n_rows;
  info->ratings_in_portion = s.ratings;
  info->r_sum_diff2 = sums[0];
  info->r_cnt = sums[1];
  info->r_sum = sums[2];
  info->time_ms = now_ms() - t0;
  if (ratings_sum_out) *ratings_sum_out = sums[3];
  return 0;
}

static int rmse_portion_async_impl(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals, int64_t tag,
                                   const PortionPre* pre);

int ycnr_rmse_portion_async(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals, int64_t tag) {
  return rmse_portion_async_impl(c, rows, indx, vals, tag, nullptr);
}

static int rmse_portion_async_impl(ycnr_ctx* c, const int32_t* rows, const int32_t* indx, const float* vals, int64_t tag,
                                   const PortionPre* pre) {
  if (!c || !rows || !indx || !vals) return fail("ycnr_rmse_portion_async: null argument");
  if (c->step_type != YCNR_RMSE_VALIDATE && c->step_type != YCNR_RMSE_TEST)
    return fail("ycnr_rmse_portion_async: call ycnr_start_calc_rmse first");
  OK(set_device(c));
  {   // the same portion of the same set under the previous shift, factors untouched: derive (see rmse_prev)
    const auto& pv = c->rmse_prev;
    if (pv.step_type == c->step_type && pv.ver[0] == c->fac_version[0] && pv.ver[1] == c->fac_version[1] &&
        c->batch.kind == 0 && c->rmse_order.empty()) {
      auto it = pv.sums.find(tag);
      const int R = rows[0];
      if (it != pv.sums.end() && it->second.rows_cnt == R && (R == 0 || it->second.rows_from == rows[1])) {
        const auto& q = it->second;
        const double d = c->rmse_shift - pv.shift;
        ycnr_portion_info pi;
        memset(&pi, 0, sizeof(pi));
        pi.rows_from = q.rows_from;
        pi.rows_cnt = q.rows_cnt;
        pi.ratings_in_portion = q.ratings;
        pi.r_sum_diff2 = q.d2 - 2.0 * d * (q.rat - q.pred) + q.cnt * d * d;
        pi.r_cnt = q.cnt;
        pi.r_sum = q.pred + q.cnt * d;
        c->rmse_done.emplace_back(tag, pi);
        c->rmse_cur.sums[tag] = {q.rows_from, q.rows_cnt, q.ratings, pi.r_sum_diff2, q.cnt, pi.r_sum, q.rat};
        return 0;
      }
    }
  }
  if (rows[0] >= 0 && rows[0] < (1 << 20)) {   // (portions of a million rows keep the device-side header unpack)
    bool taken = false;
    OK(batch_add(c, 2, rows, indx, vals, c->fac_rows[YCNR_USER_FACTORS], tag, nullptr, &taken, pre));
    if (taken) return 0;
  }
  ycnr_portion_info pi;
  double rat = 0.0;
  OK(rmse_portion_sync(c, rows, indx, vals, &pi, &rat));     // large portion: the single-portion path, synchronous
  c->rmse_done.emplace_back(tag, pi);
  c->rmse_cur.sums[tag] = {pi.rows_from, pi.rows_cnt, pi.ratings_in_portion, pi.r_sum_diff2, pi.r_cnt, pi.r_sum, rat};
  return 0;
}

int ycnr_rmse_poll(ycnr_ctx* c, int32_t wait, int32_t max_out, int64_t* tags_out, ycnr_portion_info* infos_out,
                   int32_t* n_out) {
  if (!c || !n_out || max_out < 0 || (max_out && (!tags_out || !infos_out))) return fail("ycnr_rmse_poll: bad argument");
  OK(set_device(c));
  *n_out = 0;
  const double tc0 = now_ms();
  OK(rmse_collect(c, wait != 0));
  if (c->trace && wait) {
    fprintf(stderr, "[ycnr trace] rmse pass %d: %d portions, calls %.2f ms (batch_add %.2f incl. flushes: plan/pack %.2f, "
            "slot wait %.2f, copy issue %.2f), final collect %.2f ms\n", c->step_type, c->t_portions, c->t_rmse_calls,
            c->t_add, c->t_parse, c->t_slot_wait, c->t_copy_issue, now_ms() - tc0);
    c->t_parse = c->t_slot_wait = c->t_copy_issue = c->t_launch = c->t_add = c->t_rmse_calls = 0;
    c->t_portions = 0;
  }
  if (*c->h_bad) {
    *c->h_bad = 0;
    CU(cudaMemset(c->d_bad, 0, sizeof(int32_t)));
    return fail("ycnr_rmse_portion: item id outside the item factor matrix (0..%lld)", (long long)c->fac_rows[1] - 1);
  }
  int n = 0;
  while (n < max_out && c->rmse_done_head < c->rmse_done.size()) {
    tags_out[n] = c->rmse_done[c->rmse_done_head].first;
    infos_out[n] = c->rmse_done[c->rmse_done_head].second;
    ++c->rmse_done_head;
    ++n;
  }
  if (c->rmse_done_head == c->rmse_done.size()) {
    c->rmse_done.clear();
    c->rmse_done_head = 0;
  }
  *n_out = n;
  return 0;
}

// n calls of ycnr_als_portion / ycnr_rmse_portion_async issued from native code (a binding whose per-call overhead
// matters — a Python loop spends ~10 us per message — hands over the pointers of n filled portion buffers).
// The headers are scanned by the context's worker pool first (scan_header), the calls then run in order on the
// calling thread with the scan results, and the rows of the queued portions are written by the pool again
// (run_fill_tasks) — before a batch goes to the device and before this function returns, so the caller may reuse
// the header buffers afterwards exactly as after n single calls.
// (in groups of kScanGroup portions, so that the first batch is on its way after a fraction of a millisecond)
constexpr int kScanGroup = 512;
static void scan_headers(ycnr_ctx* c, int kind, int32_t from, int32_t to, const int32_t* const* rows, int64_t lim_rows,
                         std::vector<PortionPre>& pre) {
  const double t0 = c->trace ? now_ms() : 0.0;
  constexpr int kPer = 8;
  const int n = to - from;
  const std::function<void(int)> fn = [&](int blk) {
    const int i1 = std::min<int>(to, from + (blk + 1) * kPer);
    for (int i = from + blk * kPer; i < i1; ++i)
      if (rows[i]) pre[i] = scan_header(kind, rows[i], lim_rows);
  };
  c->pool.run((n + kPer - 1) / kPer, fn);
  if (c->trace) c->t_scan += now_ms() - t0;
}

int ycnr_als_portions(ycnr_ctx* c, int32_t n, const int32_t* const* rows, const int32_t* const* indx,
                      const float* const* vals, ycnr_portion_info* infos) {
  if (!c || n < 0 || (n && (!rows || !indx || !vals))) return fail("ycnr_als_portions: bad argument");
  if (c->step_type != YCNR_BY_USER && c->step_type != YCNR_BY_ITEM)
    return fail("ycnr_als_portion: call ycnr_start_train_step first");
  const int solved_w = c->step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
  std::vector<PortionPre> pre((size_t)n);
  int rc = 0;
  for (int g0 = 0; g0 < n && !rc; g0 += kScanGroup) {
    const int g1 = std::min<int>(n, g0 + kScanGroup);
    scan_headers(c, 1, g0, g1, rows, c->fac_rows[solved_w], pre);
    for (int i = g0; i < g1 && !rc; ++i)
      rc = als_portion_impl(c, rows[i], indx[i], vals[i], infos ? infos + i : nullptr, &pre[i]);
  }
  const int rf = run_fill_tasks(c);
  return rc ? rc : rf;
}

int ycnr_rmse_portions_async(ycnr_ctx* c, int32_t n, const int32_t* const* rows, const int32_t* const* indx,
                             const float* const* vals, const int64_t* tags) {
  if (!c || n < 0 || (n && (!rows || !indx || !vals))) return fail("ycnr_rmse_portions_async: bad argument");
  const double t0 = now_ms();
  std::vector<PortionPre> pre((size_t)n);
  int rc = 0;
  for (int g0 = 0; g0 < n && !rc; g0 += kScanGroup) {
    const int g1 = std::min<int>(n, g0 + kScanGroup);
    scan_headers(c, 2, g0, g1, rows, c->fac_rows[YCNR_USER_FACTORS], pre);
    for (int i = g0; i < g1 && !rc; ++i)
      rc = rmse_portion_async_impl(c, rows[i], indx[i], vals[i], tags ? tags[i] : i, &pre[i]);
  }
  const int rf = run_fill_tasks(c);
  c->t_rmse_calls += now_ms() - t0;
  return rc ? rc : rf;
}

// Compatibility export (upstream never calls it from lib/).  When `fixed` is one of the attached
// factor matrices and its device replica is current, the rows are gathered from the replica and only the ids
// travel up; any other host matrix is gathered row by row with the copy engine (cols copies of k floats) —
// never a full upload of `fixed`.
int ycnr_s_als_build_sub_fixed_facts(ycnr_ctx* c, float* sub, const float* fixed, int64_t fixed_rows,
                                     const int32_t* indx, int32_t cols, int32_t k) {
  if (!c || !sub || !fixed || !indx || cols < 0 || k <= 0) return fail("ycnr_s_als_build_sub_fixed_facts: bad argument");
  if (cols == 0) return 0;
  for (int i = 0; i < cols; ++i)
    if (indx[i] < 0 || indx[i] >= fixed_rows) return fail("ycnr_s_als_build_sub_fixed_facts: index %d out of range", indx[i]);
  OK(set_device(c));
  const size_t sb = (size_t)cols * k * sizeof(float);
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  OK(c->gather_tmp.ensure(al(sb) + al((size_t)cols * sizeof(int32_t))));
  char* d = (char*)c->gather_tmp.p;
  float* d_sub = (float*)d;
  int32_t* d_idx = (int32_t*)(d + al(sb));
  int attached = -1;
  for (int w = 0; w < 2; ++w)
    if (fixed == c->h_fac[w] && k == c->k && fixed_rows == c->fac_rows[w] && c->device_current[w]) attached = w;
  if (attached >= 0) {
    CU(cudaMemcpyAsync(d_idx, indx, (size_t)cols * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    ProfScope ps(c, YCNR_K_GATHER, cols, cols);
    const int64_t total = (int64_t)cols * k;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)c->num_sms * 8);
    ycnr::gather_rows_kernel<<<grid, 256, 0, c->stream>>>(d_sub, c->d_fac[attached], d_idx, cols, k);
    CU(cudaGetLastError());
  } else {
    for (int i = 0; i < cols; ++i)
      CU(cudaMemcpyAsync(d_sub + (size_t)i * k, fixed + (size_t)indx[i] * k, (size_t)k * sizeof(float),
                         cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaMemcpyAsync(sub, d_sub, sb, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ycnr_s_als_build_sub_fixed_facts_noctx(float* sub, const float* fixed, int64_t fixed_rows, const int32_t* indx,
                                           int32_t cols, int32_t k) {
  if (!g_default_ctx) return fail("sAlsBuildSubFixedFacts: no live context in this process (call ycnr_create first)");
  return ycnr_s_als_build_sub_fixed_facts(g_default_ctx, sub, fixed, fixed_rows, indx, cols, k);
}

int ycnr_check_portion(const int32_t* rows, int64_t rows_len, int64_t indx_len, int64_t vals_len) {
  if (!rows || rows_len < 1) return fail("portion: the rows array needs at least the row count");
  const int64_t R = rows[0];
  if (R < 0) return fail("portion header: negative row count");
  if (2 * R + 1 > rows_len) return fail("portion header: %lld rows do not fit the rows array (%lld words)", (long long)R, (long long)rows_len);
  int64_t off = 0;
  for (int64_t r = 0; r < R; ++r) {
    const int32_t n = rows[2 + 2 * r];
    if (n < 0) return fail("portion header: negative cols in row %lld", (long long)r);
    off += n;
  }
  if (off > indx_len || off > vals_len)
    return fail("portion header: %lld ratings do not fit the indx/vals arrays (%lld / %lld)", (long long)off, (long long)indx_len, (long long)vals_len);
  return 0;
}

int ycnr_factor_elems(ycnr_ctx* c, int32_t which, int64_t* out) {
  if (!c || which < 0 || which > 1 || !out) return fail("ycnr_factor_elems: bad argument");
  *out = c->fac_rows[which] * (int64_t)c->k;
  return 0;
}

int ycnr_memory_usage(ycnr_ctx* c, int64_t out[4]) {
  if (!c || !out) return fail("ycnr_memory_usage: bad argument");
  OK(set_device(c));
  size_t dev = 0, pinned = 0;
  for (int w = 0; w < 2; ++w) {
    dev += (size_t)c->fac_rows[w] * c->k * sizeof(float);
    if (c->h_registered[w]) pinned += (size_t)c->fac_rows[w] * c->k * sizeof(float);
  }
  for (auto& rs : c->rowsets) dev += rs.rows.cap + rs.ratings.cap + rs.plan.cap + rs.sums.cap;
  for (auto& sl : c->slots) { dev += sl.dev.cap; pinned += sl.host_cap; }
  dev += c->table.buf.cap + c->ingest_tmp.cap + c->partial.cap + c->gather_tmp.cap;
  for (auto& r : c->pinned) pinned += r.second;
  size_t fr = 0, tot = 0;
  CU(cudaMemGetInfo(&fr, &tot));
  out[0] = (int64_t)dev;
  out[1] = (int64_t)pinned;
  out[2] = (int64_t)fr;
  out[3] = (int64_t)tot;
  return 0;
}

// ---- bulk path --------------------------------------------------------------------------
int ycnr_rowset_create(ycnr_ctx* c, int32_t step_type, int32_t n_rows, const int32_t* row_ids,
                       const int64_t* row_start, const int32_t* row_len, const int32_t* indx,
                       const float* vals, int64_t span, const int32_t* portion_first, int32_t n_portions,
                       int32_t* out) {
  if (!c || !out || n_rows < 0 || span < 0) return fail("ycnr_rowset_create: bad argument");
  if (step_type < YCNR_BY_USER || step_type > YCNR_RMSE_TEST) return fail("ycnr_rowset_create: bad stepType");
  if (n_rows && (!row_ids || !row_start || !row_len)) return fail("ycnr_rowset_create: null row arrays");
  if (span && (!indx || !vals)) return fail("ycnr_rowset_create: null ratings arrays");
  OK(set_device(c));
  const int64_t lim_cols = (step_type == YCNR_BY_ITEM) ? c->fac_rows[0] : c->fac_rows[1];
  const int64_t lim_rows = (step_type == YCNR_BY_ITEM) ? c->fac_rows[1] : c->fac_rows[0];
  int64_t nnz = 0;
  for (int r = 0; r < n_rows; ++r) {
    if (row_ids[r] < 0 || row_ids[r] >= lim_rows) return fail("ycnr_rowset_create: row id %d out of range", row_ids[r]);
    if (row_len[r] < 0 || row_start[r] < 0 || row_start[r] + row_len[r] > span)
      return fail("ycnr_rowset_create: row %d addresses ratings outside [0, span)", r);
    nnz += row_len[r];
  }
  for (int64_t e = 0; e < span; ++e)
    if (indx[e] < 0 || indx[e] >= lim_cols) return fail("ycnr_rowset_create: column id %d out of range", indx[e]);
  int id = -1;
  for (size_t i = 0; i < c->rowsets.size(); ++i)
    if (!c->rowsets[i].used) { id = (int)i; break; }
  if (id < 0) { c->rowsets.emplace_back(); id = (int)c->rowsets.size() - 1; }
  RowSet& rs = c->rowsets[id];
  rs = RowSet();
  rs.used = true;
  rs.step_type = step_type;
  rs.n_rows = n_rows;
  rs.span = span;
  rs.nnz = nnz;
  const bool rmse = step_type >= YCNR_RMSE_VALIDATE;
  int32_t pf_default[2] = {0, n_rows};
  if (!portion_first || n_portions <= 0) { portion_first = pf_default; n_portions = 1; }
  rs.n_portions = n_portions;
  // rows buffer: start[R] i64 | ids[R] | len[R] | pfirst[P+1]
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_ids = al((size_t)n_rows * 8), o_len = al(o_ids + (size_t)n_rows * 4), o_pf = al(o_len + (size_t)n_rows * 4);
  const size_t rows_bytes = al(o_pf + (size_t)(n_portions + 1) * 4);
  OK(rs.rows.ensure(rows_bytes));
  char* d = (char*)rs.rows.p;
  if (n_rows) {
    CU(cudaMemcpyAsync(d, row_start, (size_t)n_rows * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + o_ids, row_ids, (size_t)n_rows * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + o_len, row_len, (size_t)n_rows * 4, cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaMemcpyAsync(d + o_pf, portion_first, (size_t)(n_portions + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  const size_t o_vals = al((size_t)span * 4);
  OK(rs.ratings.ensure(o_vals + al((size_t)span * 4)));
  char* dr = (char*)rs.ratings.p;
  if (span) {
    CU(cudaMemcpyAsync(dr, indx, (size_t)span * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dr + o_vals, vals, (size_t)span * 4, cudaMemcpyHostToDevice, c->stream));
  }
  rs.view.row_start = (const int64_t*)d;
  rs.view.row_ids = (const int32_t*)(d + o_ids);
  rs.view.row_len = (const int32_t*)(d + o_len);
  rs.d_portion_first = (const int32_t*)(d + o_pf);
  rs.view.indx = (const int32_t*)dr;
  rs.view.vals = (const float*)(dr + o_vals);
  std::vector<int32_t> packed;
  if (!rmse) {
    const PlanCfg cfg{c->dual_max, c->split_cols, c->fused_max};
    plan_count(row_len, 1, n_rows, cfg, rs.dplan);
    packed.resize(rs.dplan.words + 1);
    plan_fill(row_len, 1, n_rows, cfg, rs.dplan, packed.data(), c->opts.solve_chunks);
    OK(rs.plan.ensure(packed.size() * 4));
    CU(cudaMemcpyAsync(rs.plan.p, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, c->stream));
  } else {
    OK(build_rmse_entries(c, rs, row_ids, row_start, row_len, portion_first));
  }
  CU(cudaStreamSynchronize(c->stream));  // host sources may be freed by the caller on return
  *out = id;
  return 0;
}

int ycnr_rowset_destroy(ycnr_ctx* c, int32_t id) {
  if (!c || id < 0 || id >= (int)c->rowsets.size() || !c->rowsets[id].used) return fail("ycnr_rowset_destroy: bad id");
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  RowSet& rs = c->rowsets[id];
  rs.rows.release(); rs.ratings.release(); rs.plan.release(); rs.sums.release(); rs.order.release(); rs.erows.release();
  if (rs.h_sums) cudaFreeHost(rs.h_sums);
  if (rs.sums_ready) cudaEventDestroy(rs.sums_ready);
  rs.h_sums = nullptr;
  rs.sums_ready = nullptr;
  rs.pending = rs.cached = false;
  rs.used = false;
  return 0;
}

int ycnr_als_rowset(ycnr_ctx* c, int32_t id) {
  if (!c || id < 0 || id >= (int)c->rowsets.size() || !c->rowsets[id].used) return fail("ycnr_als_rowset: bad id");
  RowSet& rs = c->rowsets[id];
  if (rs.step_type != YCNR_BY_USER && rs.step_type != YCNR_BY_ITEM) return fail("ycnr_als_rowset: row set is an RMSE set");
  OK(set_device(c));
  const int solved = rs.step_type == YCNR_BY_USER ? YCNR_USER_FACTORS : YCNR_ITEM_FACTORS;
  OK(ensure_fixed_current(c, 1 - solved));
  OK(ensure_fixed_current(c, solved));
  c->fac_version[solved]++;   // (a rank without rows still receives its peers' rows)
  if (rs.n_rows == 0) return 0;
  const bool spread = c->spread_bulk < 0 ? rs.n_rows < kSpreadBulkRows : c->spread_bulk != 0;
  return run_als(c, rs.step_type, rs.view, rs.dplan, (const int32_t*)rs.plan.p, spread);
}

// Queue the RMSE pass of a row set at `shift` (kernels + the copy of the per-portion sums), without waiting.
// No-op when the sums of a pass over the same factor versions are already there or on their way.
int ycnr_rmse_rowset_begin(ycnr_ctx* c, int32_t id, double shift) {
  if (!c || id < 0 || id >= (int)c->rowsets.size() || !c->rowsets[id].used) return fail("ycnr_rmse_rowset_begin: bad argument");
  RowSet& rs = c->rowsets[id];
  if (rs.step_type != YCNR_RMSE_VALIDATE && rs.step_type != YCNR_RMSE_TEST) return fail("ycnr_rmse_rowset: row set is an ALS set");
  OK(set_device(c));
  OK(ensure_fixed_current(c, YCNR_USER_FACTORS));
  OK(ensure_fixed_current(c, YCNR_ITEM_FACTORS));
  if (rs.n_rows == 0) return 0;
  const bool fresh = rs.cache_ver[0] == c->fac_version[0] && rs.cache_ver[1] == c->fac_version[1];
  if ((rs.cached || rs.pending) && fresh) return 0;
  if (rs.pending) {   // a pass over older factors is still in flight: let it land before its buffers are reused
    CU(cudaEventSynchronize(rs.sums_ready));
    rs.pending = false;
  }
  if (!rs.h_sums) CU(cudaMallocHost(&rs.h_sums, sizeof(double) * 4 * rs.n_portions));
  if (!rs.sums_ready) CU(cudaEventCreateWithFlags(&rs.sums_ready, cudaEventDisableTiming));
  double* d_rows = (double*)rs.sums.p;
  double* d_port = d_rows + 3 * (size_t)rs.n_entries;
  OK(run_rmse(c, rs.eview, rs.n_entries, rs.nnz, shift, d_rows, rs.d_efirst, rs.n_portions, d_port, nullptr,
              (const int32_t*)rs.order.p));
  CU(cudaMemcpyAsync(rs.h_sums, d_port, sizeof(double) * 4 * rs.n_portions, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaEventRecord(rs.sums_ready, c->stream));
  rs.pending = true;
  rs.cached = false;
  rs.cache_shift = shift;
  rs.cache_ver[0] = c->fac_version[0];
  rs.cache_ver[1] = c->fac_version[1];
  return 0;
}

int ycnr_rmse_rowset(ycnr_ctx* c, int32_t id, double shift, double* totals, double* portion_sums) {
  if (!c || !totals || id < 0 || id >= (int)c->rowsets.size() || !c->rowsets[id].used) return fail("ycnr_rmse_rowset: bad argument");
  OK(ycnr_rmse_rowset_begin(c, id, shift));
  RowSet& rs = c->rowsets[id];
  totals[0] = totals[1] = totals[2] = 0.0;
  if (rs.n_rows == 0) {
    if (portion_sums) memset(portion_sums, 0, sizeof(double) * 3 * rs.n_portions);
    return 0;
  }
  if (rs.pending) {
    CU(cudaEventSynchronize(rs.sums_ready));
    rs.pending = false;
    rs.cached = true;
  }
  // the pass at `shift` from the sums at cache_shift: pred' = pred + d for every rating
  const double d = shift - rs.cache_shift;
  for (int p = 0; p < rs.n_portions; ++p) {  // EmfMaster.m_completedPortion 770-774, portion order
    const double* q = rs.h_sums + 4 * (size_t)p;
    const double d2 = d == 0.0 ? q[0] : q[0] - 2.0 * d * (q[3] - q[2]) + q[1] * d * d;
    const double sp = d == 0.0 ? q[2] : q[2] + q[1] * d;
    if (portion_sums) {
      portion_sums[3 * p] = d2;
      portion_sums[3 * p + 1] = q[1];
      portion_sums[3 * p + 2] = sp;
    }
    totals[0] += d2;
    totals[1] += q[1];
    totals[2] += sp;
  }
  return 0;
}

// Sum of the ratings themselves over the row set and over its last portion (from the sums of the last pass):
// with them a host layer that has gathered the pass at shift 0 from all ranks derives the pass at any other shift.
int ycnr_rmse_rowset_ratings(ycnr_ctx* c, int32_t id, double* total_out, double* last_portion_out) {
  if (!c || !total_out || id < 0 || id >= (int)c->rowsets.size() || !c->rowsets[id].used) return fail("ycnr_rmse_rowset_ratings: bad argument");
  RowSet& rs = c->rowsets[id];
  *total_out = 0.0;
  if (last_portion_out) *last_portion_out = 0.0;
  if (rs.n_rows == 0) return 0;
  if (rs.pending) {
    CU(cudaEventSynchronize(rs.sums_ready));
    rs.pending = false;
    rs.cached = true;
  }
  if (!rs.cached) return fail("ycnr_rmse_rowset_ratings: no RMSE pass has run on this row set yet");
  for (int p = 0; p < rs.n_portions; ++p) *total_out += rs.h_sums[4 * (size_t)p + 3];
  if (last_portion_out) *last_portion_out = rs.h_sums[4 * (size_t)(rs.n_portions - 1) + 3];
  return 0;
}

// ---- multi-GPU ------------------------------------------------------------------------------
int ycnr_ipc_export(ycnr_ctx* c, int32_t which, uint8_t handle_out[64]) {
  if (!c || which < 0 || which > 1 || !handle_out) return fail("ycnr_ipc_export: bad argument");
  OK(set_device(c));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, c->d_fac[which]));
  memcpy(handle_out, &h, 64);
  return 0;
}

int ycnr_ipc_import(ycnr_ctx* c, const uint8_t handle[64], void** out) {
  if (!c || !handle || !out) return fail("ycnr_ipc_import: bad argument");
  OK(set_device(c));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int ycnr_ipc_close(ycnr_ctx* c, void* p) {
  if (!c || !p) return fail("ycnr_ipc_close: bad argument");
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaIpcCloseMemHandle(p));
  return 0;
}

int ycnr_set_peers(ycnr_ctx* c, int32_t which, int32_t n, void* const* ptrs) {
  if (!c || which < 0 || which > 1 || n < 0 || n > YCNR_MAX_DST - 1 || (n && !ptrs)) return fail("ycnr_set_peers: bad argument");
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  c->peers[which].assign(ptrs, ptrs + n);
  return 0;
}

// ---- device-side front end (ingest) ------------------------------------------------------------
namespace {

// exclusive scan of n int32 values into out[n + 1] (int64); scratch: block sums
int device_scan(ycnr_ctx* c, const int32_t* d_v, int n, int64_t* d_out, int64_t* d_block_sums) {
  if (n <= 0) {
    CU(cudaMemsetAsync(d_out, 0, sizeof(int64_t), c->stream));
    return 0;
  }
  const int nb = (n + ycnr::kScanRowsPerBlock - 1) / ycnr::kScanRowsPerBlock;
  ycnr::scan_block_sums_kernel<<<nb, ycnr::kScanThreads, 0, c->stream>>>(d_v, n, d_block_sums);
  ycnr::header_scan_blocks_kernel<<<1, 1024, 0, c->stream>>>(d_block_sums, nb);
  ycnr::scan_apply_kernel<<<nb, ycnr::kScanThreads, 0, c->stream>>>(d_v, n, d_block_sums, d_out);
  CU(cudaGetLastError());
  c->prof.launches[YCNR_K_GATHER] += 3;
  c->prof.total_launches += 3;
  return 0;
}

size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace

int ycnr_table_upload(ycnr_ctx* c, const int64_t* user_ptr, const int32_t* item_ids, const float* ratings,
                      const int8_t* dataset_type) {
  if (!c || !user_ptr || !item_ids || !ratings || !dataset_type) return fail("ycnr_table_upload: null argument");
  const int users = (int)c->fac_rows[0], items = (int)c->fac_rows[1];
  if (user_ptr[0] != 0) return fail("ycnr_table_upload: user_ptr[0] must be 0");
  for (int u = 0; u < users; ++u)
    if (user_ptr[u + 1] < user_ptr[u]) return fail("ycnr_table_upload: user_ptr not monotone at %d", u);
  const int64_t nnz = user_ptr[users];
  for (int64_t e = 0; e < nnz; ++e) {
    if (item_ids[e] < 0 || item_ids[e] >= items) return fail("ycnr_table_upload: item id %d out of range", item_ids[e]);
    if (dataset_type[e] < 0 || dataset_type[e] > 31) return fail("ycnr_table_upload: dataset_type %d out of range", (int)dataset_type[e]);
  }
  OK(set_device(c));
  const size_t o_item = al16((size_t)(users + 1) * 8), o_rat = al16(o_item + (size_t)nnz * 4);
  const size_t o_dt = al16(o_rat + (size_t)nnz * 4), o_eu = al16(o_dt + (size_t)nnz), total = al16(o_eu + (size_t)nnz * 4);
  OK(c->table.buf.ensure(total));
  char* d = (char*)c->table.buf.p;
  CU(cudaMemcpyAsync(d, user_ptr, (size_t)(users + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (nnz) {
    CU(cudaMemcpyAsync(d + o_item, item_ids, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + o_rat, ratings, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + o_dt, dataset_type, (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  }
  c->table.nnz = nnz;
  c->table.user_ptr = (const int64_t*)d;
  c->table.item = (const int32_t*)(d + o_item);
  c->table.rating = (const float*)(d + o_rat);
  c->table.dt = (const int8_t*)(d + o_dt);
  c->table.elem_user = (const int32_t*)(d + o_eu);
  {
    ProfScope ps(c, YCNR_K_GATHER, users, nnz);
    ycnr::fill_elem_user_kernel<<<(users + 7) / 8, 256, 0, c->stream>>>(c->table.user_ptr, users, (int32_t*)(d + o_eu));
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));   // host sources may be freed by the caller on return
  c->table.loaded = true;
  return 0;
}

int ycnr_table_split(ycnr_ctx* c, uint64_t seed, const int32_t pcts[3], int8_t* dataset_type_out) {
  if (!c || !pcts) return fail("ycnr_table_split: null argument");
  if (!c->table.loaded) return fail("ycnr_table_split: call ycnr_table_upload first");
  if (pcts[0] < 0 || pcts[1] < 0 || pcts[2] < 0 || pcts[0] + pcts[1] + pcts[2] != 100)
    return fail("ycnr_table_split: dataSetDistr must be three percentages summing to 100");
  OK(set_device(c));
  const int users = (int)c->fac_rows[0];
  OK(c->ingest_tmp.ensure((size_t)std::max<int64_t>(c->table.nnz, 1) * 4));
  {
    ProfScope ps(c, YCNR_K_GATHER, users, c->table.nnz);
    ycnr::split_sets_kernel<<<(users + 127) / 128, 128, 0, c->stream>>>(seed, users, c->table.user_ptr, pcts[0], pcts[1],
                                                                      (int32_t*)c->ingest_tmp.p, (int8_t*)c->table.dt);
  }
  CU(cudaGetLastError());
  if (dataset_type_out && c->table.nnz)
    CU(cudaMemcpyAsync(dataset_type_out, c->table.dt, (size_t)c->table.nnz, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ycnr_table_counts(ycnr_ctx* c, uint32_t set_mask, int32_t by_item, int32_t* counts_out) {
  if (!c || !counts_out) return fail("ycnr_table_counts: null argument");
  if (!c->table.loaded) return fail("ycnr_table_counts: call ycnr_table_upload first");
  OK(set_device(c));
  const int users = (int)c->fac_rows[0], items = (int)c->fac_rows[1];
  const int n = by_item ? items : users;
  OK(c->ingest_tmp.ensure((size_t)n * 4));
  int32_t* d_cnt = (int32_t*)c->ingest_tmp.p;
  {
    ProfScope ps(c, YCNR_K_GATHER, n, c->table.nnz);
    if (by_item) {
      CU(cudaMemsetAsync(d_cnt, 0, (size_t)n * 4, c->stream));
      const int grid = (int)std::min<int64_t>((c->table.nnz + 255) / 256 + 1, (int64_t)c->num_sms * 16);
      ycnr::count_by_item_kernel<<<grid, 256, 0, c->stream>>>(c->table.item, c->table.dt, set_mask, c->table.nnz, d_cnt);
    } else {
      ycnr::count_by_user_kernel<<<(users + 7) / 8, 256, 0, c->stream>>>(c->table.user_ptr, c->table.dt, set_mask, users, d_cnt);
    }
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(counts_out, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ycnr_rowset_from_table(ycnr_ctx* c, int32_t step_type, uint32_t set_mask, int32_t first_row,
                           const int32_t* portions_row_id_to, int32_t n_portions, int32_t* out) {
  if (!c || !out || n_portions < 0 || (n_portions && !portions_row_id_to)) return fail("ycnr_rowset_from_table: bad argument");
  if (step_type < YCNR_BY_USER || step_type > YCNR_RMSE_TEST) return fail("ycnr_rowset_from_table: bad stepType");
  if (!c->table.loaded) return fail("ycnr_rowset_from_table: call ycnr_table_upload first");
  OK(set_device(c));
  const int users = (int)c->fac_rows[0], items = (int)c->fac_rows[1];
  const bool by_item = step_type == YCNR_BY_ITEM;
  const int rows = by_item ? items : users;
  if (first_row < 0 || first_row > rows) return fail("ycnr_rowset_from_table: first_row outside [0, rows]");
  for (int p = 0; p < n_portions; ++p) {
    const int lo = p ? portions_row_id_to[p - 1] : first_row;
    if (portions_row_id_to[p] < lo || portions_row_id_to[p] > rows) return fail("ycnr_rowset_from_table: portion bounds not monotone in [first_row, rows]");
  }
  const int64_t nnz_t = c->table.nnz;
  const int P = std::max(n_portions, 1);
  // by item: per-chunk counters of n_chunks x items words.  The chunk grows with the catalog so that the counters
  // stay under 256 MB (100 M ratings x 1 M items would otherwise ask for 12 GB); the sort stays stable for any chunk.
  int64_t item_chunk = ycnr::kItemChunk;
  while (by_item && ((nnz_t + item_chunk - 1) / item_chunk) * (int64_t)items * 4 > ((int64_t)256 << 20)) item_chunk *= 2;
  const int n_chunks = (int)((nnz_t + item_chunk - 1) / item_chunk);
  // scratch: cnt i32[rows] | ptr i64[rows+1] | block sums | pto | last_row | drop_last | flag | len | pos i64[rows+1]
  //          | (by item) chunk counters i32[n_chunks][items]
  const int nb = (rows + ycnr::kScanRowsPerBlock - 1) / ycnr::kScanRowsPerBlock + 2;
  size_t o = 0;
  const size_t o_cnt = o;   o = al16(o + (size_t)rows * 4);
  const size_t o_ptr = o;   o = al16(o + (size_t)(rows + 1) * 8);
  const size_t o_bs = o;    o = al16(o + (size_t)nb * 8);
  const size_t o_pto = o;   o = al16(o + (size_t)P * 4);
  const size_t o_last = o;  o = al16(o + (size_t)P * 4);
  const size_t o_drop = o;  o = al16(o + (size_t)P * 4);
  const size_t o_flag = o;  o = al16(o + (size_t)rows * 4);
  const size_t o_len = o;   o = al16(o + (size_t)rows * 4);
  const size_t o_pos = o;   o = al16(o + (size_t)(rows + 1) * 8);
  const size_t o_cur = o;   o = al16(o + (by_item ? (size_t)n_chunks * items * 4 : 0));
  OK(c->ingest_tmp.ensure(o));
  char* t = (char*)c->ingest_tmp.p;
  int32_t* d_cnt = (int32_t*)(t + o_cnt);
  int64_t* d_ptr = (int64_t*)(t + o_ptr);
  int64_t* d_bs = (int64_t*)(t + o_bs);
  int32_t* d_pto = (int32_t*)(t + o_pto);
  int32_t* d_last = (int32_t*)(t + o_last);
  int32_t* d_drop = (int32_t*)(t + o_drop);
  int32_t* d_flag = (int32_t*)(t + o_flag);
  int32_t* d_len = (int32_t*)(t + o_len);
  int64_t* d_pos = (int64_t*)(t + o_pos);
  int32_t* d_cur = (int32_t*)(t + o_cur);
  if (n_portions) CU(cudaMemcpyAsync(d_pto, portions_row_id_to, (size_t)n_portions * 4, cudaMemcpyHostToDevice, c->stream));

  ProfScope ps(c, YCNR_K_GATHER, rows, nnz_t);
  // 1. row lengths of the fetch and its row pointer
  if (by_item) {
    CU(cudaMemsetAsync(d_cur, 0, (size_t)n_chunks * items * 4, c->stream));
    if (n_chunks) ycnr::item_hist_kernel<<<n_chunks, 256, 0, c->stream>>>(c->table.item, c->table.dt, set_mask, nnz_t, items, item_chunk, d_cur);
    ycnr::item_chunk_offsets_kernel<<<(items + 255) / 256, 256, 0, c->stream>>>(d_cur, n_chunks, items, d_cnt);
  } else {
    ycnr::count_by_user_kernel<<<(users + 7) / 8, 256, 0, c->stream>>>(c->table.user_ptr, c->table.dt, set_mask, users, d_cnt);
  }
  CU(cudaGetLastError());
  OK(device_scan(c, d_cnt, rows, d_ptr, d_bs));
  int64_t span = 0;
  CU(cudaMemcpyAsync(&span, d_ptr + rows, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));

  int id = -1;
  for (size_t i = 0; i < c->rowsets.size(); ++i)
    if (!c->rowsets[i].used) { id = (int)i; break; }
  if (id < 0) { c->rowsets.emplace_back(); id = (int)c->rowsets.size() - 1; }
  RowSet& rs = c->rowsets[id];
  rs = RowSet();
  rs.used = true;
  rs.step_type = step_type;
  rs.span = span;
  rs.n_portions = P;
  // 2. the fetch itself
  const size_t o_vals = al16((size_t)span * 4);
  OK(rs.ratings.ensure(o_vals + al16((size_t)span * 4) + 16));
  char* dr = (char*)rs.ratings.p;
  if (by_item) {
    if (n_chunks)
      ycnr::item_scatter_kernel<<<n_chunks, 32, 0, c->stream>>>(c->table.item, c->table.rating, c->table.dt, c->table.elem_user,
                                                               set_mask, nnz_t, items, item_chunk, d_cur, d_ptr, (int32_t*)dr, (float*)(dr + o_vals));
  } else {
    ycnr::fill_by_user_kernel<<<(users + 7) / 8, 256, 0, c->stream>>>(c->table.user_ptr, c->table.item, c->table.rating, c->table.dt,
                                                                     set_mask, users, d_ptr, (int32_t*)dr, (float*)(dr + o_vals));
  }
  CU(cudaGetLastError());
  // 3. portion headers with quirk Q2
  int32_t pto_all[1] = {rows};
  if (!n_portions) CU(cudaMemcpyAsync(d_pto, pto_all, 4, cudaMemcpyHostToDevice, c->stream));   // one portion: everything
  ycnr::portion_tail_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(d_ptr, d_pto, P, first_row, d_last, d_drop);
  ycnr::row_emit_kernel<<<(rows + 255) / 256, 256, 0, c->stream>>>(d_ptr, rows, d_pto, P, first_row, d_last, d_drop, d_flag, d_len);
  CU(cudaGetLastError());
  OK(device_scan(c, d_flag, rows, d_pos, d_bs));
  int64_t n_rows64 = 0;
  CU(cudaMemcpyAsync(&n_rows64, d_pos + rows, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const int n_rows = (int)n_rows64;
  rs.n_rows = n_rows;
  const size_t o_ids = al16((size_t)n_rows * 8), o_rl = al16(o_ids + (size_t)n_rows * 4), o_pf = al16(o_rl + (size_t)n_rows * 4);
  OK(rs.rows.ensure(al16(o_pf + (size_t)(P + 1) * 4) + 16));
  char* d = (char*)rs.rows.p;
  ycnr::row_scatter_kernel<<<(rows + 255) / 256, 256, 0, c->stream>>>(d_ptr, rows, d_flag, d_len, d_pos, (int32_t*)(d + o_ids),
                                                                    (int64_t*)d, (int32_t*)(d + o_rl));
  ycnr::portion_first_kernel<<<(P + 1 + 255) / 256, 256, 0, c->stream>>>(d_pos, rows, d_pto, P, first_row, (int32_t*)(d + o_pf));
  CU(cudaGetLastError());
  c->prof.launches[YCNR_K_GATHER] += 6;
  c->prof.total_launches += 6;
  rs.view.row_start = (const int64_t*)d;
  rs.view.row_ids = (const int32_t*)(d + o_ids);
  rs.view.row_len = (const int32_t*)(d + o_rl);
  rs.d_portion_first = (const int32_t*)(d + o_pf);
  rs.view.indx = (const int32_t*)dr;
  rs.view.vals = (const float*)(dr + o_vals);
  // 4. the launch plan is built on the host from the emitted row lengths (one small copy each way)
  std::vector<int32_t> h_len((size_t)std::max(n_rows, 1));
  if (n_rows) CU(cudaMemcpyAsync(h_len.data(), d + o_rl, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  int64_t nnz = 0;
  for (int r = 0; r < n_rows; ++r) nnz += h_len[r];
  rs.nnz = nnz;
  if (step_type < YCNR_RMSE_VALIDATE) {
    const PlanCfg cfg{c->dual_max, c->split_cols, c->fused_max};
    plan_count(h_len.data(), 1, n_rows, cfg, rs.dplan);
    std::vector<int32_t> packed(rs.dplan.words + 1);
    plan_fill(h_len.data(), 1, n_rows, cfg, rs.dplan, packed.data(), c->opts.solve_chunks);
    OK(rs.plan.ensure(packed.size() * 4));
    CU(cudaMemcpyAsync(rs.plan.p, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  } else {
    std::vector<int32_t> h_ids((size_t)std::max(n_rows, 1)), h_pf((size_t)P + 1);
    std::vector<int64_t> h_start((size_t)std::max(n_rows, 1));
    if (n_rows) {
      CU(cudaMemcpyAsync(h_ids.data(), d + o_ids, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaMemcpyAsync(h_start.data(), d, (size_t)n_rows * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(h_pf.data(), d + o_pf, (size_t)(P + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    OK(build_rmse_entries(c, rs, h_ids.data(), h_start.data(), h_len.data(), h_pf.data()));
  }
  *out = id;
  return 0;
}

int ycnr_rowset_info(ycnr_ctx* c, int32_t id, int32_t* n_rows, int64_t* span, int32_t* n_portions) {
  if (!c || id < 0 || id >= (int)c->rowsets.size() || !c->rowsets[id].used) return fail("ycnr_rowset_info: bad id");
  const RowSet& rs = c->rowsets[id];
  if (n_rows) *n_rows = rs.n_rows;
  if (span) *span = rs.span;
  if (n_portions) *n_portions = rs.n_portions;
  return 0;
}

int ycnr_rowset_read(ycnr_ctx* c, int32_t id, int32_t* row_ids, int64_t* row_start, int32_t* row_len,
                     int32_t* portion_first, int32_t* indx, float* vals) {
  if (!c || id < 0 || id >= (int)c->rowsets.size() || !c->rowsets[id].used) return fail("ycnr_rowset_read: bad id");
  const RowSet& rs = c->rowsets[id];
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  if (rs.n_rows) {
    if (row_ids) CU(cudaMemcpy(row_ids, rs.view.row_ids, (size_t)rs.n_rows * 4, cudaMemcpyDeviceToHost));
    if (row_start) CU(cudaMemcpy(row_start, rs.view.row_start, (size_t)rs.n_rows * 8, cudaMemcpyDeviceToHost));
    if (row_len) CU(cudaMemcpy(row_len, rs.view.row_len, (size_t)rs.n_rows * 4, cudaMemcpyDeviceToHost));
  }
  if (portion_first) CU(cudaMemcpy(portion_first, rs.d_portion_first, (size_t)(rs.n_portions + 1) * 4, cudaMemcpyDeviceToHost));
  if (rs.span) {
    if (indx) CU(cudaMemcpy(indx, rs.view.indx, (size_t)rs.span * 4, cudaMemcpyDeviceToHost));
    if (vals) CU(cudaMemcpy(vals, rs.view.vals, (size_t)rs.span * 4, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ---- serving ----------------------------------------------------------------------------------
int ycnr_recommend_batch(ycnr_ctx* c, int32_t n_users, const int32_t* user_ids, const int64_t* skip_ptr,
                         const int32_t* skip_ids, int32_t limit, double min_recommend_rating,
                         double global_avg_shift, int32_t* out_item_ids, double* out_predict, int32_t* out_count) {
  if (!c || n_users < 0 || limit < 1 || (n_users && (!user_ids || !skip_ptr || !out_count)))
    return fail("ycnr_recommend_batch: bad argument");
  const int keep = limit - 1;   // YcnrController.js:281-282
  if (keep > 0 && n_users && (!out_item_ids || !out_predict)) return fail("ycnr_recommend_batch: null output");
  for (int u = 0; u < n_users; ++u) {
    if (user_ids[u] < 0 || user_ids[u] >= c->fac_rows[0]) return fail("ycnr_recommend_batch: user id %d out of range", user_ids[u]);
    if (skip_ptr[u + 1] < skip_ptr[u]) return fail("ycnr_recommend_batch: skip_ptr not monotone");
  }
  if (n_users && skip_ptr[n_users] > skip_ptr[0] && !skip_ids) return fail("ycnr_recommend_batch: null skip_ids");
  if (n_users == 0) return 0;
  if (keep == 0) { memset(out_count, 0, sizeof(int32_t) * n_users); return 0; }
  OK(set_device(c));
  OK(ensure_fixed_current(c, YCNR_USER_FACTORS));
  OK(ensure_fixed_current(c, YCNR_ITEM_FACTORS));
  const int n_items = (int)c->fac_rows[1];
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const int chunk = std::max(1, std::min(n_users, (int)(((size_t)512 << 20) / ((size_t)n_items * 8))));   // <= 512 MB of scratch
  for (int u0 = 0; u0 < n_users; u0 += chunk) {
    const int nu = std::min(chunk, n_users - u0);
    const int64_t s0 = skip_ptr[u0], ns = skip_ptr[u0 + nu] - s0;
    std::vector<int64_t> rel(nu + 1);
    for (int u = 0; u <= nu; ++u) rel[u] = skip_ptr[u0 + u] - s0;
    const size_t o_ptr = 0, o_uid = al(o_ptr + (size_t)(nu + 1) * 8), o_skip = al(o_uid + (size_t)nu * 4);
    const size_t o_cnt = al(o_skip + (size_t)ns * 4), o_oid = al(o_cnt + (size_t)nu * 4);
    const size_t o_opr = al(o_oid + (size_t)nu * keep * 4), o_pred = al(o_opr + (size_t)nu * keep * 8);
    OK(c->gather_tmp.ensure(al(o_pred + (size_t)nu * n_items * 8)));
    char* d = (char*)c->gather_tmp.p;
    CU(cudaMemcpyAsync(d + o_ptr, rel.data(), (size_t)(nu + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + o_uid, user_ids + u0, (size_t)nu * 4, cudaMemcpyHostToDevice, c->stream));
    if (ns) CU(cudaMemcpyAsync(d + o_skip, skip_ids + s0, (size_t)ns * 4, cudaMemcpyHostToDevice, c->stream));
    ycnr::RecommendArgs a{};
    a.U = c->d_fac[YCNR_USER_FACTORS];
    a.V = c->d_fac[YCNR_ITEM_FACTORS];
    a.k = c->k;
    a.n_items = n_items;
    a.n_users = nu;
    a.user_ids = (const int32_t*)(d + o_uid);
    a.skip_ptr = (const int64_t*)(d + o_ptr);
    a.skip_ids = (const int32_t*)(d + o_skip);
    a.shift = global_avg_shift;
    a.min_rating = min_recommend_rating;
    a.keep = keep;
    a.pred = (double*)(d + o_pred);
    a.out_ids = (int32_t*)(d + o_oid);
    a.out_pred = (double*)(d + o_opr);
    a.out_count = (int32_t*)(d + o_cnt);
    {
      ProfScope ps(c, YCNR_K_GATHER, nu, (int64_t)nu * n_items);
      dim3 grid((n_items + ycnr::kRecItemsPerCta - 1) / ycnr::kRecItemsPerCta, nu);
      ycnr::recommend_scores_kernel<<<grid, 256, 0, c->stream>>>(a);
      ycnr::recommend_select_kernel<<<nu, 256, 0, c->stream>>>(a);
      c->prof.launches[YCNR_K_GATHER] += 1;
      c->prof.total_launches += 1;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_count + u0, d + o_cnt, (size_t)nu * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(out_item_ids + (size_t)u0 * keep, d + o_oid, (size_t)nu * keep * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(out_predict + (size_t)u0 * keep, d + o_opr, (size_t)nu * keep * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

// ---- diagnostics -----------------------------------------------------------------------------
int ycnr_debug_plan(const ycnr_options* o, const int32_t* row_len, int32_t n_rows, int32_t* summary, int32_t* words_out,
                    int64_t cap_words, int64_t* n_words_out) {
  if (!o || (n_rows && !row_len) || n_rows < 0 || !summary) return fail("ycnr_debug_plan: bad argument");
  const PlanCfg cfg = derive_plan_cfg(*o, nullptr);
  DevPlan p;
  plan_count(row_len, 1, n_rows, cfg, p);
  for (int b = 0; b < kDualBins; ++b) summary[b] = p.n_dual[b];
  summary[24] = p.n_fused;
  summary[25] = p.n_multi;
  summary[26] = p.n_items;
  summary[27] = (int32_t)p.off_fused;
  summary[28] = (int32_t)p.off_multi;
  summary[29] = (int32_t)p.off_item_row;
  summary[30] = (int32_t)p.off_item_off;
  summary[31] = (int32_t)p.off_item_order;
  if (n_words_out) *n_words_out = (int64_t)p.words;
  if (words_out) {
    if (cap_words < (int64_t)p.words) return fail("ycnr_debug_plan: %lld words needed, %lld given", (long long)p.words, (long long)cap_words);
    plan_fill(row_len, 1, n_rows, cfg, p, words_out, o->solve_chunks);
  }
  return 0;
}

// The row arrays the multi-portion entry points build for a batch of portion headers (host code only, no GPU):
// scan (entries, ratings, validity per portion) by the worker pool exactly as ycnr_als_portions /
// ycnr_rmse_portions_async do it, then the deferred fill.  kind 1 = ALS rows, 2 = RMSE work entries of at most 64
// ratings.  counts_out[3 n]: entries, ratings, bad flag per portion; the arrays receive the concatenated batch
// (capacity cap entries; portions that are invalid or do not fit are skipped and leave their counts behind).
int ycnr_debug_batch_rows(int32_t kind, int32_t n, const int32_t* const* rows, int64_t lim_rows, int32_t threads,
                          int64_t* counts_out, int32_t* ids_out, int32_t* len_out, int64_t* start_out, int64_t cap,
                          int64_t* entries_out) {
  if ((kind != 1 && kind != 2) || n < 0 || (n && !rows) || !counts_out || !entries_out) return fail("ycnr_debug_batch_rows: bad argument");
  WorkPool pool;
  pool.start(std::max(0, threads - 1));
  std::vector<PortionPre> pre((size_t)n);
  constexpr int kPer = 8;
  const std::function<void(int)> scan = [&](int blk) {
    const int i1 = std::min<int>(n, (blk + 1) * kPer);
    for (int i = blk * kPer; i < i1; ++i)
      if (rows[i]) pre[i] = scan_header(kind, rows[i], lim_rows);
  };
  pool.run((n + kPer - 1) / kPer, scan);
  std::vector<FillTask> tasks;
  int64_t e = 0, run = 0;
  for (int i = 0; i < n; ++i) {
    counts_out[3 * i] = pre[i].entries;
    counts_out[3 * i + 1] = pre[i].ratings;
    counts_out[3 * i + 2] = pre[i].bad;
    if (pre[i].bad || !rows[i] || e + pre[i].entries > cap) continue;
    tasks.push_back({kind, rows[i], (size_t)e, run});
    e += pre[i].entries;
    run += pre[i].ratings;
  }
  if (ids_out && len_out && start_out) {
    const std::function<void(int)> fill = [&](int i) {
      const FillTask& t = tasks[i];
      fill_header(t.kind, t.rows, ids_out + t.r0, len_out + t.r0, start_out + t.r0, t.base);
    };
    pool.run((int)tasks.size(), fill);
  }
  pool.shutdown();
  *entries_out = e;
  return 0;
}

int ycnr_debug_read_partials(ycnr_ctx* c, float* out, int64_t n_floats) {
  if (!c || !out || n_floats < 0) return fail("ycnr_debug_read_partials: bad argument");
  OK(set_device(c));
  if ((size_t)n_floats * sizeof(float) > c->partial.cap) return fail("ycnr_debug_read_partials: only %zu bytes of partials exist", c->partial.cap);
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(out, c->partial.p, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- measurement -----------------------------------------------------------------------------
int ycnr_profile_reset(ycnr_ctx* c) {
  if (!c) return fail("ycnr_profile_reset: null context");
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  collect_profile(c);
  memset(&c->prof, 0, sizeof(c->prof));
  memset(c->dual_bin_ms, 0, sizeof(c->dual_bin_ms));
  memset(c->dual_bin_rows, 0, sizeof(c->dual_bin_rows));
  return 0;
}

int ycnr_profile_dual_bins(ycnr_ctx* c, double* ms_out, int64_t* rows_out) {
  if (!c || !ms_out || !rows_out) return fail("ycnr_profile_dual_bins: bad argument");
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  collect_profile(c);
  memcpy(ms_out, c->dual_bin_ms, sizeof(c->dual_bin_ms));
  memcpy(rows_out, c->dual_bin_rows, sizeof(c->dual_bin_rows));
  return 0;
}

int ycnr_profile_read(ycnr_ctx* c, ycnr_profile* out) {
  if (!c || !out) return fail("ycnr_profile_read: bad argument");
  OK(set_device(c));
  CU(cudaStreamSynchronize(c->stream));
  collect_profile(c);
  *out = c->prof;
  return 0;
}

}  // extern "C"
