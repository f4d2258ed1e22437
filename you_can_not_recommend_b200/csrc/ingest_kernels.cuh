// ingest_kernels.cuh — device-side front end of a training step, sm_100a.
//
// Upstream, the master fetches every portion from PostgreSQL and converts it per rating in JS
// (EmfMaster.m_fetchPortionTrainAlsOrRmse / m_processFetchedPortionAlsOrRmse, lib/emf/EmfMaster.js:501-614).
// Here the ratings table (malrec_ratings, data/db-schema.sql:887-893: user_list_id, item_id, rating,
// dataset_type; sorted by user, item) is uploaded once and these kernels build, per step type:
//   * the fetch: ratings with dataset_type in the step's set (EmfMaster.js:502-503), grouped by user in item
//     order (ORDER BY 511-518) or by item in user order (520-529) — a stream compaction resp. a stable
//     counting sort, bit-identical to the host front end (csrc/host_frontend.cc);
//   * the per-row rating counts the planner needs (EmfLord.getStats / doUpdateStats, EmfLord.js:48-128, 255-397);
//   * the concatenated portion headers after the conversion loop's quirk Q2 (EmfMaster.js:582-609): inside
//     every portion the LAST rating is dropped — the last non-empty row is emitted one short, and not at all
//     when that rating was its only one (unless it is the portion's only row, which is emitted with 0 columns).
#pragma once
#include "common.cuh"

namespace ycnr {

constexpr int kScanThreads = 256;
constexpr int kScanRowsPerBlock = 2048;
constexpr int kItemChunk = 32768;        // table entries per block of the by-item counting sort (grows with the catalog, see ycnr_rowset_from_table)

__device__ __forceinline__ bool in_set(int8_t dt, uint32_t mask) { return (mask >> (uint32_t)(dt & 31)) & 1u; }

// ---- generic exclusive scan of int32 values into int64 (three launches; out has n + 1 entries) -----------
__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const int32_t* __restrict__ v, int n,
                                                                       int64_t* __restrict__ block_sums) {
  __shared__ int64_t ws[kScanThreads / 32];
  const int base = blockIdx.x * kScanRowsPerBlock;
  int64_t s = 0;
  for (int r = base + threadIdx.x; r < min(n, base + kScanRowsPerBlock); r += kScanThreads) s += v[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += ws[w];
    block_sums[blockIdx.x] = t;
  }
}

// (the block sums themselves are scanned by header_scan_blocks_kernel, portion_kernels.cuh)
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* __restrict__ v, int n,
                                                                  const int64_t* __restrict__ block_offs,
                                                                  int64_t* __restrict__ out) {
  __shared__ int64_t ws[kScanThreads / 32];
  __shared__ int64_t carry;
  const int base = blockIdx.x * kScanRowsPerBlock;
  if (threadIdx.x == 0) carry = block_offs[blockIdx.x];
  __syncthreads();
  for (int p0 = 0; p0 < kScanRowsPerBlock; p0 += kScanThreads) {
    const int r = base + p0 + threadIdx.x;
    const int32_t x = r < n ? v[r] : 0;
    int64_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    int64_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
      const int64_t t = ws[w];
      if (w < (threadIdx.x >> 5)) woff += t;
      total += t;
    }
    const int64_t c = carry;
    if (r < n) {
      out[r] = c + woff + inc - x;
      if (r == n - 1) out[n] = c + woff + inc;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
}

// ---- table helpers ------------------------------------------------------------------------------------
// elem_user[e] = u for every entry of user u (the table's user_list_id column), one warp per user
__global__ void __launch_bounds__(256) fill_elem_user_kernel(const int64_t* __restrict__ user_ptr, int users,
                                                             int32_t* __restrict__ elem_user) {
  const int u = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (u >= users) return;
  const int lane = threadIdx.x & 31;
  for (int64_t e = user_ptr[u] + lane; e < user_ptr[u + 1]; e += 32) elem_user[e] = u;
}

// cnt[u] = number of the user's ratings whose dataset_type is in the set; one warp per user
__global__ void __launch_bounds__(256) count_by_user_kernel(const int64_t* __restrict__ user_ptr,
                                                            const int8_t* __restrict__ dt, uint32_t mask, int users,
                                                            int32_t* __restrict__ cnt) {
  const int u = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (u >= users) return;
  const int lane = threadIdx.x & 31;
  int c = 0;
  for (int64_t e = user_ptr[u] + lane; e < user_ptr[u + 1]; e += 32) c += in_set(dt[e], mask) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[u] = c;
}

// the fetch "by user": the user's ratings of the set, in table (= item) order, to idx/vals[ptr[u] ..)
__global__ void __launch_bounds__(256) fill_by_user_kernel(const int64_t* __restrict__ user_ptr,
                                                           const int32_t* __restrict__ item, const float* __restrict__ rating,
                                                           const int8_t* __restrict__ dt, uint32_t mask, int users,
                                                           const int64_t* __restrict__ ptr, int32_t* __restrict__ idx,
                                                           float* __restrict__ vals) {
  const int u = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (u >= users) return;
  const int lane = threadIdx.x & 31;
  int64_t base = ptr[u];
  const int64_t e1 = user_ptr[u + 1];
  for (int64_t e0 = user_ptr[u]; e0 < e1; e0 += 32) {
    const int64_t e = e0 + lane;
    const bool ok = e < e1 && in_set(dt[e], mask);
    const uint32_t b = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int64_t pos = base + __popc(b & ((1u << lane) - 1u));
      idx[pos] = item[e];
      vals[pos] = rating[e];
    }
    base += __popc(b);
  }
}

// ---- the fetch "by item": stable counting sort of the table by item id ----------------------------------
// counts[b][i] = entries of item i (in the set) inside table chunk b
__global__ void __launch_bounds__(256) item_hist_kernel(const int32_t* __restrict__ item, const int8_t* __restrict__ dt,
                                                        uint32_t mask, int64_t nnz, int items, int64_t chunk,
                                                        int32_t* __restrict__ counts) {
  const int64_t e0 = (int64_t)blockIdx.x * chunk;
  const int64_t e1 = min(nnz, e0 + chunk);
  int32_t* mine = counts + (size_t)blockIdx.x * items;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x)
    if (in_set(dt[e], mask)) atomicAdd(mine + item[e], 1);
}

// per item: counts[b][i] -> offset of chunk b inside the item's row; totals[i] = the row's length
__global__ void __launch_bounds__(256) item_chunk_offsets_kernel(int32_t* __restrict__ counts, int n_chunks, int items,
                                                                 int32_t* __restrict__ totals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= items) return;
  int32_t run = 0;
  for (int b = 0; b < n_chunks; ++b) {
    const int32_t c = counts[(size_t)b * items + i];
    counts[(size_t)b * items + i] = run;
    run += c;
  }
  totals[i] = run;
}

// one warp per table chunk walks its entries IN ORDER, 32 at a time: entries of the same item inside a group
// are ranked by lane (match_any), the lowest lane advances the chunk's cursor of that item — so every item's
// row keeps the table's user order, exactly as the host front end produces it.
__global__ void __launch_bounds__(32) item_scatter_kernel(const int32_t* __restrict__ item, const float* __restrict__ rating,
                                                          const int8_t* __restrict__ dt, const int32_t* __restrict__ elem_user,
                                                          uint32_t mask, int64_t nnz, int items, int64_t chunk,
                                                          int32_t* __restrict__ cursors, const int64_t* __restrict__ item_ptr,
                                                          int32_t* __restrict__ idx, float* __restrict__ vals) {
  const int lane = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * chunk;
  const int64_t c1 = min(nnz, c0 + chunk);
  int32_t* cur = cursors + (size_t)blockIdx.x * items;
  for (int64_t e0 = c0; e0 < c1; e0 += 32) {
    const int64_t e = e0 + lane;
    const bool ok = e < c1 && in_set(dt[e], mask);
    const int it = ok ? item[e] : -1 - lane;                  // unique keys for the entries that do not take part
    const uint32_t peers = __match_any_sync(0xffffffffu, it);
    const int leader = __ffs(peers) - 1;
    int32_t base = 0;
    if (ok && lane == leader) {
      base = cur[it];
      cur[it] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) {
      const int64_t pos = item_ptr[it] + base + __popc(peers & ((1u << lane) - 1u));
      idx[pos] = elem_user[e];
      vals[pos] = rating[e];
    }
    __syncwarp();                                             // cursor stores visible to the next group's readers
  }
}

// cnt[i] over the set for the planner (ratings_count per item): plain histogram
__global__ void __launch_bounds__(256) count_by_item_kernel(const int32_t* __restrict__ item, const int8_t* __restrict__ dt,
                                                            uint32_t mask, int64_t nnz, int32_t* __restrict__ cnt) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    if (in_set(dt[e], mask)) atomicAdd(cnt + item[e], 1);
}

// ---- portion headers with quirk Q2 (EmfMaster.js:582-609) -----------------------------------------------
// Portion p covers rows [from_p, to_p) = [pto[p-1], pto[p]), the first one starts at first_row (a rank's slice
// of the plan does not start at row 0).  last_row[p] = its last non-empty row (or -1),
// drop_last[p] = 1 when that row is not emitted (its only rating is the dropped one and it is not alone).
__global__ void __launch_bounds__(256) portion_tail_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ pto,
                                                           int n_portions, int first_row, int32_t* __restrict__ last_row,
                                                           int32_t* __restrict__ drop_last) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_portions) return;
  const int from = p == 0 ? first_row : pto[p - 1], to = pto[p];
  int r = to - 1;
  while (r >= from && ptr[r + 1] == ptr[r]) --r;
  if (r < from) {
    last_row[p] = -1;
    drop_last[p] = 0;
    return;
  }
  last_row[p] = r;
  int dropped = 0;
  if (ptr[r + 1] - ptr[r] == 1) {          // its only rating is the portion's last one
    int q = r - 1;
    while (q >= from && ptr[q + 1] == ptr[q]) --q;
    dropped = q >= from ? 1 : 0;           // another row exists: this one is never emitted; alone: emitted with 0 columns
  }
  drop_last[p] = dropped;
}

// per row: emitted flag and emitted length
__global__ void __launch_bounds__(256) row_emit_kernel(const int64_t* __restrict__ ptr, int rows, const int32_t* __restrict__ pto,
                                                       int n_portions, int first_row, const int32_t* __restrict__ last_row,
                                                       const int32_t* __restrict__ drop_last, int32_t* __restrict__ flag,
                                                       int32_t* __restrict__ len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t c = ptr[r + 1] - ptr[r];
  int f = 0, l = 0;
  if (c > 0 && n_portions > 0 && r >= first_row && r < pto[n_portions - 1]) {
    int lo = 0, hi = n_portions - 1;       // first portion whose upper bound exceeds r
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (pto[mid] > r) hi = mid; else lo = mid + 1;
    }
    if (last_row[lo] == r) {
      f = drop_last[lo] ? 0 : 1;
      l = (int)c - 1;
    } else {
      f = 1;
      l = (int)c;
    }
  }
  flag[r] = f;
  len[r] = l;
}

__global__ void __launch_bounds__(256) row_scatter_kernel(const int64_t* __restrict__ ptr, int rows, const int32_t* __restrict__ flag,
                                                          const int32_t* __restrict__ len, const int64_t* __restrict__ pos,
                                                          int32_t* __restrict__ row_ids, int64_t* __restrict__ row_start,
                                                          int32_t* __restrict__ row_len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows || !flag[r]) return;
  const int64_t o = pos[r];
  row_ids[o] = r;
  row_start[o] = ptr[r];
  row_len[o] = len[r];
}

// portion_first[p] = number of emitted rows before portion p; portion_first[P] = all of them
__global__ void __launch_bounds__(256) portion_first_kernel(const int64_t* __restrict__ pos, int rows, const int32_t* __restrict__ pto,
                                                            int n_portions, int first_row, int32_t* __restrict__ portion_first) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n_portions) return;
  const int from = p == 0 ? first_row : pto[p - 1];   // rows before first_row are never emitted: pos[first_row] = 0
  portion_first[p] = (int32_t)pos[min(from, rows)];
}

// ---- first-time split into train / validate / test ----------------------------------
// EmfLord.doSplitToSets, JS path (lib/emf/EmfLord.js:450-473), per user with n ratings:
//   t0 = ceil(n*p0/100), t1 = ceil(n*(p0+p1)/100) - t0, t2 = n - t0 - t1; knuth-shuffle of the user's entries;
//   the first t0 -> train (1), the next t1 -> validate (2), the rest -> test (3).
// Upstream shuffles with Math.random(); here, as in the host front end (csrc/host_frontend.cc:208-244),
// random() = u01(mix64(seed, user, step)) — integer hashing and IEEE double multiply/floor only, so the
// device reproduces the host's dataset_type bytes exactly.  One thread per user, permutation in scratch.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t mix64_dev(uint64_t seed, uint64_t a, uint64_t b) {
  return splitmix64(splitmix64(splitmix64(seed) + a) + b);
}

__global__ void __launch_bounds__(128) split_sets_kernel(uint64_t seed, int users, const int64_t* __restrict__ user_ptr,
                                                         int p0, int p1, int32_t* __restrict__ perm,
                                                         int8_t* __restrict__ dataset_type) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= users) return;
  const int64_t beg = user_ptr[u];
  const int64_t n = user_ptr[u + 1] - beg;
  if (n <= 0) return;
  const int64_t t0 = (int64_t)ceil(__dmul_rn((double)n, (double)p0) / 100.0);
  const int64_t t1 = (int64_t)ceil(__dmul_rn((double)n, (double)(p0 + p1)) / 100.0) - t0;
  const int64_t t2 = n - (t0 + t1);
  int64_t c0 = max((int64_t)0, t0), c1 = max((int64_t)0, t1), c2 = max((int64_t)0, t2);
  if (c0 + c1 + c2 < n) c0 += n - (c0 + c1 + c2);
  int32_t* pos = perm + beg;
  for (int64_t i = 0; i < n; ++i) pos[i] = (int32_t)i;
  int64_t cur = n;
  uint64_t step = 0;
  while (cur != 0) {   // knuth-shuffle: r = floor(random() * cur); cur--; swap(a[cur], a[r])
    const double rnd = __dmul_rn((double)(mix64_dev(seed, (uint64_t)u, step++) >> 11), 1.0 / 9007199254740992.0);
    const int64_t r = (int64_t)floor(__dmul_rn(rnd, (double)cur));
    cur -= 1;
    const int32_t a = pos[cur], b = pos[r];
    pos[cur] = b;
    pos[r] = a;
  }
  int64_t offs = 0;
  const int64_t cnts[3] = {c0, c1, c2};
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    for (int64_t i = offs; i < min(n, offs + cnts[s]); ++i) dataset_type[beg + pos[i]] = (int8_t)(s + 1);
    if (cnts[s]) offs += cnts[s];
  }
}

}  // namespace ycnr
