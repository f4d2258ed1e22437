// host_frontend.cc — CPU front end feeding the ALS hot path (see include/ycnr_host.h).
//
// Builds libycnr_host.so with plain g++ (no CUDA).  Each function cites the
// upstream code whose behaviour it reproduces; the bit-exact checker is
// oracle/front_end.py (tests/test_front_end.py).
#include "ycnr_host.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t mix64(uint64_t seed, uint64_t a, uint64_t b) {
  return splitmix(splitmix(splitmix(seed) + a) + b);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

// Box-Muller on two uniforms derived from one hash.
inline double normal01(uint64_t h) {
  double u1 = u01(h);
  double u2 = u01(splitmix(h));
  if (u1 < 1e-300) u1 = 1e-300;
  return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
}

void parallel_for(int64_t n, int nthreads, const std::function<void(int64_t, int64_t, int)>& fn) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (n < 4096 || nthreads == 1) {
    fn(0, n, 0);
    return;
  }
  // fine-grained static chunks handed out round-robin keep skewed rows balanced
  std::vector<std::thread> th;
  int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    int64_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back(fn, lo, hi, t);
  }
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

const char* ycnr_host_last_error(void) { return g_err; }
uint64_t ycnr_mix64(uint64_t seed, uint64_t a, uint64_t b) { return mix64(seed, a, b); }
double ycnr_u01(uint64_t h) { return u01(h); }

// ---------------------------------------------------------------------------------
// Synthetic ratings table.  Stands in for malrec_ratings (data/db-schema.sql:887-893);
// shapes per SURVEY.md §8(d).
// ---------------------------------------------------------------------------------
int ycnr_synth_user_counts(uint64_t seed, int32_t users, int32_t items, int64_t target_nnz,
                           double alpha, int32_t* counts) {
  if (users <= 0 || items <= 0 || !counts) return fail("synth_user_counts: bad shape");
  if (target_nnz < users || target_nnz > (int64_t)users * items)
    return fail("synth_user_counts: target_nnz %lld outside [users, users*items]", (long long)target_nnz);
  std::vector<double> raw(users);
  for (int32_t u = 0; u < users; ++u) {
    double x = u01(mix64(seed, 1, (uint64_t)u));
    raw[u] = std::pow(1.0 - x, -1.0 / alpha);  // Pareto(alpha), >= 1
  }
  auto total = [&](double xm) {
    int64_t s = 0;
    for (int32_t u = 0; u < users; ++u) {
      double c = std::floor(xm * raw[u]);
      if (c < 1) c = 1;
      if (c > items) c = items;
      s += (int64_t)c;
    }
    return s;
  };
  double lo = 1e-6, hi = (double)items;
  for (int it = 0; it < 80; ++it) {
    double mid = 0.5 * (lo + hi);
    if (total(mid) < target_nnz) lo = mid; else hi = mid;
  }
  int64_t s = 0;
  for (int32_t u = 0; u < users; ++u) {
    double c = std::floor(lo * raw[u]);
    if (c < 1) c = 1;
    if (c > items) c = items;
    counts[u] = (int32_t)c;
    s += counts[u];
  }
  // top up / trim one rating at a time in user order until the total is exact
  int64_t diff = target_nnz - s;
  for (int pass = 0; diff != 0 && pass < items + 2; ++pass) {
    for (int32_t u = 0; u < users && diff != 0; ++u) {
      if (diff > 0 && counts[u] < items) { counts[u]++; diff--; }
      else if (diff < 0 && counts[u] > 1) { counts[u]--; diff++; }
    }
  }
  if (diff != 0) return fail("synth_user_counts: could not hit target");
  return 0;
}

int ycnr_synth_fill(uint64_t seed, int32_t users, int32_t items, const int64_t* user_ptr,
                    int32_t max_rating, int32_t rank, double item_skew,
                    int32_t* item_ids, float* ratings, int32_t nthreads) {
  if (users <= 0 || items <= 0 || rank <= 0 || rank > 64) return fail("synth_fill: bad shape");
  // popularity rank -> item id scramble (multiplicative permutation)
  int64_t P = 7919;
  auto gcd = [](int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; };
  while (gcd(P, items) != 1) P += 2;
  const int64_t Qoff = (int64_t)(mix64(seed, 6, 0) % (uint64_t)items);
  std::vector<float> qtab((size_t)items * rank);
  for (int64_t e = 0; e < (int64_t)items * rank; ++e) qtab[e] = (float)normal01(mix64(seed, 4, (uint64_t)e));
  const double mid = 0.65 * max_rating, amp = 0.20 * max_rating, noise = 0.12 * max_rating;
  const double inv_sqrt_rank = 1.0 / std::sqrt((double)rank);

  parallel_for(users, nthreads, [&](int64_t lo, int64_t hi, int) {
    std::vector<uint64_t> bits(((size_t)items + 63) / 64);
    std::vector<float> pu(rank);
    for (int64_t u = lo; u < hi; ++u) {
      const int64_t beg = user_ptr[u];
      const int32_t n = (int32_t)(user_ptr[u + 1] - beg);
      if (n <= 0) continue;
      std::fill(bits.begin(), bits.end(), 0ull);
      const bool invert = n > items / 2;       // dense user: pick the complement instead
      const int32_t want = invert ? items - n : n;
      int32_t got = 0;
      uint64_t j = 0;
      while (got < want) {
        double x = u01(mix64(seed ^ 0x5bd1e995u, (uint64_t)u, j++));
        double y = std::pow(x, item_skew);
        if (invert) y = 1.0 - y;
        int64_t r = (int64_t)(y * items);
        if (r >= items) r = items - 1;
        if (r < 0) r = 0;
        int64_t it = (r * P + Qoff) % items;
        if (j > (uint64_t)want * 64 + 1024) {  // pathological: linear probe to the next free id
          while (bits[it >> 6] >> (it & 63) & 1) it = (it + 1) % items;
        }
        uint64_t m = 1ull << (it & 63);
        if (bits[it >> 6] & m) continue;
        bits[it >> 6] |= m;
        ++got;
      }
      for (int t = 0; t < rank; ++t) pu[t] = (float)normal01(mix64(seed, 3, (uint64_t)u * rank + t));
      int64_t w = beg;
      const size_t nwords = bits.size();
      for (size_t wd = 0; wd < nwords; ++wd) {
        uint64_t word = invert ? ~bits[wd] : bits[wd];
        if (wd == nwords - 1 && (items & 63)) word &= (1ull << (items & 63)) - 1;
        while (word) {
          const int32_t it = (int32_t)(wd * 64 + __builtin_ctzll(word));
          word &= word - 1;
          double dot = 0;
          const float* q = &qtab[(size_t)it * rank];
          for (int t = 0; t < rank; ++t) dot += (double)pu[t] * q[t];
          double v = mid + amp * dot * inv_sqrt_rank + noise * normal01(mix64(seed, 5, (uint64_t)w));
          double rr = std::floor(v + 0.5);
          if (rr < 1) rr = 1;
          if (rr > max_rating) rr = max_rating;
          item_ids[w] = it;
          ratings[w] = (float)rr;
          ++w;
        }
      }
    }
  });
  return 0;
}

int ycnr_init_factors(uint64_t seed, int32_t which, int64_t count, double mean, double dev,
                      float* out, int32_t nthreads) {
  if (count < 0 || !out) return fail("init_factors: bad args");
  parallel_for(count, nthreads, [&](int64_t lo, int64_t hi, int) {
    for (int64_t e = lo; e < hi; ++e)
      out[e] = (float)(mean + dev * normal01(mix64(seed, 10 + (uint64_t)which, (uint64_t)e)));
  });
  return 0;
}

// ---------------------------------------------------------------------------------
// Split rule Q9 — EmfLord.js:450-473 on a first split (all ratings dataset_type 0):
//   target[0] = ceil(n*p0/100); target[1] = ceil(n*(p0+p1)/100) - target[0];
//   target[2] = n - target[0] - target[1]; shuffle(freeIds) (knuth-shuffle);
//   first target[0] ids -> train(1), next target[1] -> validate(2), rest -> test(3).
// Math.random() is replaced by u01(mix64(seed, user, step)).
// ---------------------------------------------------------------------------------
int ycnr_split_sets(uint64_t seed, int32_t users, const int64_t* user_ptr,
                    const int32_t pcts[3], int8_t* dataset_type, int32_t nthreads) {
  if (users < 0 || !user_ptr || !dataset_type) return fail("split_sets: bad args");
  const int32_t p0 = pcts[0], p1 = pcts[1];
  parallel_for(users, nthreads, [&](int64_t lo, int64_t hi, int) {
    std::vector<int32_t> pos;
    for (int64_t u = lo; u < hi; ++u) {
      const int64_t beg = user_ptr[u];
      const int64_t n = user_ptr[u + 1] - beg;
      if (n <= 0) continue;
      // JS: Math.ceil(totalCnt * pcts[0] / 100) in doubles
      int64_t t0 = (int64_t)std::ceil((double)n * (double)p0 / 100.0);
      int64_t t1 = (int64_t)std::ceil((double)n * (double)(p0 + p1) / 100.0) - t0;
      int64_t t2 = n - (t0 + t1);
      int64_t c0 = std::max<int64_t>(0, t0), c1 = std::max<int64_t>(0, t1), c2 = std::max<int64_t>(0, t2);
      if (c0 + c1 + c2 < n) c0 += n - (c0 + c1 + c2);
      pos.resize(n);
      for (int64_t i = 0; i < n; ++i) pos[i] = (int32_t)i;
      // knuth-shuffle: while (cur) { r = floor(random()*cur); cur--; swap(a[cur], a[r]); }
      int64_t cur = n;
      uint64_t step = 0;
      while (cur != 0) {
        int64_t r = (int64_t)std::floor(u01(mix64(seed, (uint64_t)u, step++)) * (double)cur);
        cur -= 1;
        std::swap(pos[cur], pos[r]);
      }
      int64_t offs = 0;
      const int64_t cnts[3] = {c0, c1, c2};
      for (int s = 0; s < 3; ++s) {
        // slice(offs, offs + cnt) clamps at the array end
        for (int64_t i = offs; i < std::min(n, offs + cnts[s]); ++i) dataset_type[beg + pos[i]] = (int8_t)(s + 1);
        if (cnts[s]) offs += cnts[s];
      }
    }
  });
  return 0;
}

// ---------------------------------------------------------------------------------
// Portion planner Q6 — EmfLord.js:510-612, one stepType per call.
// ---------------------------------------------------------------------------------
int ycnr_split_to_portions(const int32_t* cnt_per_row, int32_t total_rows,
                           int32_t ratings_in_portion_opt, int32_t num_threads_opt,
                           int32_t pct_plus1,
                           int32_t* portions_row_id_to, int32_t cap, int32_t* n_portions_out,
                           int32_t* max_ratings_in_portion_out, int32_t* max_rows_in_portion_out) {
  if (!cnt_per_row || total_rows <= 0) return fail("split_to_portions: bad args");
  // stats as getStats() builds them (EmfLord.js:95-119): rows with cnt == 0 are absent
  double rowsCnt = 0, ratingsCount = 0, maxRatingsPerRow = 0;
  for (int32_t r = 0; r < total_rows; ++r) {
    if (cnt_per_row[r] == 0) continue;
    rowsCnt += 1;
    ratingsCount += cnt_per_row[r];
    if (cnt_per_row[r] > maxRatingsPerRow) maxRatingsPerRow = cnt_per_row[r];
  }
  if (ratingsCount == 0) { *n_portions_out = 0; return 0; }
  const double frac = pct_plus1 ? ((double)pct_plus1 / 100.0) : 1.0;
  if (pct_plus1) {
    ratingsCount = std::ceil(ratingsCount * frac);
    maxRatingsPerRow = std::ceil(maxRatingsPerRow * frac);
  }
  double ratingsInPortion = ratings_in_portion_opt;
  double avgPortionsCount = std::ceil(ratingsCount / ratingsInPortion);
  double avgRowsInPortion = std::floor(rowsCnt / avgPortionsCount);
  if (avgPortionsCount < num_threads_opt) {
    avgPortionsCount = num_threads_opt;
    ratingsInPortion = std::ceil(ratingsCount / avgPortionsCount);
    avgRowsInPortion = std::floor(rowsCnt / avgPortionsCount);
  }
  if (avgRowsInPortion < 1) {
    avgRowsInPortion = 1;
    avgPortionsCount = rowsCnt;
    ratingsInPortion = std::ceil(ratingsCount / avgPortionsCount);
  }
  if (ratingsInPortion < maxRatingsPerRow) {
    ratingsInPortion = maxRatingsPerRow;
    avgPortionsCount = std::ceil(ratingsCount / ratingsInPortion);
    avgRowsInPortion = std::floor(rowsCnt / avgPortionsCount);
  }
  int32_t p = 0, rows = 0, maxRows = 0, written = 0;
  double rtgs = 0;
  for (int32_t id = 0; id < total_rows; ++id) {
    if (cnt_per_row[id] == 0) continue;  // "for (let id in ratingsCntPer)" skips holes
    double cnt = cnt_per_row[id];
    if (pct_plus1) cnt = std::ceil(cnt * frac);
    if ((rtgs + cnt) > ratingsInPortion) {
      rtgs = 0;
      rows = 0;
      p++;
    }
    rtgs += cnt;
    rows++;
    if (rows > maxRows) maxRows = rows;
    if (p >= cap) return fail("split_to_portions: capacity %d too small", cap);
    portions_row_id_to[p] = id + 1;
    written = p + 1;
  }
  *n_portions_out = written;
  if (max_ratings_in_portion_out) *max_ratings_in_portion_out = (int32_t)ratingsInPortion;
  if (max_rows_in_portion_out) *max_rows_in_portion_out = maxRows;
  return 0;
}

// ---------------------------------------------------------------------------------
// Fetch filters — EmfMaster.js:501-529.  byUser: ORDER BY user, item.  byItem:
// ORDER BY item only (upstream leaves the column order to the DB, Q4); here user ascending.
// ---------------------------------------------------------------------------------
int ycnr_count_by_user(int32_t users, const int64_t* user_ptr, const int8_t* dt, uint32_t mask,
                       int64_t* out_ptr) {
  out_ptr[0] = 0;
  for (int32_t u = 0; u < users; ++u) {
    int64_t c = 0;
    for (int64_t e = user_ptr[u]; e < user_ptr[u + 1]; ++e) c += (mask >> dt[e]) & 1u;
    out_ptr[u + 1] = out_ptr[u] + c;
  }
  return 0;
}

int ycnr_fill_by_user(int32_t users, const int64_t* user_ptr, const int32_t* item_ids,
                      const float* ratings, const int8_t* dt, uint32_t mask,
                      const int64_t* out_ptr, int32_t* out_idx, float* out_vals) {
  parallel_for(users, 0, [&](int64_t lo, int64_t hi, int) {
    for (int64_t u = lo; u < hi; ++u) {
      int64_t w = out_ptr[u];
      for (int64_t e = user_ptr[u]; e < user_ptr[u + 1]; ++e)
        if ((mask >> dt[e]) & 1u) { out_idx[w] = item_ids[e]; out_vals[w] = ratings[e]; ++w; }
    }
  });
  return 0;
}

int ycnr_count_by_item(int32_t users, int32_t items, const int64_t* user_ptr,
                       const int32_t* item_ids, const int8_t* dt, uint32_t mask, int64_t* out_ptr) {
  std::vector<int64_t> c((size_t)items + 1, 0);
  const int64_t nnz = user_ptr[users];
  for (int64_t e = 0; e < nnz; ++e)
    if ((mask >> dt[e]) & 1u) {
      if (item_ids[e] < 0 || item_ids[e] >= items) return fail("count_by_item: item id out of range");
      c[item_ids[e] + 1]++;
    }
  out_ptr[0] = 0;
  for (int32_t i = 0; i < items; ++i) out_ptr[i + 1] = out_ptr[i] + c[i + 1];
  return 0;
}

int ycnr_fill_by_item(int32_t users, int32_t items, const int64_t* user_ptr,
                      const int32_t* item_ids, const float* ratings, const int8_t* dt,
                      uint32_t mask, const int64_t* out_ptr, int32_t* out_idx, float* out_vals) {
  std::vector<int64_t> w(out_ptr, out_ptr + items);
  for (int32_t u = 0; u < users; ++u)
    for (int64_t e = user_ptr[u]; e < user_ptr[u + 1]; ++e)
      if ((mask >> dt[e]) & 1u) {
        int64_t at = w[item_ids[e]]++;
        out_idx[at] = u;
        out_vals[at] = ratings[e];
      }
  return 0;
}

// ---------------------------------------------------------------------------------
// Portion conversion with the Q2 drop — the loop of EmfMaster.js:582-609, kept in
// its upstream shape on purpose: the row-closing test fires on the last fetched
// rating *before* cols++, so that rating is never covered by a row.
//   emit(rowId, start, cols) is called once per closed row.
// ---------------------------------------------------------------------------------
}  // extern "C"

template <class Emit>
static inline int64_t convert_portion(const int64_t* ptr, int32_t row_from, int32_t row_to, Emit emit) {
  const int64_t base = ptr[row_from];
  const int64_t len = ptr[row_to] - base;  // data.length
  int32_t row = row_from;                  // row of rating i, advanced lazily
  int32_t last_r = 0;
  int64_t cols = 0, start = base;
  for (int64_t i = 0; i < len; ++i) {
    while (ptr[row + 1] <= base + i) ++row;  // data[i].r
    if (i == 0) last_r = row;
    if (last_r != row || i == len - 1) {
      emit(last_r, start, cols);
      start += cols;
      last_r = row;
      cols = 0;
    }
    cols++;
  }
  return len;
}

extern "C" {

int ycnr_build_portion(const int64_t* ptr, const int32_t* idx, const float* vals,
                       int32_t row_from, int32_t row_to,
                       int32_t* buf_rows, int32_t cap_rows_words,
                       int32_t* buf_indx, float* buf_vals, int32_t cap_ratings,
                       int32_t* fetched_out) {
  if (row_to < row_from) return fail("build_portion: bad row range");
  const int64_t base = ptr[row_from];
  const int64_t len = ptr[row_to] - base;
  if (len > cap_ratings) return fail("build_portion: %lld ratings exceed buffer %d (upstream assert, EmfMaster.js:588)",
                                     (long long)len, cap_ratings);
  memcpy(buf_indx, idx + base, (size_t)len * sizeof(int32_t));
  memcpy(buf_vals, vals + base, (size_t)len * sizeof(float));
  int32_t r = 0;
  bool overflow = false;
  convert_portion(ptr, row_from, row_to, [&](int32_t rowId, int64_t, int64_t cols) {
    if (1 + r * 2 + 1 >= cap_rows_words) { overflow = true; return; }
    buf_rows[1 + r * 2] = rowId;
    buf_rows[1 + r * 2 + 1] = (int32_t)cols;
    r++;
  });
  if (overflow) return fail("build_portion: rows buffer too small (upstream assert, EmfMaster.js:595)");
  buf_rows[0] = r;
  if (fetched_out) *fetched_out = (int32_t)len;
  return 0;
}

int ycnr_build_rowlist(const int64_t* ptr, const int32_t* portions_row_id_to, int32_t n_portions,
                       int32_t* row_ids, int64_t* row_start, int32_t* row_len, int32_t cap_rows,
                       int32_t* portion_first) {
  int32_t r = 0;
  bool overflow = false;
  for (int32_t p = 0; p < n_portions; ++p) {
    const int32_t from = p == 0 ? 0 : portions_row_id_to[p - 1];
    const int32_t to = portions_row_id_to[p];
    portion_first[p] = r;
    convert_portion(ptr, from, to, [&](int32_t rowId, int64_t start, int64_t cols) {
      if (r >= cap_rows) { overflow = true; return; }
      row_ids[r] = rowId;
      row_start[r] = start;
      row_len[r] = (int32_t)cols;
      r++;
    });
    if (overflow) return fail("build_rowlist: capacity %d too small", cap_rows);
  }
  portion_first[n_portions] = r;
  return 0;
}

}  // extern "C"
