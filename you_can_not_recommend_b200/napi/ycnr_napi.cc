// ycnr_napi.cc — thin N-API shim over the C ABI of libycnr_als.so (include/ycnr_als.h).
//
// Takes the slot of the reference's raw-V8 addon (cpp_utils/cpp_utils.cc:3-8, built by
// binding.gyp:4-9 as build/Release/cpp_utils): the module keeps the two upstream exports
//   sAlsBuildSubFixedFacts(sub, fixed, indx, cols, k)   (cpp_utils/als_utils.cc:22-38)
//   dAlsBuildSubFixedFacts(...)                         (float64: rejected, GPU path is float32)
// and adds the calls EmfWorker's portion handlers make instead of their BLAS/LAPACK loops
// (lib/emf/EmfWorker.js:169-261, 266-315).  No arithmetic here: argument marshalling and
// status -> thrown Error only, so the worker's uncaughtException path fires on failure
// (lib/emf/EmfWorkerProcess.js:30-45).  Cannot be executed in this image (no node); it is
// syntax-checked against node_api_min.h by tests/test_napi_syntax.py.
#ifdef YCNR_NAPI_MIN
#include "node_api_min.h"
#else
#include <node_api.h>
#endif
#include <stdint.h>
#include <string.h>

#include <vector>

#include "ycnr_als.h"

namespace {

bool fail(napi_env env, const char* what) {
  napi_throw_error(env, "YCNR", what);
  return false;
}
bool check(napi_env env, int rc) { return rc == 0 ? true : fail(env, ycnr_last_error()); }

bool args(napi_env env, napi_callback_info info, size_t want, napi_value* argv) {
  size_t argc = want;
  if (napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr) != napi_ok || argc < want)
    return fail(env, "wrong number of arguments");
  return true;
}

template <class T>
bool typed(napi_env env, napi_value v, napi_typedarray_type want, T** data, size_t* len) {
  napi_typedarray_type t;
  void* p = nullptr;
  if (napi_get_typedarray_info(env, v, &t, len, &p, nullptr, nullptr) != napi_ok || t != want)
    return fail(env, "invalid type!");  // message of cpp_utils.js:13
  *data = static_cast<T*>(p);
  return true;
}

bool handle(napi_env env, napi_value v, ycnr_ctx** out) {
  void* p = nullptr;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p) return fail(env, "invalid context handle");
  *out = static_cast<ycnr_ctx*>(p);
  return true;
}

napi_value undefined(napi_env env) {
  napi_value u;
  napi_get_undefined(env, &u);
  return u;
}

double get_num(napi_env env, napi_value obj, const char* name, double dflt) {
  bool has = false;
  napi_value v;
  double d = dflt;
  if (napi_has_named_property(env, obj, name, &has) == napi_ok && has &&
      napi_get_named_property(env, obj, name, &v) == napi_ok)
    napi_get_value_double(env, v, &d);
  return d;
}
bool get_flag(napi_env env, napi_value obj, const char* name) {
  bool has = false, b = false;
  napi_value v;
  if (napi_has_named_property(env, obj, name, &has) == napi_ok && has &&
      napi_get_named_property(env, obj, name, &v) == napi_ok)
    napi_get_value_bool(env, v, &b);
  return b;
}

void finalize_ctx(napi_env, void* data, void*) { ycnr_destroy(static_cast<ycnr_ctx*>(data)); }

// create({factorsCount, totalUsersCount, totalItemsCount, userFactReg, itemFactReg,
//         useDoublePrecision, lowmem, device}) -> handle
napi_value Create(napi_env env, napi_callback_info info) {
  napi_value a[1];
  if (!args(env, info, 1, a)) return nullptr;
  ycnr_options o;
  memset(&o, 0, sizeof(o));
  o.factors_count = (int32_t)get_num(env, a[0], "factorsCount", 100);
  o.total_users = (int32_t)get_num(env, a[0], "totalUsersCount", 0);
  o.total_items = (int32_t)get_num(env, a[0], "totalItemsCount", 0);
  o.user_fact_reg = get_num(env, a[0], "userFactReg", 0.05);
  o.item_fact_reg = get_num(env, a[0], "itemFactReg", 0.05);
  o.use_double_precision = get_flag(env, a[0], "useDoublePrecision");
  o.lowmem = get_flag(env, a[0], "lowmem");
  o.device = (int32_t)get_num(env, a[0], "device", 0);
  o.dual_max_cols = -1;
  ycnr_ctx* c = nullptr;
  if (!check(env, ycnr_create(&o, &c))) return nullptr;
  napi_value ext;
  napi_create_external(env, c, finalize_ctx, nullptr, &ext);
  return ext;
}

// attachFactors(handle, Float32Array userFactors, Float32Array itemFactors)   (shm-typed-array views)
napi_value AttachFactors(napi_env env, napi_callback_info info) {
  napi_value a[3];
  ycnr_ctx* c;
  float *u, *v;
  size_t nu, nv;
  if (!args(env, info, 3, a) || !handle(env, a[0], &c) || !typed(env, a[1], napi_float32_array, &u, &nu) ||
      !typed(env, a[2], napi_float32_array, &v, &nv))
    return nullptr;
  int64_t want_u = 0, want_v = 0;
  if (!check(env, ycnr_factor_elems(c, YCNR_USER_FACTORS, &want_u)) || !check(env, ycnr_factor_elems(c, YCNR_ITEM_FACTORS, &want_v)))
    return nullptr;
  if ((int64_t)nu < want_u || (int64_t)nv < want_v) {
    fail(env, "attachFactors: factor arrays are smaller than total x factorsCount");
    return nullptr;
  }
  check(env, ycnr_attach_factors(c, u, v));
  return undefined(env);
}

napi_value StartTrainStep(napi_env env, napi_callback_info info) {
  napi_value a[2];
  ycnr_ctx* c;
  int32_t step;
  if (!args(env, info, 2, a) || !handle(env, a[0], &c) || napi_get_value_int32(env, a[1], &step) != napi_ok) return nullptr;
  check(env, ycnr_start_train_step(c, step));
  return undefined(env);
}

napi_value portion_result(napi_env env, const ycnr_portion_info& pi, bool rmse) {
  napi_value o, v;
  napi_create_object(env, &o);
  napi_create_int32(env, pi.rows_from, &v); napi_set_named_property(env, o, "rowsFrom", v);
  napi_create_int32(env, pi.rows_cnt, &v); napi_set_named_property(env, o, "rowsCnt", v);
  napi_create_int64(env, pi.ratings_in_portion, &v); napi_set_named_property(env, o, "ratingsInPortion", v);
  napi_create_double(env, pi.time_ms, &v); napi_set_named_property(env, o, "time", v);
  if (rmse) {
    napi_create_double(env, pi.r_sum_diff2, &v); napi_set_named_property(env, o, "rSumDiff2", v);
    napi_create_double(env, pi.r_cnt, &v); napi_set_named_property(env, o, "rCnt", v);
    napi_create_double(env, pi.r_sum, &v); napi_set_named_property(env, o, "rSum", v);
  }
  return o;
}

// alsPortion(handle, Int32Array alsRows, Int32Array alsIndx, Float32Array alsVals) -> completedPortion fields
napi_value AlsPortion(napi_env env, napi_callback_info info) {
  napi_value a[4];
  ycnr_ctx* c;
  int32_t *rows, *indx;
  float* vals;
  size_t n0, n1, n2;
  if (!args(env, info, 4, a) || !handle(env, a[0], &c) || !typed(env, a[1], napi_int32_array, &rows, &n0) ||
      !typed(env, a[2], napi_int32_array, &indx, &n1) || !typed(env, a[3], napi_float32_array, &vals, &n2))
    return nullptr;
  ycnr_portion_info pi;
  if (!check(env, ycnr_check_portion(rows, (int64_t)n0, (int64_t)n1, (int64_t)n2))) return nullptr;
  if (!check(env, ycnr_als_portion(c, rows, indx, vals, &pi))) return nullptr;
  return portion_result(env, pi, false);
}

napi_value EndTrainStep(napi_env env, napi_callback_info info) {
  napi_value a[1];
  ycnr_ctx* c;
  if (!args(env, info, 1, a) || !handle(env, a[0], &c)) return nullptr;
  check(env, ycnr_end_train_step(c));
  return undefined(env);
}

napi_value StartCalcRmse(napi_env env, napi_callback_info info) {
  napi_value a[3];
  ycnr_ctx* c;
  int32_t step;
  double shift;
  if (!args(env, info, 3, a) || !handle(env, a[0], &c) || napi_get_value_int32(env, a[1], &step) != napi_ok ||
      napi_get_value_double(env, a[2], &shift) != napi_ok)
    return nullptr;
  check(env, ycnr_start_calc_rmse(c, step, shift));
  return undefined(env);
}

napi_value RmsePortion(napi_env env, napi_callback_info info) {
  napi_value a[4];
  ycnr_ctx* c;
  int32_t *rows, *indx;
  float* vals;
  size_t n0, n1, n2;
  if (!args(env, info, 4, a) || !handle(env, a[0], &c) || !typed(env, a[1], napi_int32_array, &rows, &n0) ||
      !typed(env, a[2], napi_int32_array, &indx, &n1) || !typed(env, a[3], napi_float32_array, &vals, &n2))
    return nullptr;
  ycnr_portion_info pi;
  if (!check(env, ycnr_check_portion(rows, (int64_t)n0, (int64_t)n1, (int64_t)n2))) return nullptr;
  if (!check(env, ycnr_rmse_portion(c, rows, indx, vals, &pi))) return nullptr;
  return portion_result(env, pi, true);
}

// rmsePortionAsync(handle, rmseRows, rmseIndx, rmseVals, portionNo): queue the portion (small portions are launched
// in batches); its sums come back from rmsePoll.  The worker's 'completedPortion' replies are asynchronous
// messages upstream as well (EmfWorker.js:304-314).
napi_value RmsePortionAsync(napi_env env, napi_callback_info info) {
  napi_value a[5];
  ycnr_ctx* c;
  int32_t *rows, *indx;
  float* vals;
  size_t n0, n1, n2;
  int64_t tag;
  if (!args(env, info, 5, a) || !handle(env, a[0], &c) || !typed(env, a[1], napi_int32_array, &rows, &n0) ||
      !typed(env, a[2], napi_int32_array, &indx, &n1) || !typed(env, a[3], napi_float32_array, &vals, &n2) ||
      napi_get_value_int64(env, a[4], &tag) != napi_ok)
    return nullptr;
  if (!check(env, ycnr_check_portion(rows, (int64_t)n0, (int64_t)n1, (int64_t)n2))) return nullptr;
  check(env, ycnr_rmse_portion_async(c, rows, indx, vals, tag));
  return undefined(env);
}

// Portion buffers of a whole half-step / RMSE pass handed over at once (the master's portion cache,
// usePortionsCache: EmfMaster.js:434-494): three JS arrays of typed arrays -> pointer tables for
// ycnr_als_portions / ycnr_rmse_portions_async, which scan and queue the headers with the library's host threads.
bool portion_tables(napi_env env, napi_value rows_a, napi_value indx_a, napi_value vals_a, std::vector<const int32_t*>& rows,
                    std::vector<const int32_t*>& indx, std::vector<const float*>& vals) {
  uint32_t n = 0, n1 = 0, n2 = 0;
  if (napi_get_array_length(env, rows_a, &n) != napi_ok || napi_get_array_length(env, indx_a, &n1) != napi_ok ||
      napi_get_array_length(env, vals_a, &n2) != napi_ok || n1 != n || n2 != n)
    return fail(env, "portion arrays of different lengths");
  rows.resize(n); indx.resize(n); vals.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    napi_value r, x, v;
    int32_t *pr, *pi;
    float* pv;
    size_t l0, l1, l2;
    if (napi_get_element(env, rows_a, i, &r) != napi_ok || napi_get_element(env, indx_a, i, &x) != napi_ok ||
        napi_get_element(env, vals_a, i, &v) != napi_ok || !typed(env, r, napi_int32_array, &pr, &l0) ||
        !typed(env, x, napi_int32_array, &pi, &l1) || !typed(env, v, napi_float32_array, &pv, &l2))
      return false;
    if (!check(env, ycnr_check_portion(pr, (int64_t)l0, (int64_t)l1, (int64_t)l2))) return false;
    rows[i] = pr; indx[i] = pi; vals[i] = pv;
  }
  return true;
}

// alsPortions(handle, [alsRows...], [alsIndx...], [alsVals...]) -> [completedPortion fields per portion]
napi_value AlsPortions(napi_env env, napi_callback_info info) {
  napi_value a[4];
  ycnr_ctx* c;
  std::vector<const int32_t*> rows, indx;
  std::vector<const float*> vals;
  if (!args(env, info, 4, a) || !handle(env, a[0], &c) || !portion_tables(env, a[1], a[2], a[3], rows, indx, vals)) return nullptr;
  std::vector<ycnr_portion_info> infos(rows.size() ? rows.size() : 1);
  if (!check(env, ycnr_als_portions(c, (int32_t)rows.size(), rows.data(), indx.data(), vals.data(), infos.data()))) return nullptr;
  napi_value arr;
  napi_create_array(env, &arr);
  for (size_t i = 0; i < rows.size(); ++i) napi_set_element(env, arr, (uint32_t)i, portion_result(env, infos[i], false));
  return arr;
}

// rmsePortionsAsync(handle, [rmseRows...], [rmseIndx...], [rmseVals...], firstPortionNo): portion i is queued under
// tag firstPortionNo + i; the sums come back from rmsePoll
napi_value RmsePortionsAsync(napi_env env, napi_callback_info info) {
  napi_value a[5];
  ycnr_ctx* c;
  int64_t first = 0;
  std::vector<const int32_t*> rows, indx;
  std::vector<const float*> vals;
  if (!args(env, info, 5, a) || !handle(env, a[0], &c) || !portion_tables(env, a[1], a[2], a[3], rows, indx, vals) ||
      napi_get_value_int64(env, a[4], &first) != napi_ok)
    return nullptr;
  std::vector<int64_t> tags(rows.size());
  for (size_t i = 0; i < rows.size(); ++i) tags[i] = first + (int64_t)i;
  check(env, ycnr_rmse_portions_async(c, (int32_t)rows.size(), rows.data(), indx.data(), vals.data(), tags.data()));
  return undefined(env);
}

// rmsePoll(handle, wait) -> [{portionNo, rowsFrom, rowsCnt, ratingsInPortion, time, rSumDiff2, rCnt, rSum}, ...] in
// the order the portions were queued; wait = true flushes the queue and waits for all of them.
napi_value RmsePoll(napi_env env, napi_callback_info info) {
  napi_value a[2];
  ycnr_ctx* c;
  bool wait = false;
  if (!args(env, info, 2, a) || !handle(env, a[0], &c) || napi_get_value_bool(env, a[1], &wait) != napi_ok) return nullptr;
  napi_value arr;
  napi_create_array(env, &arr);
  uint32_t out = 0;
  for (;;) {
    int64_t tags[256];
    ycnr_portion_info infos[256];
    int32_t n = 0;
    if (!check(env, ycnr_rmse_poll(c, wait ? 1 : 0, 256, tags, infos, &n))) return nullptr;
    for (int32_t i = 0; i < n; ++i) {
      napi_value o = portion_result(env, infos[i], true), v;
      napi_create_int64(env, tags[i], &v);
      napi_set_named_property(env, o, "portionNo", v);
      napi_set_element(env, arr, out++, o);
    }
    if (n < 256) break;
  }
  return arr;
}

// sAlsBuildSubFixedFacts(sub, fixed, indx, cols, k) — upstream's own signature (cpp_utils/cpp_utils.js:15-19,
// als_utils.cc:22-38): runs on the process's current context; a 6-argument call with the context handle first
// is accepted as well.
napi_value SAlsBuildSubFixedFacts(napi_env env, napi_callback_info info) {
  napi_value a[6];
  size_t argc = 6;
  if (napi_get_cb_info(env, info, &argc, a, nullptr, nullptr) != napi_ok || argc < 5) {
    fail(env, "wrong number of arguments");
    return nullptr;
  }
  const size_t o = argc >= 6 ? 1 : 0;
  ycnr_ctx* c = nullptr;
  float *sub, *fixed;
  int32_t* indx;
  size_t ns, nf, ni;
  int32_t cols, k;
  if ((o && !handle(env, a[0], &c)) || !typed(env, a[o], napi_float32_array, &sub, &ns) ||
      !typed(env, a[o + 1], napi_float32_array, &fixed, &nf) || !typed(env, a[o + 2], napi_int32_array, &indx, &ni) ||
      napi_get_value_int32(env, a[o + 3], &cols) != napi_ok || napi_get_value_int32(env, a[o + 4], &k) != napi_ok)
    return nullptr;
  if (k <= 0 || cols < 0 || (size_t)cols > ni || (size_t)cols * (size_t)k > ns) return fail(env, "buffer too small"), nullptr;
  if (o) check(env, ycnr_s_als_build_sub_fixed_facts(c, sub, fixed, (int64_t)(nf / (size_t)k), indx, cols, k));
  else check(env, ycnr_s_als_build_sub_fixed_facts_noctx(sub, fixed, (int64_t)(nf / (size_t)k), indx, cols, k));
  return undefined(env);
}

// getMemoryUsage(handle) -> {device, pinned, deviceFree, deviceTotal}   (EmfWorker.mw_getMemoryUsage, EmfWorker.js:109-113)
napi_value GetMemoryUsage(napi_env env, napi_callback_info info) {
  napi_value a[1];
  ycnr_ctx* c;
  if (!args(env, info, 1, a) || !handle(env, a[0], &c)) return nullptr;
  int64_t mu[4];
  if (!check(env, ycnr_memory_usage(c, mu))) return nullptr;
  napi_value o, v;
  napi_create_object(env, &o);
  napi_create_int64(env, mu[0], &v); napi_set_named_property(env, o, "device", v);
  napi_create_int64(env, mu[1], &v); napi_set_named_property(env, o, "pinned", v);
  napi_create_int64(env, mu[2], &v); napi_set_named_property(env, o, "deviceFree", v);
  napi_create_int64(env, mu[3], &v); napi_set_named_property(env, o, "deviceTotal", v);
  return o;
}

napi_value DAlsBuildSubFixedFacts(napi_env env, napi_callback_info) {
  fail(env, "useDoublePrecision is not supported by the B200 path (float32 only)");
  return nullptr;
}

// ---- bulk mode: a whole step resident on the device ---------------------------
// tableUpload(handle, BigInt64Array userPtr, Int32Array itemIds, Float32Array ratings, Int8Array datasetType)
napi_value TableUpload(napi_env env, napi_callback_info info) {
  napi_value a[5];
  ycnr_ctx* c;
  int64_t* uptr;
  int32_t* items;
  float* ratings;
  int8_t* dt;
  size_t n0, n1, n2, n3;
  if (!args(env, info, 5, a) || !handle(env, a[0], &c) || !typed(env, a[1], napi_bigint64_array, &uptr, &n0) ||
      !typed(env, a[2], napi_int32_array, &items, &n1) || !typed(env, a[3], napi_float32_array, &ratings, &n2) ||
      !typed(env, a[4], napi_int8_array, &dt, &n3))
    return nullptr;
  if (n0 < 1 || n1 != n2 || n1 != n3 || (int64_t)n1 != uptr[n0 - 1]) {
    fail(env, "tableUpload: array sizes do not match userPtr");
    return nullptr;
  }
  check(env, ycnr_table_upload(c, uptr, items, ratings, dt));
  return undefined(env);
}

// tableSplit(handle, seed, p0, p1, p2, Int8Array datasetTypeOut)   (EmfLord.doSplitToSets, EmfLord.js:450-473)
napi_value TableSplit(napi_env env, napi_callback_info info) {
  napi_value a[6];
  ycnr_ctx* c;
  double seed;
  int32_t pcts[3];
  int8_t* out;
  size_t n;
  if (!args(env, info, 6, a) || !handle(env, a[0], &c) || napi_get_value_double(env, a[1], &seed) != napi_ok ||
      napi_get_value_int32(env, a[2], &pcts[0]) != napi_ok || napi_get_value_int32(env, a[3], &pcts[1]) != napi_ok ||
      napi_get_value_int32(env, a[4], &pcts[2]) != napi_ok || !typed(env, a[5], napi_int8_array, &out, &n))
    return nullptr;
  check(env, ycnr_table_split(c, (uint64_t)seed, pcts, out));
  return undefined(env);
}

// tableCounts(handle, setMask, byItem, Int32Array countsOut)   (ratings_count per row for splitToPortions)
napi_value TableCounts(napi_env env, napi_callback_info info) {
  napi_value a[4];
  ycnr_ctx* c;
  uint32_t mask;
  int32_t by_item, *out;
  size_t n;
  if (!args(env, info, 4, a) || !handle(env, a[0], &c) || napi_get_value_uint32(env, a[1], &mask) != napi_ok ||
      napi_get_value_int32(env, a[2], &by_item) != napi_ok || !typed(env, a[3], napi_int32_array, &out, &n))
    return nullptr;
  check(env, ycnr_table_counts(c, mask, by_item, out));
  return undefined(env);
}

// rowsetFromTable(handle, stepType, setMask, firstRow, Int32Array portionsRowIdTo) -> rowset id
napi_value RowsetFromTable(napi_env env, napi_callback_info info) {
  napi_value a[5];
  ycnr_ctx* c;
  int32_t step, first_row, *pto, id = -1;
  uint32_t mask;
  size_t n;
  if (!args(env, info, 5, a) || !handle(env, a[0], &c) || napi_get_value_int32(env, a[1], &step) != napi_ok ||
      napi_get_value_uint32(env, a[2], &mask) != napi_ok || napi_get_value_int32(env, a[3], &first_row) != napi_ok ||
      !typed(env, a[4], napi_int32_array, &pto, &n))
    return nullptr;
  if (!check(env, ycnr_rowset_from_table(c, step, mask, first_row, pto, (int32_t)n, &id))) return nullptr;
  napi_value v;
  napi_create_int32(env, id, &v);
  return v;
}

// alsRowset(handle, rowset): one half-step over a resident row set (EmfLord.alsTrainStep, EmfLord.js:963-984)
napi_value AlsRowset(napi_env env, napi_callback_info info) {
  napi_value a[2];
  ycnr_ctx* c;
  int32_t id;
  if (!args(env, info, 2, a) || !handle(env, a[0], &c) || napi_get_value_int32(env, a[1], &id) != napi_ok) return nullptr;
  check(env, ycnr_als_rowset(c, id));
  return undefined(env);
}

// rmseRowset(handle, rowset, globalAvgShift, Float64Array portionSums /*[3 * portions]*/) -> {rSumDiff2, rCnt, rSum}
// portionSums carries the per-portion partials quirk Q7 needs (EmfMaster.js:777-783 uses the LAST portion's).
napi_value RmseRowset(napi_env env, napi_callback_info info) {
  napi_value a[4];
  ycnr_ctx* c;
  int32_t id, np = 0;
  double shift, *ps, totals[3];
  size_t n;
  if (!args(env, info, 4, a) || !handle(env, a[0], &c) || napi_get_value_int32(env, a[1], &id) != napi_ok ||
      napi_get_value_double(env, a[2], &shift) != napi_ok || !typed(env, a[3], napi_float64_array, &ps, &n))
    return nullptr;
  if (!check(env, ycnr_rowset_info(c, id, nullptr, nullptr, &np))) return nullptr;
  if (n < (size_t)3 * (size_t)np) {
    fail(env, "rmseRowset: portionSums needs 3 doubles per portion");
    return nullptr;
  }
  if (!check(env, ycnr_rmse_rowset(c, id, shift, totals, ps))) return nullptr;
  napi_value o, v;
  napi_create_object(env, &o);
  napi_create_double(env, totals[0], &v); napi_set_named_property(env, o, "rSumDiff2", v);
  napi_create_double(env, totals[1], &v); napi_set_named_property(env, o, "rCnt", v);
  napi_create_double(env, totals[2], &v); napi_set_named_property(env, o, "rSum", v);
  return o;
}

// downloadFactors(handle, which): device replica -> the attached host segment (bulk mode keeps factors on the device)
napi_value DownloadFactors(napi_env env, napi_callback_info info) {
  napi_value a[2];
  ycnr_ctx* c;
  int32_t which;
  if (!args(env, info, 2, a) || !handle(env, a[0], &c) || napi_get_value_int32(env, a[1], &which) != napi_ok) return nullptr;
  check(env, ycnr_download_factors(c, which, 0, -1));
  return undefined(env);
}

// hostRegister(handle, typedArray): page-lock a portion-cache buffer (usePortionsCache) for direct DMA
napi_value HostRegister(napi_env env, napi_callback_info info) {
  napi_value a[2];
  ycnr_ctx* c;
  if (!args(env, info, 2, a) || !handle(env, a[0], &c)) return nullptr;
  napi_typedarray_type t;
  size_t len = 0;
  void* p = nullptr;
  if (napi_get_typedarray_info(env, a[1], &t, &len, &p, nullptr, nullptr) != napi_ok || !p) {
    fail(env, "invalid type!");
    return nullptr;
  }
  const size_t elem = (t == napi_float64_array) ? 8 : (t == napi_int8_array || t == napi_uint8_array || t == napi_uint8_clamped_array) ? 1
                      : (t == napi_int16_array || t == napi_uint16_array) ? 2 : 4;
  check(env, ycnr_host_register(c, p, len * elem));
  return undefined(env);
}

// recommend(handle, Int32Array userIds, Float64Array skipPtr /*[n+1] offsets*/, Int32Array skipIds, limit,
//           minRecommendRating, globalAvgShift, Int32Array outIds /*[n*(limit-1)]*/, Float64Array outPredict,
//           Int32Array outCount /*[n]*/)      — YcnrController.recommendItemsForUser (YcnrController.js:227-284), 0-based ids
napi_value Recommend(napi_env env, napi_callback_info info) {
  napi_value a[10];
  ycnr_ctx* c;
  int32_t *uids, *skip, *oids, *ocnt, limit;
  double *sptr, *opred, min_rating, shift;
  size_t nu, nsp, ns, noi, nop, noc;
  if (!args(env, info, 10, a) || !handle(env, a[0], &c) || !typed(env, a[1], napi_int32_array, &uids, &nu) ||
      !typed(env, a[2], napi_float64_array, &sptr, &nsp) || !typed(env, a[3], napi_int32_array, &skip, &ns) ||
      napi_get_value_int32(env, a[4], &limit) != napi_ok || napi_get_value_double(env, a[5], &min_rating) != napi_ok ||
      napi_get_value_double(env, a[6], &shift) != napi_ok || !typed(env, a[7], napi_int32_array, &oids, &noi) ||
      !typed(env, a[8], napi_float64_array, &opred, &nop) || !typed(env, a[9], napi_int32_array, &ocnt, &noc))
    return nullptr;
  if (limit < 1 || nsp != nu + 1 || noc < nu || noi < nu * (size_t)(limit - 1) || nop < nu * (size_t)(limit - 1)) {
    fail(env, "recommend: array sizes do not match");
    return nullptr;
  }
  int64_t* ptr = new int64_t[nu + 1];   // JS has no Int64Array before BigInt: offsets travel as doubles
  for (size_t i = 0; i <= nu; ++i) ptr[i] = (int64_t)sptr[i];
  const int rc = (ptr[nu] < 0 || (size_t)ptr[nu] > ns)
                     ? -1
                     : ycnr_recommend_batch(c, (int32_t)nu, uids, ptr, skip, limit, min_rating, shift, oids, opred, ocnt);
  delete[] ptr;
  if (rc == -1) fail(env, "recommend: skipPtr exceeds skipIds");
  else check(env, rc);
  return undefined(env);
}

napi_value Destroy(napi_env env, napi_callback_info info) {
  // contexts are released by the external's finalizer; explicit destroy only synchronises
  napi_value a[1];
  ycnr_ctx* c;
  if (!args(env, info, 1, a) || !handle(env, a[0], &c)) return nullptr;
  check(env, ycnr_synchronize(c));
  return undefined(env);
}

napi_value Init(napi_env env, napi_value exports) {
  const napi_property_descriptor props[] = {
      {"create", nullptr, Create, nullptr, nullptr, nullptr, 0, nullptr},
      {"attachFactors", nullptr, AttachFactors, nullptr, nullptr, nullptr, 0, nullptr},
      {"startTrainStep", nullptr, StartTrainStep, nullptr, nullptr, nullptr, 0, nullptr},
      {"alsPortion", nullptr, AlsPortion, nullptr, nullptr, nullptr, 0, nullptr},
      {"alsPortions", nullptr, AlsPortions, nullptr, nullptr, nullptr, 0, nullptr},
      {"rmsePortionsAsync", nullptr, RmsePortionsAsync, nullptr, nullptr, nullptr, 0, nullptr},
      {"endTrainStep", nullptr, EndTrainStep, nullptr, nullptr, nullptr, 0, nullptr},
      {"startCalcRmse", nullptr, StartCalcRmse, nullptr, nullptr, nullptr, 0, nullptr},
      {"rmsePortion", nullptr, RmsePortion, nullptr, nullptr, nullptr, 0, nullptr},
      {"rmsePortionAsync", nullptr, RmsePortionAsync, nullptr, nullptr, nullptr, 0, nullptr},
      {"rmsePoll", nullptr, RmsePoll, nullptr, nullptr, nullptr, 0, nullptr},
      {"sAlsBuildSubFixedFacts", nullptr, SAlsBuildSubFixedFacts, nullptr, nullptr, nullptr, 0, nullptr},
      {"dAlsBuildSubFixedFacts", nullptr, DAlsBuildSubFixedFacts, nullptr, nullptr, nullptr, 0, nullptr},
      {"getMemoryUsage", nullptr, GetMemoryUsage, nullptr, nullptr, nullptr, 0, nullptr},
      {"tableUpload", nullptr, TableUpload, nullptr, nullptr, nullptr, 0, nullptr},
      {"tableSplit", nullptr, TableSplit, nullptr, nullptr, nullptr, 0, nullptr},
      {"tableCounts", nullptr, TableCounts, nullptr, nullptr, nullptr, 0, nullptr},
      {"rowsetFromTable", nullptr, RowsetFromTable, nullptr, nullptr, nullptr, 0, nullptr},
      {"alsRowset", nullptr, AlsRowset, nullptr, nullptr, nullptr, 0, nullptr},
      {"rmseRowset", nullptr, RmseRowset, nullptr, nullptr, nullptr, 0, nullptr},
      {"downloadFactors", nullptr, DownloadFactors, nullptr, nullptr, nullptr, 0, nullptr},
      {"hostRegister", nullptr, HostRegister, nullptr, nullptr, nullptr, 0, nullptr},
      {"recommend", nullptr, Recommend, nullptr, nullptr, nullptr, 0, nullptr},
      {"destroy", nullptr, Destroy, nullptr, nullptr, nullptr, 0, nullptr},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}

}  // namespace

NAPI_MODULE(cpp_utils, Init)
