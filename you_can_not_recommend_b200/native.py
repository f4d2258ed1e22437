"""ctypes binding of libycnr_als.so (include/ycnr_als.h) — the same C ABI the N-API shim binds.

Loading the library never touches the GPU; every compute call fails loudly (RuntimeError
carrying ycnr_last_error()) when no B200 is usable.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import build

BY_USER, BY_ITEM, RMSE_VALIDATE, RMSE_TEST = 0, 1, 2, 3
USER_FACTORS, ITEM_FACTORS = 0, 1
GRAM_AUTO, GRAM_FFMA, GRAM_TC3XTF32 = 0, 1, 2
STEP_TYPES = {"byUser": BY_USER, "byItem": BY_ITEM, "rmseValidate": RMSE_VALIDATE, "rmseTest": RMSE_TEST}

KERNEL_CLASSES = ["primal_fused", "dual_fused", "gram_partial", "reduce_solve", "rmse_rows", "rmse_reduce",
                  "gather", "gram_tc"]

# every symbol include/ycnr_als.h declares (tests/test_abi.py checks the export table)
EXPORTS = [
    "ycnr_last_error", "ycnr_device_count", "ycnr_create", "ycnr_destroy", "ycnr_attach_factors",
    "ycnr_upload_factors", "ycnr_download_factors", "ycnr_invalidate_device", "ycnr_device_factors",
    "ycnr_stream", "ycnr_synchronize", "ycnr_host_register", "ycnr_host_unregister", "ycnr_start_train_step", "ycnr_als_portion", "ycnr_end_train_step",
    "ycnr_start_calc_rmse", "ycnr_rmse_portion", "ycnr_rmse_portion_async", "ycnr_rmse_poll", "ycnr_als_portions",
    "ycnr_rmse_portions_async", "ycnr_s_als_build_sub_fixed_facts", "ycnr_s_als_build_sub_fixed_facts_noctx",
    "ycnr_check_portion", "ycnr_factor_elems", "ycnr_memory_usage", "ycnr_rowset_create",
    "ycnr_rowset_destroy", "ycnr_als_rowset", "ycnr_rmse_rowset", "ycnr_rmse_rowset_begin", "ycnr_rmse_rowset_ratings", "ycnr_ipc_export", "ycnr_ipc_import",
    "ycnr_ipc_close", "ycnr_set_peers", "ycnr_table_upload", "ycnr_table_split", "ycnr_table_counts", "ycnr_rowset_from_table",
    "ycnr_rowset_info", "ycnr_rowset_read", "ycnr_recommend_batch", "ycnr_debug_plan", "ycnr_debug_batch_rows", "ycnr_debug_read_partials", "ycnr_profile_reset", "ycnr_profile_read",
    "ycnr_profile_dual_bins",
]


class Options(C.Structure):
    _fields_ = [
        ("factors_count", C.c_int32), ("total_users", C.c_int32), ("total_items", C.c_int32),
        ("user_fact_reg", C.c_double), ("item_fact_reg", C.c_double),
        ("use_double_precision", C.c_int32), ("lowmem", C.c_int32), ("device", C.c_int32),
        ("gram_path", C.c_int32), ("dual_max_cols", C.c_int32), ("split_cols", C.c_int32),
        ("profile", C.c_int32), ("tc_min_cols", C.c_int32), ("tc_variant", C.c_int32),
        ("solve_chunks", C.c_int32), ("reserved", C.c_int32 * 1),
    ]


class PortionInfo(C.Structure):
    _fields_ = [
        ("rows_from", C.c_int32), ("rows_cnt", C.c_int32), ("ratings_in_portion", C.c_int64),
        ("time_ms", C.c_double), ("r_sum_diff2", C.c_double), ("r_cnt", C.c_double), ("r_sum", C.c_double),
    ]


class Profile(C.Structure):
    _fields_ = [
        ("ms", C.c_double * 8), ("launches", C.c_int64 * 8), ("rows", C.c_int64 * 8),
        ("ratings", C.c_int64 * 8), ("total_launches", C.c_int64),
    ]


_lib = None


def lib():
    """Build (if stale) and load libycnr_als.so."""
    global _lib
    if _lib is None:
        L = C.CDLL(build.build_cuda())
        L.ycnr_last_error.restype = C.c_char_p
        # per-message entry points: prototypes with plain address arguments, so the hot path passes the arrays'
        # addresses as ints (numpy's .ctypes.data_as costs ~1.5 us per argument)
        L.ycnr_check_portion.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
        L.ycnr_als_portion.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ycnr_rmse_portion_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.ycnr_rmse_poll.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("ycnr_als: " + lib().ycnr_last_error().decode())


def _addr(a):
    """Address of a numpy array's first element (int)."""
    return a.__array_interface__["data"][0]


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _i64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def debug_plan(row_len, factors_count=100, gram_path=GRAM_AUTO, dual_max_cols=-1, split_cols=0, tc_min_cols=0):
    """The library's launch plan for a row list (CPU only): dict with rows per dual bin, the fused / split row
    lists, the work items (row, offset) of the split rows and their processing order."""
    o = Options()
    o.factors_count, o.gram_path, o.dual_max_cols, o.split_cols, o.tc_min_cols = factors_count, gram_path, dual_max_cols, split_cols, tc_min_cols
    o.total_users = o.total_items = 1
    row_len = np.ascontiguousarray(row_len, np.int32)
    summ = np.zeros(32, np.int32)
    nw = C.c_int64(0)
    _check(lib().ycnr_debug_plan(C.byref(o), _i32(row_len), C.c_int32(len(row_len)), _i32(summ), None, C.c_int64(0), C.byref(nw)))
    words = np.zeros(max(1, nw.value), np.int32)
    _check(lib().ycnr_debug_plan(C.byref(o), _i32(row_len), C.c_int32(len(row_len)), _i32(summ), _i32(words), C.c_int64(len(words)), C.byref(nw)))
    n_items, n_multi, n_fused = int(summ[26]), int(summ[25]), int(summ[24])
    dual, off = [], 0
    for b in range(24):
        dual.append(words[off:off + int(summ[b])].copy())
        off += int(summ[b])
    return {"dual": dual, "fused": words[summ[27]:summ[27] + n_fused].copy(), "multi": words[summ[28]:summ[28] + n_multi].copy(),
            "item_row": words[summ[29]:summ[29] + n_items].copy(), "item_off": words[summ[30]:summ[30] + n_items].copy(),
            "item_order": words[summ[31]:summ[31] + n_items].copy()}


def debug_batch_rows(kind, headers, lim_rows, threads=4):
    """Row arrays (kind 1) / RMSE work entries (kind 2) the multi-portion entry points build from portion headers
    (CPU only): returns (counts[n, 3] = entries, ratings, bad; ids, len, start of the valid portions)."""
    n = len(headers)
    headers = [np.ascontiguousarray(h, np.int32) for h in headers]
    ptrs = (C.c_void_p * max(n, 1))(*[h.ctypes.data for h in headers])
    counts = np.zeros((n, 3), np.int64)
    cap = int(sum(int(h[0]) + int(np.maximum(h[2:2 * int(h[0]) + 1:2], 0).sum()) // 64 + 1 for h in headers if h[0] >= 0)) + 8
    ids, ln, st = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int64)
    e = C.c_int64(0)
    _check(lib().ycnr_debug_batch_rows(C.c_int32(kind), C.c_int32(n), ptrs, C.c_int64(lim_rows), C.c_int32(threads),
                                       counts.ctypes.data_as(C.POINTER(C.c_int64)), _i32(ids), _i32(ln),
                                       st.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(cap), C.byref(e)))
    return counts, ids[:e.value], ln[:e.value], st[:e.value]


def build_sub_fixed_facts_noctx(sub, fixed, indx, cols, k):
    """cpp_utils.sAlsBuildSubFixedFacts(sub, fixed, indx, cols, k) with upstream's own arity (cpp_utils.js:15-19)."""
    _check(lib().ycnr_s_als_build_sub_fixed_facts_noctx(_f32(sub), _f32(fixed), C.c_int64(fixed.size // k), _i32(indx),
                                                        C.c_int32(cols), C.c_int32(k)))


def device_count():
    n = C.c_int32(0)
    rc = lib().ycnr_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class Context:
    """Owns one ycnr_ctx (one GPU). Thin: argument marshalling only."""

    def __init__(self, factors_count, total_users, total_items, user_fact_reg=0.05, item_fact_reg=0.05,
                 use_double_precision=False, lowmem=False, device=0, gram_path=GRAM_AUTO, dual_max_cols=-1,
                 split_cols=0, profile=False, tc_min_cols=0, tc_variant=0, solve_chunks=0):
        o = Options()
        o.factors_count, o.total_users, o.total_items = factors_count, total_users, total_items
        o.user_fact_reg, o.item_fact_reg = user_fact_reg, item_fact_reg
        o.use_double_precision, o.lowmem, o.device = int(use_double_precision), int(lowmem), device
        o.gram_path, o.dual_max_cols, o.split_cols, o.profile = gram_path, dual_max_cols, split_cols, int(profile)
        o.tc_min_cols, o.tc_variant, o.solve_chunks = tc_min_cols, tc_variant, solve_chunks
        self._h = C.c_void_p()
        self.k, self.total_users, self.total_items = factors_count, total_users, total_items
        self._keep = []
        _check(lib().ycnr_create(C.byref(o), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().ycnr_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # -- factor store
    def attach_factors(self, user_factors, item_factors):
        for a, rows in ((user_factors, self.total_users), (item_factors, self.total_items)):
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape == (rows, self.k)
        self._keep = [user_factors, item_factors]
        _check(lib().ycnr_attach_factors(self._h, _f32(user_factors), _f32(item_factors)))

    def upload_factors(self, which):
        _check(lib().ycnr_upload_factors(self._h, C.c_int32(which)))

    def download_factors(self, which, row_from=0, row_cnt=-1):
        _check(lib().ycnr_download_factors(self._h, C.c_int32(which), C.c_int32(row_from), C.c_int32(row_cnt)))

    def invalidate_device(self, which):
        _check(lib().ycnr_invalidate_device(self._h, C.c_int32(which)))

    def device_factors_ptr(self, which):
        p = C.c_void_p()
        _check(lib().ycnr_device_factors(self._h, C.c_int32(which), C.byref(p)))
        return p.value

    def stream_ptr(self):
        p = C.c_void_p()
        _check(lib().ycnr_stream(self._h, C.byref(p)))
        return p.value or 0

    def synchronize(self):
        _check(lib().ycnr_synchronize(self._h))

    def host_register(self, array):
        """Page-lock a numpy array that holds portion buffers (direct DMA instead of staging)."""
        _check(lib().ycnr_host_register(self._h, C.c_void_p(array.ctypes.data), C.c_size_t(array.nbytes)))
        self._keep.append(array)

    def host_unregister(self, array):
        _check(lib().ycnr_host_unregister(self._h, C.c_void_p(array.ctypes.data)))

    # -- per-portion path
    def start_train_step(self, step_type):
        _check(lib().ycnr_start_train_step(self._h, C.c_int32(step_type)))

    def als_portion(self, rows, indx, vals):
        info = PortionInfo()
        L = lib()
        ra = _addr(rows)
        if L.ycnr_check_portion(ra, len(rows), len(indx), len(vals)) or \
                L.ycnr_als_portion(self._h, ra, _addr(indx), _addr(vals), C.addressof(info)):
            _check(1)
        return info

    def end_train_step(self):
        _check(lib().ycnr_end_train_step(self._h))

    def start_calc_rmse(self, step_type, global_avg_shift):
        _check(lib().ycnr_start_calc_rmse(self._h, C.c_int32(step_type), C.c_double(global_avg_shift)))

    def rmse_portion(self, rows, indx, vals):
        info = PortionInfo()
        _check(lib().ycnr_check_portion(_i32(rows), C.c_int64(len(rows)), C.c_int64(len(indx)), C.c_int64(len(vals))))
        _check(lib().ycnr_rmse_portion(self._h, _i32(rows), _i32(indx), _f32(vals), C.byref(info)))
        return info

    def rmse_portion_async(self, rows, indx, vals, tag):
        """Queue an RMSE portion under `tag`; results come back from rmse_poll in queue order."""
        L = lib()
        ra = _addr(rows)
        if L.ycnr_check_portion(ra, len(rows), len(indx), len(vals)) or \
                L.ycnr_rmse_portion_async(self._h, ra, _addr(indx), _addr(vals), tag):
            _check(1)

    def rmse_poll(self, wait=False, max_out=4096):
        """[(tag, PortionInfo)] of completed RMSE portions, oldest first; wait=True flushes and waits for all."""
        out = []
        buf = getattr(self, "_poll_buf", None)
        if buf is None or len(buf[0]) != max_out:       # landing arrays are reused: a poll per message must stay cheap
            buf = self._poll_buf = ((C.c_int64 * max_out)(), (PortionInfo * max_out)(), C.c_int32(0))
        tags, infos, n = buf
        L = lib()
        while True:
            if L.ycnr_rmse_poll(self._h, int(wait), max_out, tags, infos, C.byref(n)):
                _check(1)
            out.extend((tags[i], PortionInfo.from_buffer_copy(infos[i])) for i in range(n.value))
            if n.value < max_out:
                return out

    _PORTION_DTYPE = None

    def rmse_poll_arrays(self, wait=False, max_out=4096):
        """The same as rmse_poll, as arrays: (tags int64[n], infos structured[n] with the PortionInfo fields) — what a
        binding hands to a master that accumulates a whole pass at once (EmfMaster.wm_completedPortions)."""
        import numpy as np
        if Context._PORTION_DTYPE is None:
            Context._PORTION_DTYPE = np.dtype([("rows_from", "<i4"), ("rows_cnt", "<i4"), ("ratings_in_portion", "<i8"),
                                               ("time_ms", "<f8"), ("r_sum_diff2", "<f8"), ("r_cnt", "<f8"), ("r_sum", "<f8")])
            assert Context._PORTION_DTYPE.itemsize == C.sizeof(PortionInfo)
        tags_all, infos_all = [], []
        while True:
            tags = np.empty(max_out, dtype=np.int64)
            infos = np.empty(max_out, dtype=Context._PORTION_DTYPE)
            n = C.c_int32(0)
            _check(lib().ycnr_rmse_poll(self._h, C.c_int32(int(wait)), C.c_int32(max_out), tags.ctypes.data_as(C.POINTER(C.c_int64)),
                                        infos.ctypes.data_as(C.POINTER(PortionInfo)), C.byref(n)))
            tags_all.append(tags[:n.value])
            infos_all.append(infos[:n.value])
            if n.value < max_out:
                break
        return (np.concatenate(tags_all), np.concatenate(infos_all)) if len(tags_all) > 1 else (tags_all[0], infos_all[0])

    @staticmethod
    def _ptr_array(arrays, ctype):
        return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])

    def prepare_portions(self, portions, tags=None):
        """Pointer tables for n filled portion buffers [(rows, indx, vals)] (kept alive by the returned object):
        what a native binding hands to ycnr_als_portions / ycnr_rmse_portions_async."""
        n = len(portions)
        for r, i, v in portions:
            _check(lib().ycnr_check_portion(_i32(r), C.c_int64(len(r)), C.c_int64(len(i)), C.c_int64(len(v))))
        return {"n": n, "keep": portions,
                "rows": self._ptr_array([p[0] for p in portions], C.c_int32),
                "indx": self._ptr_array([p[1] for p in portions], C.c_int32),
                "vals": self._ptr_array([p[2] for p in portions], C.c_float),
                "infos": (PortionInfo * max(n, 1))(),
                "tags": (C.c_int64 * max(n, 1))(*(tags if tags is not None else range(n)))}

    def als_portions(self, prepared):
        """One ycnr_als_portion call per prepared portion, issued from native code; returns the PortionInfo array."""
        if not isinstance(prepared, dict):
            prepared = self.prepare_portions(prepared)
        _check(lib().ycnr_als_portions(self._h, C.c_int32(prepared["n"]), prepared["rows"], prepared["indx"],
                                       prepared["vals"], prepared["infos"]))
        return prepared["infos"]

    def rmse_portions_async(self, prepared):
        if not isinstance(prepared, dict):
            prepared = self.prepare_portions(prepared)
        _check(lib().ycnr_rmse_portions_async(self._h, C.c_int32(prepared["n"]), prepared["rows"], prepared["indx"],
                                              prepared["vals"], prepared["tags"]))

    def build_sub_fixed_facts(self, sub, fixed, indx):
        k = fixed.shape[1]
        _check(lib().ycnr_s_als_build_sub_fixed_facts(self._h, _f32(sub), _f32(fixed), C.c_int64(fixed.shape[0]),
                                                      _i32(indx), C.c_int32(len(indx)), C.c_int32(k)))

    def memory_usage(self):
        """'getMemoryUsage': dict of device / page-locked bytes held by the context and free / total device memory."""
        out = (C.c_int64 * 4)()
        _check(lib().ycnr_memory_usage(self._h, out))
        return {"device": out[0], "pinned": out[1], "deviceFree": out[2], "deviceTotal": out[3]}

    # -- bulk path
    def rowset_create(self, step_type, row_ids, row_start, row_len, indx, vals, portion_first=None):
        rid = C.c_int32(-1)
        npor = 0 if portion_first is None else len(portion_first) - 1
        _check(lib().ycnr_rowset_create(
            self._h, C.c_int32(step_type), C.c_int32(len(row_ids)), _i32(row_ids), _i64(row_start), _i32(row_len),
            _i32(indx), _f32(vals), C.c_int64(len(indx)),
            None if portion_first is None else _i32(portion_first), C.c_int32(npor), C.byref(rid)))
        return rid.value

    def rowset_destroy(self, rowset):
        _check(lib().ycnr_rowset_destroy(self._h, C.c_int32(rowset)))

    def als_rowset(self, rowset):
        _check(lib().ycnr_als_rowset(self._h, C.c_int32(rowset)))

    def rmse_rowset_ratings(self, rowset):
        tot, last = C.c_double(0), C.c_double(0)
        _check(lib().ycnr_rmse_rowset_ratings(self._h, C.c_int32(rowset), C.byref(tot), C.byref(last)))
        return tot.value, last.value

    def rmse_rowset_begin(self, rowset, shift):
        _check(lib().ycnr_rmse_rowset_begin(self._h, C.c_int32(rowset), C.c_double(shift)))

    def rmse_rowset(self, rowset, shift, n_portions=0):
        totals = (C.c_double * 3)()
        psums = np.zeros((max(n_portions, 1), 3), np.float64)
        _check(lib().ycnr_rmse_rowset(self._h, C.c_int32(rowset), C.c_double(shift), totals,
                                      psums.ctypes.data_as(C.POINTER(C.c_double)) if n_portions else None))
        return (totals[0], totals[1], totals[2]), psums

    # -- device-side front end
    def table_upload(self, user_ptr, item_ids, ratings, dataset_type):
        assert user_ptr.dtype == np.int64 and item_ids.dtype == np.int32 and ratings.dtype == np.float32
        assert dataset_type.dtype == np.int8 and len(user_ptr) == self.total_users + 1
        _check(lib().ycnr_table_upload(self._h, _i64(user_ptr), _i32(item_ids), _f32(ratings),
                                       dataset_type.ctypes.data_as(C.POINTER(C.c_int8))))

    def table_split(self, seed, pcts, nnz):
        """First-time train/validate/test split on the device; returns the dataset_type column."""
        out = np.zeros(nnz, np.int8)
        p = (C.c_int32 * 3)(*[int(x) for x in pcts])
        _check(lib().ycnr_table_split(self._h, C.c_uint64(seed), p, out.ctypes.data_as(C.POINTER(C.c_int8))))
        return out

    def table_counts(self, set_mask, by_item):
        out = np.zeros(self.total_items if by_item else self.total_users, np.int32)
        _check(lib().ycnr_table_counts(self._h, C.c_uint32(set_mask), C.c_int32(int(by_item)), _i32(out)))
        return out

    def rowset_from_table(self, step_type, set_mask, portions_row_id_to, first_row=0):
        pto = np.ascontiguousarray(portions_row_id_to, np.int32)
        rid = C.c_int32(-1)
        _check(lib().ycnr_rowset_from_table(self._h, C.c_int32(step_type), C.c_uint32(set_mask), C.c_int32(first_row),
                                            _i32(pto) if len(pto) else None, C.c_int32(len(pto)), C.byref(rid)))
        return rid.value

    def rowset_info(self, rowset):
        n, span, p = C.c_int32(0), C.c_int64(0), C.c_int32(0)
        _check(lib().ycnr_rowset_info(self._h, C.c_int32(rowset), C.byref(n), C.byref(span), C.byref(p)))
        return n.value, span.value, p.value

    def rowset_read(self, rowset):
        """dict of the row set's device arrays (tests, diagnostics)."""
        n, span, p = self.rowset_info(rowset)
        out = {"row_ids": np.zeros(n, np.int32), "row_start": np.zeros(n, np.int64), "row_len": np.zeros(n, np.int32),
               "portion_first": np.zeros(p + 1, np.int32), "indx": np.zeros(span, np.int32), "vals": np.zeros(span, np.float32)}
        _check(lib().ycnr_rowset_read(self._h, C.c_int32(rowset), _i32(out["row_ids"]), _i64(out["row_start"]),
                                      _i32(out["row_len"]), _i32(out["portion_first"]), _i32(out["indx"]), _f32(out["vals"])))
        return out

    # -- multi-GPU
    def ipc_export(self, which):
        buf = (C.c_uint8 * 64)()
        _check(lib().ycnr_ipc_export(self._h, C.c_int32(which), buf))
        return bytes(buf)

    def ipc_import(self, handle):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(lib().ycnr_ipc_import(self._h, buf, C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        _check(lib().ycnr_ipc_close(self._h, C.c_void_p(ptr)))

    def set_peers(self, which, ptrs):
        arr = (C.c_void_p * max(1, len(ptrs)))(*ptrs)
        _check(lib().ycnr_set_peers(self._h, C.c_int32(which), C.c_int32(len(ptrs)), arr))

    # -- serving
    def recommend_batch(self, user_ids, skip_lists, limit=20, min_recommend_rating=0.0, global_avg_shift=0.0):
        """Top-N for a batch of users (0-based ids). skip_lists: one iterable of 0-based item ids per user.
        Returns a list (per user) of (item_id, predict) pairs, best first, at most limit-1 of them."""
        user_ids = np.ascontiguousarray(user_ids, np.int32)
        n = len(user_ids)
        ptr = np.zeros(n + 1, np.int64)
        for i, sl in enumerate(skip_lists):
            ptr[i + 1] = ptr[i] + len(sl)
        skip = np.zeros(max(1, int(ptr[-1])), np.int32)
        for i, sl in enumerate(skip_lists):
            skip[ptr[i]:ptr[i + 1]] = np.asarray(sl, np.int32)
        keep = max(limit - 1, 0)
        ids = np.zeros((n, max(keep, 1)), np.int32)
        pred = np.zeros((n, max(keep, 1)), np.float64)
        cnt = np.zeros(max(n, 1), np.int32)
        _check(lib().ycnr_recommend_batch(self._h, C.c_int32(n), _i32(user_ids), _i64(ptr), _i32(skip), C.c_int32(limit),
                                          C.c_double(min_recommend_rating), C.c_double(global_avg_shift),
                                          _i32(ids), pred.ctypes.data_as(C.POINTER(C.c_double)), _i32(cnt)))
        return [[(int(ids[u, j]), float(pred[u, j])) for j in range(int(cnt[u]))] for u in range(n)]

    def debug_read_partials(self, n_items, n_tiles):
        out = np.zeros((n_items, n_tiles, 4, 4), np.float32)
        _check(lib().ycnr_debug_read_partials(self._h, _f32(out), C.c_int64(out.size)))
        return out

    # -- measurement
    def profile_reset(self):
        _check(lib().ycnr_profile_reset(self._h))

    def profile_dual_bins(self):
        ms = (C.c_double * 24)()
        rows = (C.c_int64 * 24)()
        _check(lib().ycnr_profile_dual_bins(self._h, ms, rows))
        return [(mt + 1, ms[mt], rows[mt]) for mt in range(24) if rows[mt]]

    def profile_read(self):
        p = Profile()
        _check(lib().ycnr_profile_read(self._h, C.byref(p)))
        out = {name: dict(ms=p.ms[i], launches=p.launches[i], rows=p.rows[i], ratings=p.ratings[i])
               for i, name in enumerate(KERNEL_CLASSES)}
        out["total_launches"] = p.total_launches
        return out
