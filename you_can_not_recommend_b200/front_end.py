"""Host front end: the in-memory stand-in for the reference master's data side.

Produces what `EmfMaster`/`EmfLord` hand to the workers:
a ratings table with dataset_type, per-row stats, portion plans and portion
buffers in the upstream wire format.  All heavy loops run in libycnr_host.so
(csrc/host_frontend.cc, C ABI in include/ycnr_host.h) through ctypes.

Reference: lib/emf/EmfLord.js:48-128 (stats), 450-473 (split), 510-612 (planner);
lib/emf/EmfMaster.js:501-529 (fetch filters), 571-614 (portion conversion).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import build

_lib = None


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def lib():
    global _lib
    if _lib is None:
        path = build.build_host()
        L = C.CDLL(path)
        L.ycnr_host_last_error.restype = C.c_char_p
        L.ycnr_mix64.restype = C.c_uint64
        L.ycnr_mix64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.ycnr_u01.restype = C.c_double
        L.ycnr_u01.argtypes = [C.c_uint64]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("ycnr_host: " + lib().ycnr_host_last_error().decode())


# dataset_type values (EmfBase.js:229-247)
NOT_SPLIT, TRAIN, VALIDATE, TEST, NEW = 0, 1, 2, 3, 4
MASK_TRAIN = (1 << TRAIN) | (1 << VALIDATE)   # Q1: "IN (1, 2)"  EmfMaster.js:502-503
MASK_VALIDATE = 1 << VALIDATE
MASK_TEST = 1 << TEST

# The BASELINE.json shapes: users, items, ratings, max_rating, k
SHAPES = {
    "ml-100k": dict(users=943, items=1682, ratings=100_000, max_rating=5, factors=20),
    "ml-1m": dict(users=6040, items=3883, ratings=1_000_000, max_rating=5, factors=100),
    "mal": dict(users=1_750_000, items=12_700, ratings=121_000_000, max_rating=10, factors=100),
    "netflix": dict(users=480_000, items=17_800, ratings=100_000_000, max_rating=5, factors=100),
}
DEFAULT_SEED = 20261017


@dataclass
class Csr:
    """Rows x ratings in fetch order: ptr[rows+1] int64, idx int32 (0-based), vals float32."""
    ptr: np.ndarray
    idx: np.ndarray
    vals: np.ndarray

    @property
    def rows(self):
        return len(self.ptr) - 1

    @property
    def nnz(self):
        return int(self.ptr[-1])


@dataclass
class RowList:
    """Concatenated portion headers after the Q2 drop, addressing a Csr's idx/vals."""
    row_ids: np.ndarray        # int32 [R]
    row_start: np.ndarray      # int64 [R]
    row_len: np.ndarray        # int32 [R]
    portion_first: np.ndarray  # int32 [P+1]

    @property
    def nnz(self):
        return int(self.row_len.sum(dtype=np.int64))


@dataclass
class RatingsTable:
    """malrec_ratings (db-schema.sql:887-893), sorted by (user, item), ids 0-based."""
    users: int
    items: int
    user_ptr: np.ndarray       # int64 [users+1]
    item_ids: np.ndarray       # int32 [nnz]
    ratings: np.ndarray        # float32 [nnz] (smallint values)
    dataset_type: np.ndarray   # int8 [nnz]
    max_rating: int = 5
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def nnz(self):
        return int(self.user_ptr[-1])

    # -- stats (EmfLord.js:48-128; ratings_count over sets 1,2,3, EmfLord.js:281,345)
    def counts_per_user(self):
        m = (self.dataset_type >= 1) & (self.dataset_type <= 3)
        cs = np.zeros(self.nnz + 1, np.int64)
        np.cumsum(m, out=cs[1:])
        return (cs[self.user_ptr[1:]] - cs[self.user_ptr[:-1]]).astype(np.int32)

    def counts_per_item(self):
        m = (self.dataset_type >= 1) & (self.dataset_type <= 3)
        return np.bincount(self.item_ids[m], minlength=self.items).astype(np.int32)

    def avg_per_user(self):
        """avg_rating per user over sets 1,2,3 (malrec_users.avg_rating as doUpdateStats leaves it,
        EmfLord.js:255-397); NaN where the user has no such rating (upstream: row skipped -> undefined)."""
        m = (self.dataset_type >= 1) & (self.dataset_type <= 3)
        cs = np.zeros(self.nnz + 1, np.float64)
        np.cumsum(np.where(m, self.ratings, 0).astype(np.float64), out=cs[1:])
        s = cs[self.user_ptr[1:]] - cs[self.user_ptr[:-1]]
        c = self.counts_per_user()
        return np.where(c > 0, s / np.maximum(c, 1), np.nan)

    def avg_per_item(self):
        m = (self.dataset_type >= 1) & (self.dataset_type <= 3)
        s = np.bincount(self.item_ids[m], weights=self.ratings[m].astype(np.float64), minlength=self.items)
        c = self.counts_per_item()
        return np.where(c > 0, s / np.maximum(c, 1), np.nan)

    def total_ratings_avg(self):
        # "select avg(r.rating) where dataset_type in (1,2,3)"  EmfLord.js:224-228
        m = (self.dataset_type >= 1) & (self.dataset_type <= 3)
        return float(self.ratings[m].astype(np.float64).mean()) if m.any() else 0.0

    # -- fetch filters (EmfMaster.js:501-529)
    def csr_by_user(self, set_mask):
        key = ("u", set_mask)
        if key not in self._cache:
            L = lib()
            ptr = np.zeros(self.users + 1, np.int64)
            _check(L.ycnr_count_by_user(C.c_int32(self.users), _ptr(self.user_ptr, C.c_int64),
                                        _ptr(self.dataset_type, C.c_int8), C.c_uint32(set_mask),
                                        _ptr(ptr, C.c_int64)))
            idx = np.empty(int(ptr[-1]), np.int32)
            vals = np.empty(int(ptr[-1]), np.float32)
            _check(L.ycnr_fill_by_user(C.c_int32(self.users), _ptr(self.user_ptr, C.c_int64),
                                       _ptr(self.item_ids, C.c_int32), _ptr(self.ratings, C.c_float),
                                       _ptr(self.dataset_type, C.c_int8), C.c_uint32(set_mask),
                                       _ptr(ptr, C.c_int64), _ptr(idx, C.c_int32), _ptr(vals, C.c_float)))
            self._cache[key] = Csr(ptr, idx, vals)
        return self._cache[key]

    def csr_by_item(self, set_mask):
        key = ("i", set_mask)
        if key not in self._cache:
            L = lib()
            ptr = np.zeros(self.items + 1, np.int64)
            _check(L.ycnr_count_by_item(C.c_int32(self.users), C.c_int32(self.items),
                                        _ptr(self.user_ptr, C.c_int64), _ptr(self.item_ids, C.c_int32),
                                        _ptr(self.dataset_type, C.c_int8), C.c_uint32(set_mask),
                                        _ptr(ptr, C.c_int64)))
            idx = np.empty(int(ptr[-1]), np.int32)
            vals = np.empty(int(ptr[-1]), np.float32)
            _check(L.ycnr_fill_by_item(C.c_int32(self.users), C.c_int32(self.items),
                                       _ptr(self.user_ptr, C.c_int64), _ptr(self.item_ids, C.c_int32),
                                       _ptr(self.ratings, C.c_float), _ptr(self.dataset_type, C.c_int8),
                                       C.c_uint32(set_mask), _ptr(ptr, C.c_int64),
                                       _ptr(idx, C.c_int32), _ptr(vals, C.c_float)))
            self._cache[key] = Csr(ptr, idx, vals)
        return self._cache[key]


def synth_table(shape="ml-100k", seed=DEFAULT_SEED, alpha=1.2, item_skew=2.0, rank=8,
                users=None, items=None, ratings=None, max_rating=None, nthreads=0):
    """Deterministic synthetic ratings table of a named BASELINE shape (or explicit sizes)."""
    s = dict(SHAPES[shape]) if shape in SHAPES else {}
    users = users or s["users"]
    items = items or s["items"]
    ratings = ratings or s["ratings"]
    max_rating = max_rating or s.get("max_rating", 5)
    L = lib()
    counts = np.zeros(users, np.int32)
    _check(L.ycnr_synth_user_counts(C.c_uint64(seed), C.c_int32(users), C.c_int32(items),
                                    C.c_int64(ratings), C.c_double(alpha), _ptr(counts, C.c_int32)))
    user_ptr = np.zeros(users + 1, np.int64)
    np.cumsum(counts, out=user_ptr[1:])
    nnz = int(user_ptr[-1])
    item_ids = np.empty(nnz, np.int32)
    vals = np.empty(nnz, np.float32)
    _check(L.ycnr_synth_fill(C.c_uint64(seed), C.c_int32(users), C.c_int32(items),
                             _ptr(user_ptr, C.c_int64), C.c_int32(max_rating), C.c_int32(rank),
                             C.c_double(item_skew), _ptr(item_ids, C.c_int32), _ptr(vals, C.c_float),
                             C.c_int32(nthreads)))
    return RatingsTable(users, items, user_ptr, item_ids, vals, np.zeros(nnz, np.int8), max_rating)


def table_from_triples(users, items, u, i, r, max_rating=5):
    """Ratings table from explicit 0-based (user, item, rating) triples (tests, golden vectors)."""
    u = np.asarray(u, np.int64)
    i = np.asarray(i, np.int32)
    r = np.asarray(r, np.float32)
    order = np.lexsort((i, u))
    u, i, r = u[order], i[order], r[order]
    user_ptr = np.zeros(users + 1, np.int64)
    np.cumsum(np.bincount(u, minlength=users), out=user_ptr[1:])
    return RatingsTable(users, items, user_ptr, np.ascontiguousarray(i), np.ascontiguousarray(r),
                        np.zeros(len(r), np.int8), max_rating)


def split_sets(table, pcts=(85, 10, 5), seed=DEFAULT_SEED + 1, nthreads=0):
    """First-time split of every user's ratings by rule Q9 (EmfLord.js:450-473)."""
    p = (C.c_int32 * 3)(*pcts)
    _check(lib().ycnr_split_sets(C.c_uint64(seed), C.c_int32(table.users), _ptr(table.user_ptr, C.c_int64),
                                 p, _ptr(table.dataset_type, C.c_int8), C.c_int32(nthreads)))
    table._cache.clear()
    return table


def init_factors(rows, k, which, seed=DEFAULT_SEED + 2, mean=0.0, dev=None, nthreads=0):
    """EmfBase.initSharedFactorsRandom (EmfBase.js:457-513): N(mean, 1/k) per element, row-major."""
    out = np.empty((rows, k), np.float32)
    _check(lib().ycnr_init_factors(C.c_uint64(seed), C.c_int32(which), C.c_int64(rows * k),
                                   C.c_double(mean), C.c_double(1.0 / k if dev is None else dev),
                                   _ptr(out, C.c_float), C.c_int32(nthreads)))
    return out


def split_to_portions(cnt_per_row, ratings_in_portion, num_threads=1, pct_plus1=0):
    """EmfLord.splitToPortions (EmfLord.js:510-612) for one stepType.

    Returns (portionsRowIdTo int32[P], maxRatingsInPortion, maxRowsInPortion)."""
    cnt = np.ascontiguousarray(cnt_per_row, np.int32)
    cap = max(1, int((cnt > 0).sum()))
    out = np.zeros(cap, np.int32)
    n = C.c_int32(0)
    mr = C.c_int32(0)
    mrows = C.c_int32(0)
    _check(lib().ycnr_split_to_portions(_ptr(cnt, C.c_int32), C.c_int32(len(cnt)),
                                        C.c_int32(int(ratings_in_portion)), C.c_int32(int(num_threads)),
                                        C.c_int32(int(pct_plus1)), _ptr(out, C.c_int32), C.c_int32(cap),
                                        C.byref(n), C.byref(mr), C.byref(mrows)))
    return out[:n.value].copy(), mr.value, mrows.value


def build_portion_into(csr, row_from, row_to, rows, indx, vals):
    """EmfMaster.m_processFetchedPortionAlsOrRmse (EmfMaster.js:571-614): convert rows
    [row_from, row_to) of the fetch into the worker's portion buffers, in place.
    rows Int32[2*maxRows+1], indx Int32[maxRatings], vals Float32[maxRatings]. Returns data.length."""
    fetched = C.c_int32(0)
    _check(lib().ycnr_build_portion(_ptr(csr.ptr, C.c_int64), _ptr(csr.idx, C.c_int32), _ptr(csr.vals, C.c_float),
                                    C.c_int32(row_from), C.c_int32(row_to), _ptr(rows, C.c_int32),
                                    C.c_int32(len(rows)), _ptr(indx, C.c_int32), _ptr(vals, C.c_float),
                                    C.c_int32(len(indx)), C.byref(fetched)))
    return fetched.value


def build_portion(csr, row_from, row_to, max_rows, max_ratings):
    """Same, into fresh buffers: returns (rows, indx, vals, fetched)."""
    rows = np.zeros(2 * max_rows + 1, np.int32)
    indx = np.zeros(max_ratings, np.int32)
    vals = np.zeros(max_ratings, np.float32)
    return rows, indx, vals, build_portion_into(csr, row_from, row_to, rows, indx, vals)


def build_rowlist(csr, portions_row_id_to):
    """All portions of a step as one row list (bulk form of build_portion, same Q2 drop)."""
    pto = np.ascontiguousarray(portions_row_id_to, np.int32)
    cap = csr.rows + len(pto) + 1
    row_ids = np.empty(cap, np.int32)
    row_start = np.empty(cap, np.int64)
    row_len = np.empty(cap, np.int32)
    first = np.zeros(len(pto) + 1, np.int32)
    _check(lib().ycnr_build_rowlist(_ptr(csr.ptr, C.c_int64), _ptr(pto, C.c_int32), C.c_int32(len(pto)),
                                    _ptr(row_ids, C.c_int32), _ptr(row_start, C.c_int64),
                                    _ptr(row_len, C.c_int32), C.c_int32(cap), _ptr(first, C.c_int32)))
    n = int(first[-1])
    return RowList(row_ids[:n].copy(), row_start[:n].copy(), row_len[:n].copy(), first)
