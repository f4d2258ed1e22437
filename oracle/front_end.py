"""Pure-Python restatement of the reference master's data side (TEST INFRASTRUCTURE ONLY).

Literal, loop-for-loop restatements used to check the product front end
(you_can_not_recommend_b200/front_end.py -> csrc/host_frontend.cc) bit-exactly on
small inputs.  PARITY UNPINNED upstream: the reference ships no tests or fixtures
(package.json:30); the golden cases in tests/golden/ are hand-derived from the code
cited below.

  split rule Q9 ............ lib/emf/EmfLord.js:450-473 (+ knuth-shuffle's while-loop)
  planner Q6 ............... lib/emf/EmfLord.js:510-612
  fetch filter / order ..... lib/emf/EmfMaster.js:501-529
  portion conversion Q2 .... lib/emf/EmfMaster.js:571-614
"""
import math

import numpy as np

M64 = (1 << 64) - 1


def _splitmix(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def mix64(seed, a, b):
    """The seeded stand-in for Math.random (the reference is unseeded)."""
    return _splitmix((_splitmix((_splitmix(seed & M64) + a) & M64) + b) & M64)


def u01(h):
    return (h >> 11) * (1.0 / 9007199254740992.0)


def split_sets(seed, user_ptr, pcts=(85, 10, 5)):
    """EmfLord.doSplitToSets on a first split: every rating starts as dataset_type 0."""
    nnz = int(user_ptr[-1])
    out = np.zeros(nnz, np.int8)
    for u in range(len(user_ptr) - 1):
        beg, end = int(user_ptr[u]), int(user_ptr[u + 1])
        free_ids = list(range(end - beg))          # positions stand for item ids (ascending)
        if not free_ids:
            continue
        total = len(free_ids)
        target = [0, 0, 0]
        target[0] = math.ceil(total * pcts[0] / 100)
        target[1] = math.ceil(total * (pcts[0] + pcts[1]) / 100) - target[0]
        target[2] = total - (target[0] + target[1])
        new = [max(0, target[0]), max(0, target[1]), max(0, target[2])]
        if sum(new) < len(free_ids):
            new[0] += len(free_ids) - sum(new)
        # knuth-shuffle
        cur = len(free_ids)
        step = 0
        while cur != 0:
            r = math.floor(u01(mix64(seed, u, step)) * cur)
            step += 1
            cur -= 1
            free_ids[cur], free_ids[r] = free_ids[r], free_ids[cur]
        offs = 0
        for i in range(3):
            if new[i]:
                for p in free_ids[offs:offs + new[i]]:
                    out[beg + p] = i + 1
                offs += new[i]
    return out


def split_to_portions(cnt_per_row, ratings_in_portion_opt, num_threads_opt=1, pct_plus1=0):
    """EmfLord.splitToPortions for one stepType. cnt_per_row[id] == 0 means the id is a hole."""
    present = [(i, int(c)) for i, c in enumerate(cnt_per_row) if c != 0]
    rows_cnt = len(present)
    ratings_count = sum(c for _, c in present)
    max_per_row = max((c for _, c in present), default=0)
    if ratings_count == 0:
        return [], 0, 0
    if pct_plus1:
        ratings_count = math.ceil(ratings_count * (pct_plus1 / 100))
        max_per_row = math.ceil(max_per_row * (pct_plus1 / 100))
    ratings_in_portion = ratings_in_portion_opt
    avg_portions = math.ceil(ratings_count / ratings_in_portion)
    avg_rows = math.floor(rows_cnt / avg_portions)
    if avg_portions < num_threads_opt:
        avg_portions = num_threads_opt
        ratings_in_portion = math.ceil(ratings_count / avg_portions)
        avg_rows = math.floor(rows_cnt / avg_portions)
    if avg_rows < 1:
        avg_rows = 1
        avg_portions = rows_cnt
        ratings_in_portion = math.ceil(ratings_count / avg_portions)
    if ratings_in_portion < max_per_row:
        ratings_in_portion = max_per_row
        avg_portions = math.ceil(ratings_count / ratings_in_portion)
        avg_rows = math.floor(rows_cnt / avg_portions)
    portions = []
    p = rtgs = rows = max_rows = 0
    for idx, cnt in present:
        if pct_plus1:
            cnt = math.ceil(cnt * (pct_plus1 / 100))
        if rtgs + cnt > ratings_in_portion:
            rtgs = 0
            rows = 0
            p += 1
        rtgs += cnt
        rows += 1
        max_rows = max(max_rows, rows)
        while len(portions) <= p:
            portions.append(None)
        portions[p] = idx + 1
    return portions, int(ratings_in_portion), max_rows


def fetch(table_user_ptr, item_ids, ratings, dataset_type, step_type, row_from, row_to):
    """m_fetchPortionTrainAlsOrRmse: list of dicts {r, c, rating} with 1-based ids,
    WHERE id > row_from AND id <= row_to, dataset_type filter by stepType."""
    sel = {"rmseValidate": (2,), "rmseTest": (3,)}.get(step_type, (1, 2))
    data = []
    users = len(table_user_ptr) - 1
    if step_type == "byItem":
        for u in range(users):
            for e in range(int(table_user_ptr[u]), int(table_user_ptr[u + 1])):
                it = int(item_ids[e]) + 1
                if row_from < it <= row_to and int(dataset_type[e]) in sel:
                    data.append({"r": it, "c": u + 1, "rating": float(ratings[e])})
        data.sort(key=lambda d: (d["r"], d["c"]))  # ORDER BY item_id (column order: see Q4)
    else:
        for u in range(row_from, min(row_to, users)):
            for e in range(int(table_user_ptr[u]), int(table_user_ptr[u + 1])):
                if int(dataset_type[e]) in sel:
                    data.append({"r": u + 1, "c": int(item_ids[e]) + 1, "rating": float(ratings[e])})
    return data


def convert_portion(data, max_rows, max_ratings):
    """m_processFetchedPortionAlsOrRmse, loop kept verbatim in structure (incl. the Q2 drop)."""
    buf_rows = np.zeros(2 * max_rows + 1, np.int32)
    buf_indx = np.zeros(max_ratings, np.int32)
    buf_vals = np.zeros(max_ratings, np.float32)
    last_r = None
    r = 0
    cols = 0
    n = len(data)
    for i in range(n):
        d = dict(data[i])
        d["c"] -= 1
        d["r"] -= 1
        assert i < len(buf_vals)
        buf_vals[i] = d["rating"]
        buf_indx[i] = d["c"]
        if i == 0:
            last_r = d["r"]
        if last_r != d["r"] or i == n - 1:
            assert (1 + r * 2 + 1) < len(buf_rows)
            buf_rows[1 + r * 2] = last_r
            buf_rows[1 + r * 2 + 1] = cols
            last_r = d["r"]
            cols = 0
            r += 1
        cols += 1
    buf_rows[0] = r
    return buf_rows, buf_indx, buf_vals
