"""Shared test scaffolding: seeded problems, portion lists, error metrics."""
import numpy as np

from you_can_not_recommend_b200 import front_end as fe
from you_can_not_recommend_b200.emf_master import EmfMaster

STEPS = ("byUser", "byItem", "rmseValidate", "rmseTest")


def make_problem(shape="ml-100k", k=None, seed=fe.DEFAULT_SEED, options=None, **synth_kw):
    """Ratings table + split + portion plan (product front end) + seeded initial factors."""
    table = fe.synth_table(shape, seed=seed, **synth_kw)
    k = k or fe.SHAPES.get(shape, {}).get("factors", 20)
    opts = {"factorsCount": k, "seed": seed}
    opts.update(options or {})
    m = EmfMaster(table, opts)
    m.splitDataForTrain()
    U0 = fe.init_factors(table.users, k, 0, seed + 2)
    V0 = fe.init_factors(table.items, k, 1, seed + 2)
    return {"table": table, "master": m, "k": k, "seed": seed, "U0": U0, "V0": V0,
            "total_ratings_avg": m.stats["totalRatingsAvg"], "options": m.options}


def step_csr(prob, step):
    return prob["master"]._csr(step)


def oracle_portions(prob, steps=STEPS):
    """dict stepType -> list of (bufRows, bufIndx, bufVals) in the upstream wire format."""
    m = prob["master"]
    out = {}
    for step in steps:
        csr = m._csr(step)
        pto = m.portionsRowIdTo[step]
        plist = []
        for p in range(len(pto)):
            row_from = 0 if p == 0 else int(pto[p - 1])
            rows, indx, vals, fetched = fe.build_portion(csr, row_from, int(pto[p]),
                                                         m.maxRowsInPortion[step] + 1,
                                                         max(1, m.maxRatingsInPortion[step]))
            plist.append((rows, indx, vals))
        out[step] = plist
    return out


def portion_from_rows(row_ids, cols_per_row, vals_per_row):
    """Hand-built portion in the wire format."""
    R = len(row_ids)
    rows = np.zeros(2 * R + 1, np.int32)
    rows[0] = R
    indx, vals = [], []
    for r in range(R):
        rows[1 + 2 * r] = row_ids[r]
        rows[2 + 2 * r] = len(cols_per_row[r])
        indx.extend(cols_per_row[r])
        vals.extend(vals_per_row[r])
    return rows, np.asarray(indx + [0], np.int32), np.asarray(vals + [0], np.float32)


def rel_fro(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def worst_row_rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    num = np.linalg.norm(a - b, axis=1)
    den = np.maximum(np.linalg.norm(b, axis=1), 1e-30)
    return float((num / den).max())
