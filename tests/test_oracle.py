"""The oracle against independent known answers (exact rationals, numpy float64) and its own frozen trajectory."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import make_problem, oracle_portions, portion_from_rows, rel_fro

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tiny_als_exact_rationals(dtype):
    g = json.load(open(os.path.join(GOLD, "tiny_als.json")))
    V = np.asarray(g["V"], dtype)
    U = np.zeros((3, 2), dtype)
    ids = sorted(int(u) for u in g["users"])
    cols = [[i for i, _ in g["users"][str(u)]["ratings"]] for u in ids]
    vals = [[r for _, r in g["users"][str(u)]["ratings"]] for u in ids]
    rows, indx, v = portion_from_rows(ids, cols, vals)
    n = oracle.als_portion(rows, indx, v, V, U, g["lambda"])
    assert n == sum(len(c) for c in cols)
    for u in ids:
        want = np.asarray(g["users"][str(u)]["x"])
        assert np.allclose(U[u], want, rtol=2e-6 if dtype == np.float32 else 1e-13)


def test_zero_cols_row_is_skipped():
    V = np.eye(3, dtype=np.float32)
    U = np.full((2, 3), 7.0, np.float32)
    rows, indx, v = portion_from_rows([1], [[]], [[]])     # the Q2 "A" pattern
    oracle.als_portion(rows, indx, v, V, U, 0.05)
    assert (U == 7.0).all()


def test_o32_loops_blas_and_o64_agree():
    rng = np.random.default_rng(0)
    k, items = 24, 50
    V = rng.normal(0, 0.3, (items, k))
    lens = [1, 3, 23, 24, 25, 50]
    cols = [sorted(rng.choice(items, n, replace=False).tolist()) for n in lens]
    vals = [rng.integers(1, 6, n).astype(float).tolist() for n in lens]
    rows, indx, v = portion_from_rows(list(range(len(lens))), cols, vals)
    U64 = np.zeros((len(lens), k))
    oracle.als_portion(rows, indx, v, V, U64, 0.05)
    # independent numpy float64 normal equations
    for r, n in enumerate(lens):
        Y = V[cols[r]]
        x = np.linalg.solve(Y.T @ Y + 0.05 * n * np.eye(k), Y.T @ np.asarray(vals[r]))
        assert np.allclose(U64[r], x, rtol=1e-9, atol=1e-12)
    U32 = np.zeros((len(lens), k), np.float32)
    oracle.als_portion(rows, indx, v, V.astype(np.float32), U32, 0.05)
    assert rel_fro(U32, U64) < 2e-4
    if oracle.set_blas(threads=2):
        U32b = np.zeros((len(lens), k), np.float32)
        oracle.als_portion(rows, indx, v, V.astype(np.float32), U32b, 0.05, use_blas=True)
        assert rel_fro(U32b, U64) < 2e-4
        U64b = np.zeros((len(lens), k))
        oracle.als_portion(rows, indx, v, V, U64b, 0.05, use_blas=True)
        assert rel_fro(U64b, U64) < 1e-10


def test_rmse_portion_matches_numpy():
    rng = np.random.default_rng(1)
    k = 10
    U = rng.normal(0, 0.5, (4, k)).astype(np.float32)
    V = rng.normal(0, 0.5, (6, k)).astype(np.float32)
    cols = [[0, 2, 5], [1], [], [3, 4]]
    vals = [[3, 4, 5], [1], [], [2, 2]]
    rows, indx, v = portion_from_rows([0, 1, 2, 3], cols, vals)
    d2, cnt, s = oracle.rmse_portion(rows, indx, v, U, V, 0.25)
    pred = [float(np.dot(U[u].astype(np.float64), V[i].astype(np.float64))) + 0.25 for u in range(4) for i in cols[u]]
    truth = [x for r in vals for x in r]
    assert cnt == 6
    assert abs(s - sum(pred)) < 1e-5
    assert abs(d2 - sum((t - p) ** 2 for t, p in zip(truth, pred))) < 1e-4


def test_frozen_c1_trajectory_and_precision_gap():
    """ML-100k-shaped config (BASELINE configs[0]): O64 reproduces its committed trajectory,
    O32 stays within the north-star tolerances of O64 (RMSE 1e-4, factors 1e-3)."""
    g = json.load(open(os.path.join(GOLD, "c1_trajectory.json")))
    prob = make_problem("ml-100k", k=20)
    portions = oracle_portions(prob)
    t64 = oracle.OracleTrainer(prob["U0"], prob["V0"], portions, 0.05, 0.05, prob["total_ratings_avg"], np.float64)
    t32 = oracle.OracleTrainer(prob["U0"], prob["V0"], portions, 0.05, 0.05, prob["total_ratings_avg"], np.float32)
    for it in range(3):
        h64, h32 = t64.train_iter(), t32.train_iter()
        for key in ("rmseValidate", "rmseTest", "rmseTestShift", "globalAvgShift"):
            assert abs(h64[key] - g["history"][it][key]) < 1e-9, (it, key)
            assert abs(h32[key] - h64[key]) < 1e-4, (it, key)
    assert rel_fro(t32.U, t64.U) < 1e-3 and rel_fro(t32.V, t64.V) < 1e-3


def test_q7_shift_uses_last_portion_only():
    prob = make_problem("ml-100k", k=8, options={"ratingsInPortionForRmse": 500})
    portions = oracle_portions(prob, steps=("rmseTest",))
    assert len(portions["rmseTest"]) > 2
    tr = oracle.OracleTrainer(prob["U0"], prob["V0"], portions, total_ratings_avg=3.0)
    tr.calc_rmse("rmseTest", False)
    rows, indx, vals = portions["rmseTest"][-1]
    _, cnt, s = oracle.rmse_portion(rows, indx, vals, tr.U, tr.V, 0.0)
    assert tr.global_avg_shift == 3.0 - s / cnt


def test_recommend_restatement_quirks():
    """YcnrController.recommendItemsForUser (lib/YcnrController.js:227-284) by hand: k = 1, user factor 1, so
    predict = item factor + shift.  At most limit-1 items come back (the pop at 281-282), best first, skip list
    and threshold respected, ties keep the lower item id."""
    U = np.asarray([[1.0]], np.float32)
    V = np.asarray([[3.0], [5.0], [1.0], [5.0], [4.0], [2.0]], np.float32)
    rec = oracle.recommend_items_for_user(U, V, 0, [], limit=4, min_recommend_rating=0.0, global_avg_shift=0.5)
    assert rec == [(1, 5.5), (3, 5.5), (4, 4.5)]                      # 3 = limit - 1 entries
    rec = oracle.recommend_items_for_user(U, V, 0, [1], limit=4, min_recommend_rating=2.6, global_avg_shift=0.5)
    assert rec == [(3, 5.5), (4, 4.5), (0, 3.5)]
    rec = oracle.recommend_items_for_user(U, V, 0, [1, 3], limit=10, min_recommend_rating=3.6, global_avg_shift=0.5)
    assert rec == [(4, 4.5)]
    assert oracle.recommend_items_for_user(U, V, 0, [], limit=1) == []
