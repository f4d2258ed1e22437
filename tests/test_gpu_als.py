"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle
on the same seeded inputs.  Tolerances are the north star's: factors <= 1e-3 relative,
RMSE <= 1e-4 per iteration; integer / copy work bit-exact."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import (make_problem, oracle_portions, portion_from_rows, rel_fro, worst_row_rel)
from you_can_not_recommend_b200 import front_end as fe
from you_can_not_recommend_b200 import native
from you_can_not_recommend_b200.emf_master import EmfMaster

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FACTOR_TOL = 1e-3     # north star: factors within 1e-3 relative error
RMSE_TOL = 1e-4       # north star: RMSE within 1e-4 per iteration


def run_portion(ctx, step, rows, indx, vals):
    ctx.start_train_step(step)
    info = ctx.als_portion(rows, indx, vals)
    ctx.end_train_step()
    return info


def random_rows(rng, lens, n_cols):
    cols = [np.sort(rng.choice(n_cols, n, replace=False)).tolist() for n in lens]
    vals = [rng.integers(1, 11, n).astype(float).tolist() for n in lens]
    return cols, vals


def test_tiny_golden_exact_rationals():
    g = json.load(open(os.path.join(GOLD, "tiny_als.json")))
    V = np.asarray(g["V"], np.float32)
    U = np.zeros((3, 2), np.float32)
    ids = sorted(int(u) for u in g["users"])
    cols = [[i for i, _ in g["users"][str(u)]["ratings"]] for u in ids]
    vals = [[r for _, r in g["users"][str(u)]["ratings"]] for u in ids]
    rows, indx, v = portion_from_rows(ids, cols, vals)
    for dual in (-1, 0):
        U[:] = 0
        ctx = native.Context(2, 3, 4, g["lambda"], g["lambda"], dual_max_cols=dual)
        ctx.attach_factors(U, V)
        info = run_portion(ctx, native.BY_USER, rows, indx, v)
        assert info.rows_cnt == 3 and info.ratings_in_portion == 8 and info.rows_from == 0
        for u in ids:
            assert np.allclose(U[u], g["users"][str(u)]["x"], rtol=5e-6), (dual, u)
        ctx.close()


@pytest.mark.parametrize("k", [20, 100, 7, 32, 50, 64, 128])
@pytest.mark.parametrize("dual", [-1, 0])
@pytest.mark.parametrize("gram", [native.GRAM_FFMA, native.GRAM_AUTO])
def test_adversarial_row_lengths_vs_oracle(k, dual, gram):
    rng = np.random.default_rng(100 + k)
    n_fixed, n_solved = 3000, 64
    F = rng.normal(0, 0.3, (n_fixed, k)).astype(np.float32)
    lens = [1, 2, 3, 4, 5, 15, 16, 17, 31, 33, k - 1, k, k + 1, 2 * k, 95, 96, 97, 130, 500, 2999]
    lens = [n for n in lens if n >= 1]
    ids = list(range(3, 3 + 2 * len(lens), 2))            # gaps between solved rows
    cols, vals = random_rows(rng, lens, n_fixed)
    rows, indx, v = portion_from_rows(ids, cols, vals)
    S0 = rng.normal(0, 0.1, (n_solved, k)).astype(np.float32)
    S64 = S0.astype(np.float64)
    oracle.als_portion(rows, indx, v, F.astype(np.float64), S64, 0.05)
    S32 = S0.copy()
    oracle.als_portion(rows, indx, v, F, S32, 0.05)
    for step in (native.BY_USER, native.BY_ITEM):
        S = S0.copy()
        if step == native.BY_USER:
            ctx = native.Context(k, n_solved, n_fixed, 0.05, 0.05, dual_max_cols=dual, gram_path=gram)
            ctx.attach_factors(S, F)
        else:
            ctx = native.Context(k, n_fixed, n_solved, 0.05, 0.05, dual_max_cols=dual, gram_path=gram)
            ctx.attach_factors(F, S)
        info = run_portion(ctx, step, rows, indx, v)
        assert info.ratings_in_portion == sum(lens)
        untouched = np.setdiff1d(np.arange(n_solved), ids)
        assert (S[untouched] == S0[untouched]).all()       # rows not in the portion keep their bytes
        assert worst_row_rel(S[ids], S64[ids]) < FACTOR_TOL
        assert worst_row_rel(S[ids], S32[ids]) < FACTOR_TOL
        assert rel_fro(S[ids], S64[ids]) < 2e-4
        ctx.close()


@pytest.mark.parametrize("k,split", [(100, 64), (100, 96), (20, 32), (64, 200)])
def test_split_rows_match_unsplit(k, split):
    rng = np.random.default_rng(5)
    n_fixed = 5000
    F = rng.normal(0, 0.3, (n_fixed, k)).astype(np.float32)
    lens = [split, split + 1, 2 * split, 2 * split + 1, 7 * split - 3, 4000, 40]
    ids = list(range(len(lens)))
    cols, vals = random_rows(rng, lens, n_fixed)
    rows, indx, v = portion_from_rows(ids, cols, vals)
    S64 = np.zeros((len(lens), k))
    oracle.als_portion(rows, indx, v, F.astype(np.float64), S64, 0.05)
    out = []
    for sc in (split, 0):
        S = np.zeros((len(lens), k), np.float32)
        ctx = native.Context(k, len(lens), n_fixed, 0.05, 0.05, split_cols=sc, profile=True,
                             gram_path=native.GRAM_FFMA)
        ctx.attach_factors(S, F)
        run_portion(ctx, native.BY_USER, rows, indx, v)
        prof = ctx.profile_read()
        if sc:
            assert prof["gram_partial"]["launches"] == 1 and prof["reduce_solve"]["rows"] >= 5
        ctx.close()
        assert worst_row_rel(S, S64) < FACTOR_TOL
        out.append(S)
    assert rel_fro(out[0], out[1]) < 1e-5


@pytest.mark.parametrize("k,split,tc_min", [(100, 4096, 0), (100, 64, 0), (100, 1000, 300), (64, 128, 0), (20, 64, 0), (124, 512, 0)])
def test_tensor_core_gram_matches_ffma_and_oracle(k, split, tc_min):
    """tcgen05 3xTF32 Gram (split-precision, fp32 accumulate in TMEM) vs the FFMA path and float64."""
    rng = np.random.default_rng(9)
    n_fixed = 7000
    F = rng.normal(0, 0.3, (n_fixed, k)).astype(np.float32)
    lens = [97, 128, 129, 255, 256, 257, 999, 4096, 4097, 6999]
    ids = list(range(len(lens)))
    cols, vals = random_rows(rng, lens, n_fixed)
    rows, indx, v = portion_from_rows(ids, cols, vals)
    S64 = np.zeros((len(lens), k))
    oracle.als_portion(rows, indx, v, F.astype(np.float64), S64, 0.05)
    out = {}
    for name, gram in (("ffma", native.GRAM_FFMA), ("tc", native.GRAM_TC3XTF32)):
        S = np.zeros((len(lens), k), np.float32)
        ctx = native.Context(k, len(lens), n_fixed, 0.05, 0.05, split_cols=split, profile=True, gram_path=gram,
                             tc_min_cols=tc_min)
        ctx.attach_factors(S, F)
        run_portion(ctx, native.BY_USER, rows, indx, v)
        prof = ctx.profile_read()
        assert (prof["gram_tc"]["launches"] > 0) == (name == "tc")
        ctx.close()
        assert worst_row_rel(S, S64) < FACTOR_TOL / 10, name
        out[name] = S
    assert worst_row_rel(out["tc"], out["ffma"]) < 1e-4


def test_zero_cols_rows_are_skipped_and_empty_portion():
    k = 20
    F = np.random.default_rng(0).normal(0, 0.3, (50, k)).astype(np.float32)
    S = np.full((4, k), 3.0, np.float32)
    ctx = native.Context(k, 4, 50)
    ctx.attach_factors(S, F)
    rows, indx, v = portion_from_rows([2], [[]], [[]])                 # Q2 pattern "A" -> (A, 0)
    info = run_portion(ctx, native.BY_USER, rows, indx, v)
    assert info.rows_cnt == 1 and info.ratings_in_portion == 0
    assert (S == 3.0).all()
    empty = np.zeros(3, np.int32)
    info = run_portion(ctx, native.BY_USER, empty, np.zeros(1, np.int32), np.zeros(1, np.float32))
    assert info.rows_cnt == 0 and info.rows_from == -1
    assert (S == 3.0).all()
    ctx.close()


def test_gather_is_bit_exact():
    rng = np.random.default_rng(2)
    for k in (20, 100, 7):
        fixed = rng.normal(0, 1, (500, k)).astype(np.float32)
        indx = rng.integers(0, 500, 77).astype(np.int32)
        sub = np.zeros((77, k), np.float32)
        ctx = native.Context(k, 500, 500)
        ctx.build_sub_fixed_facts(sub, fixed, indx)
        assert (sub == fixed[indx]).all()
        with pytest.raises(RuntimeError, match="out of range"):
            ctx.build_sub_fixed_facts(sub, fixed, np.asarray([500], np.int32))
        ctx.close()


@pytest.mark.parametrize("k", [20, 100, 7])
def test_rmse_portion_vs_oracle(k):
    rng = np.random.default_rng(3)
    U = rng.normal(0, 0.4, (40, k)).astype(np.float32)
    V = rng.normal(0, 0.4, (300, k)).astype(np.float32)
    lens = [0, 1, 2, 3, 4, 5, 9, 64, 257]
    ids = list(range(2, 2 + len(lens)))
    cols = [np.sort(rng.choice(300, n, replace=False)).tolist() for n in lens]
    vals = [rng.integers(1, 6, n).astype(float).tolist() for n in lens]
    rows, indx, v = portion_from_rows(ids, cols, vals)
    ctx = native.Context(k, 40, 300)
    ctx.attach_factors(U, V)
    for shift in (0.0, 0.37):
        want = oracle.rmse_portion(rows, indx, v, U.astype(np.float64), V.astype(np.float64), shift)
        ctx.start_calc_rmse(native.RMSE_TEST, shift)
        info = ctx.rmse_portion(rows, indx, v)
        assert info.r_cnt == want[1] == sum(lens)
        assert abs(info.r_sum_diff2 - want[0]) <= 1e-5 * want[0]
        assert abs(info.r_sum - want[2]) <= 1e-5 * abs(want[2]) + 1e-6
    ctx.close()


@pytest.mark.parametrize("registered", [False, True])
def test_rmse_portion_large_header_device_unpack(registered):
    """The RMSE path unpacks the raw header on the device (three-kernel prefix sum over the row lengths):
    70 000 rows = 35 scan blocks, ragged lengths incl. zeros, with plain and with page-locked (cached)
    portion buffers.  Counts are exact, sums match the oracle to fp64 summation-order noise."""
    k, users, items = 16, 70_000, 500
    rng = np.random.default_rng(11)
    U = rng.normal(0, 0.4, (users, k)).astype(np.float32)
    V = rng.normal(0, 0.4, (items, k)).astype(np.float32)
    lens = rng.integers(0, 7, users).astype(np.int32)
    lens[::997] = 40
    R = users
    rows = np.zeros(2 * R + 1, np.int32)
    rows[0] = R
    rows[1::2] = np.arange(R, dtype=np.int32)
    rows[2::2] = lens
    nnz = int(lens.sum())
    indx = rng.integers(0, items, nnz + 1).astype(np.int32)
    vals = rng.integers(1, 6, nnz + 1).astype(np.float32)
    want = oracle.rmse_portion(rows, indx, vals, U.astype(np.float64), V.astype(np.float64), 0.25)
    ctx = native.Context(k, users, items)
    ctx.attach_factors(U, V)
    if registered:
        for a in (rows, indx, vals):
            ctx.host_register(a)
    ctx.start_calc_rmse(native.RMSE_VALIDATE, 0.25)
    for _ in range(2):                               # slot reuse
        info = ctx.rmse_portion(rows, indx, vals)
        assert info.r_cnt == want[1] == nnz
        assert info.rows_from == 0 and info.rows_cnt == R and info.ratings_in_portion == nnz
        assert abs(info.r_sum_diff2 - want[0]) <= 1e-5 * want[0]
        assert abs(info.r_sum - want[2]) <= 1e-5 * abs(want[2]) + 1e-6
    ctx.close()


@pytest.mark.parametrize("bulk", [False, True])
def test_c1_ten_iteration_trajectory(bulk):
    """BASELINE configs[0]: ML-100k shape, k=20, 85/10/5, 10 iterations through the worker interface."""
    g = json.load(open(os.path.join(GOLD, "c1_trajectory.json")))
    prob = make_problem("ml-100k", k=20)
    o32 = oracle.OracleTrainer(prob["U0"], prob["V0"], oracle_portions(prob), 0.05, 0.05,
                               prob["total_ratings_avg"], np.float32)
    m = EmfMaster(prob["table"], {"factorsCount": 20, "seed": prob["seed"], "gpu": {"bulk": bulk}})
    m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
    for it in range(10):
        got = m.trainIter()
        want32 = o32.train_iter()
        want64 = g["history"][it]
        for key in ("rmseValidate", "rmseTest", "rmseTestShift", "globalAvgShift"):
            assert abs(got[key] - want64[key]) < RMSE_TOL, (it, key, got[key], want64[key])
            assert abs(got[key] - want32[key]) < RMSE_TOL, (it, key)
    if bulk:
        m.syncFactorsToHost()
    assert rel_fro(m.userFactors, o32.U) < FACTOR_TOL and rel_fro(m.itemFactors, o32.V) < FACTOR_TOL
    assert abs(float(np.abs(m.userFactors.astype(np.float64)).sum()) - g["U_checksum"]) < 1e-3 * g["U_checksum"]
    m.endTrain()


def test_bulk_cached_and_per_portion_agree_bitwise():
    """Three ways to feed the same step: work-buffer portions, cached page-locked portions (direct DMA),
    device-resident row sets.  Same kernels, same rows => identical bytes, identical RMSE sums."""
    rip = {"byUser": 3000, "byItem": 3000}
    prob = make_problem("ml-100k", k=20, options={"ratingsInPortionForAls": rip, "ratingsInPortionForRmse": 700})
    res = []
    for gpu in ({"bulk": False}, {"bulk": False, "cachePortions": True}, {"bulk": True}):
        m = EmfMaster(prob["table"], {"factorsCount": 20, "seed": prob["seed"], "ratingsInPortionForAls": rip,
                                      "ratingsInPortionForRmse": 700, "gpu": gpu})
        m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
        out = m.trainIter()
        if gpu["bulk"]:
            m.syncFactorsToHost()
        res.append((m.userFactors.copy(), m.itemFactors.copy(), out))
        m.endTrain()
    for other in res[1:]:
        assert (res[0][0] == other[0]).all() and (res[0][1] == other[1]).all()
        for key in ("rmseValidate", "rmseTest", "rmseTestShift", "globalAvgShift"):
            assert abs(res[0][2][key] - other[2][key]) < 1e-12


def test_c2_two_iterations_k100():
    """BASELINE configs[1]: ML-1M shape, k=100 (first iterations; O64 through LAPACK as referee)."""
    prob = make_problem("ml-1m", k=100, options={"ratingsInPortionForAls": {"byUser": 200000, "byItem": 200000},
                                                  "ratingsInPortionForRmse": 50000})
    have_blas = oracle.set_blas(threads=0)
    ref = oracle.OracleTrainer(prob["U0"], prob["V0"], oracle_portions(prob), 0.05, 0.05,
                               prob["total_ratings_avg"], np.float64, use_blas=have_blas)
    m = EmfMaster(prob["table"], dict(prob["options"], gpu={"bulk": True, "profile": True}))
    m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
    for it in range(2):
        got, want = m.trainIter(), ref.train_iter()
        for key in ("rmseValidate", "rmseTest", "rmseTestShift", "globalAvgShift"):
            assert abs(got[key] - want[key]) < RMSE_TOL, (it, key, got[key], want[key])
    m.syncFactorsToHost()
    assert rel_fro(m.userFactors, ref.U) < FACTOR_TOL and rel_fro(m.itemFactors, ref.V) < FACTOR_TOL
    assert worst_row_rel(m.userFactors, ref.U) < 5 * FACTOR_TOL
    prof = m.ctx.profile_read()
    assert prof["total_launches"] > 0 and prof["dual_fused"]["rows"] > 0 and (prof["primal_fused"]["rows"] + prof["gram_tc"]["rows"]) > 0
    m.endTrain()


def test_rejected_options_and_errors():
    prob = make_problem("ml-100k", k=20)
    for bad in ({"useDoublePrecision": True}, {"lowmem": True}):
        m = EmfMaster(prob["table"], dict({"factorsCount": 20}, **bad))
        with pytest.raises(ValueError):
            m.prepareToTrain()
    ctx = native.Context(20, 10, 10)
    with pytest.raises(RuntimeError, match="start_train_step"):
        ctx.als_portion(np.zeros(3, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float32))
    U = np.zeros((10, 20), np.float32)
    V = np.zeros((10, 20), np.float32)
    ctx.attach_factors(U, V)
    with pytest.raises(RuntimeError, match="out of range"):
        ctx.rowset_create(native.BY_USER, np.asarray([0], np.int32), np.asarray([0], np.int64),
                          np.asarray([1], np.int32), np.asarray([99], np.int32), np.asarray([1.0], np.float32))
    ctx.close()


def test_full_size_mal_properties():
    """BASELINE configs[2] at full size (1.75M x 12.7K, 121M ratings, k=100): size-independent checks —
    normal-equation residuals of sampled rows in float64, run-to-run bit reproducibility, untouched rows."""
    prob = make_problem("mal", k=100, options={"gpu": {"bulk": True}})
    t = prob["table"]
    k = 100
    m = EmfMaster(t, dict(prob["options"]))
    m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
    rng = np.random.default_rng(0)

    def residuals(step, fixed, solved, lam, sample):
        ids, ln, _ = m.rowlists[step]
        csr = m._csr(step)
        worst = 0.0
        for r in sample:
            rid, n = int(ids[r]), int(ln[r])
            beg = int(csr.ptr[rid])
            Y = fixed[csr.idx[beg:beg + n]].astype(np.float64)
            b = Y.T @ csr.vals[beg:beg + n].astype(np.float64)
            x = solved[rid].astype(np.float64)
            res = Y.T @ (Y @ x) + lam * n * x - b
            worst = max(worst, float(np.linalg.norm(res) / np.linalg.norm(b)))
        return worst

    m.alsTrainStep("byUser")
    m.ctx.download_factors(native.USER_FACTORS)
    U1 = m.userFactors.copy()
    ids, ln, _ = m.rowlists["byUser"]
    order = np.argsort(ln)
    sample = np.concatenate([order[:8], order[-8:], rng.choice(len(ids), 48, replace=False)])
    assert residuals("byUser", prob["V0"], U1, 0.05, sample) < 2e-4
    m.alsTrainStep("byItem")
    m.ctx.download_factors(native.ITEM_FACTORS)
    V1 = m.itemFactors.copy()
    ids_i, ln_i, _ = m.rowlists["byItem"]
    order = np.argsort(ln_i)
    sample = np.concatenate([order[:4], order[-4:], rng.choice(len(ids_i), 8, replace=False)])
    assert residuals("byItem", U1, V1, 0.05, sample) < 2e-4
    # Q2: one user per portion lost its last rating; rows never emitted keep their init bytes
    emitted = np.zeros(t.users, bool)
    emitted[ids[ln > 0]] = True
    assert (U1[~emitted] == prob["U0"][~emitted]).all()
    # reproducibility: same inputs, same bits
    m.ctx.upload_factors(native.USER_FACTORS)          # host still holds U1
    m.itemFactors[...] = prob["V0"]
    m.ctx.upload_factors(native.ITEM_FACTORS)
    m.alsTrainStep("byUser")
    m.ctx.download_factors(native.USER_FACTORS)
    assert (m.userFactors == U1).all()
    tot, _ = m.ctx.rmse_rowset(m.rowsets["rmseValidate"], 0.0, 0)
    assert tot[1] == int(m.rowlists["rmseValidate"][1].sum(dtype=np.int64))
    assert np.isfinite(tot[0]) and tot[0] > 0
    m.endTrain()


@pytest.mark.parametrize("k,limit", [(20, 20), (100, 5), (7, 1)])
def test_recommend_batch_vs_oracle(k, limit):
    """Top-N serving (YcnrController.js:227-284): same items in the same order as the literal restatement of
    the controller loop, predictions within fp32 summation-order noise (1e-5 relative), including the
    upstream quirk that at most limit-1 items come back, the skip list and the threshold."""
    users, items = 50, 700
    rng = np.random.default_rng(5)
    U = rng.normal(0.6, 0.3, (users, k)).astype(np.float32)
    V = rng.normal(0.6, 0.3, (items, k)).astype(np.float32)
    ctx = native.Context(k, users, items)
    ctx.attach_factors(U, V)
    uids = [0, 7, 49, 7]
    skips = [rng.choice(items, n, replace=False).astype(np.int32) for n in (0, 150, 699, 3)]
    shift = 0.123
    thr = float(np.median(U[7] @ V.T)) + shift          # a threshold that cuts about half of the items
    for min_rating in (-1e30, thr):
        got = ctx.recommend_batch(uids, skips, limit, min_rating, shift)
        for u, sk, rec in zip(uids, skips, got):
            want = oracle.recommend_items_for_user(U, V, u, sk, limit, min_rating, shift)
            assert len(rec) == len(want) <= max(limit - 1, 0)
            assert [i for i, _ in rec] == [i for i, _ in want]
            for (_, p), (_, q) in zip(rec, want):
                assert abs(p - q) <= 1e-5 * max(1.0, abs(q))
            assert not set(i for i, _ in rec) & set(sk.tolist())
    ctx.close()


def test_recommend_through_master_mirror():
    """recommendItemsForUser on the host mirror: 1-based ids, rated items skipped, after one ALS iteration."""
    prob = make_problem("ml-100k", k=20)
    m = EmfMaster(prob["table"], {"factorsCount": 20, "seed": prob["seed"], "gpu": {"bulk": True}})
    m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
    m.trainIter()
    m.syncFactorsToHost()
    user = {"list_id": 5, "unrated_items": [1, 2, 3]}
    rec = m.recommendItemsForUser(user, limit=10, minRecommendRating=1.0)
    t = prob["table"]
    rated = set((t.item_ids[t.user_ptr[4]:t.user_ptr[5]] + 1).tolist()) | {1, 2, 3}
    want = oracle.recommend_items_for_user(m.userFactors, m.itemFactors, 4, [i - 1 for i in rated], 10, 1.0,
                                           m.globalAvgShift)
    assert 0 < len(rec) <= 9 and [r["id"] for r in rec] == [i + 1 for i, _ in want]
    assert all(r["id"] not in rated and r["predict"] >= 1.0 for r in rec)
    assert all(a["predict"] >= b["predict"] for a, b in zip(rec, rec[1:]))
    m.endTrain()


def _host_rowset_arrays(m, step):
    """What the host front end hands to ycnr_rowset_create for a whole step."""
    from tests.helpers import step_csr  # noqa: F401
    csr = m._csr(step)
    pto = np.asarray(m.portionsRowIdTo[step], np.int32)
    rl = fe.build_rowlist(csr, pto)
    return csr, rl


@pytest.mark.parametrize("shape,kw", [("ml-100k", {}), ("sparse", dict(users=3000, items=400, ratings=9000))])
def test_device_ingest_matches_host_front_end_bitwise(shape, kw):
    """the device-built fetch (by user: compaction; by item: stable counting sort) and portion
    headers (quirk Q2) equal the host front end's arrays bit for bit, for all four step types; the planner's
    per-row counts too.  'sparse' has users/items without ratings in a set, 1-rating rows and tiny portions."""
    if shape == "sparse":
        table = fe.synth_table("ml-100k", seed=77, **kw)
        opts = {"factorsCount": 8, "seed": 77, "ratingsInPortionForRmse": 40,
                "ratingsInPortionForAls": {"byUser": 50, "byItem": 300}}
    else:
        table = fe.synth_table(shape)
        opts = {"factorsCount": 8}
    m = EmfMaster(table, opts)
    m.splitDataForTrain()
    ctx = native.Context(8, table.users, table.items)
    ctx.table_upload(table.user_ptr, table.item_ids, table.ratings, table.dataset_type)
    all_sets = (1 << fe.TRAIN) | (1 << fe.VALIDATE) | (1 << fe.TEST)
    assert (ctx.table_counts(all_sets, False) == table.counts_per_user()).all()
    assert (ctx.table_counts(all_sets, True) == table.counts_per_item()).all()
    from you_can_not_recommend_b200.emf_master import STEP_MASK
    for step in ("byUser", "byItem", "rmseValidate", "rmseTest"):
        csr, rl = _host_rowset_arrays(m, step)
        rid = ctx.rowset_from_table(native.STEP_TYPES[step], STEP_MASK[step], m.portionsRowIdTo[step])
        got = ctx.rowset_read(rid)
        assert len(got["indx"]) == csr.nnz
        assert (got["indx"] == csr.idx).all() and (got["vals"] == csr.vals).all(), step
        assert (got["row_ids"] == rl.row_ids).all() and (got["row_len"] == rl.row_len).all(), step
        assert (got["row_start"] == rl.row_start).all() and (got["portion_first"] == rl.portion_first).all(), step
        ctx.rowset_destroy(rid)
    ctx.close()


def test_device_ingest_trains_identically():
    """One ALS iteration from device-built row sets equals the host-built bulk path bitwise."""
    prob = make_problem("ml-100k", k=20)
    outs = []
    for dev in (False, True):
        m = EmfMaster(prob["table"], {"factorsCount": 20, "seed": prob["seed"], "gpu": {"bulk": True, "deviceIngest": dev}})
        m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
        h = m.trainIter()
        m.syncFactorsToHost()
        outs.append((h, m.userFactors.copy(), m.itemFactors.copy()))
        m.endTrain()
    assert outs[0][0] == outs[1][0]
    assert (outs[0][1] == outs[1][1]).all() and (outs[0][2] == outs[1][2]).all()


@pytest.mark.parametrize("pcts", [(85, 10, 5), (70, 30, 0), (100, 0, 0), (33, 33, 34)])
def test_device_split_matches_host_split_bitwise(pcts):
    """the per-user split rule of EmfLord.doSplitToSets (EmfLord.js:450-473) on the device gives
    the same dataset_type bytes as the host front end (same counter-based PRNG instead of Math.random)."""
    table = fe.synth_table("ml-100k", seed=123)
    want = fe.split_sets(table, pcts, seed=999).dataset_type.copy()
    ctx = native.Context(8, table.users, table.items)
    ctx.table_upload(table.user_ptr, table.item_ids, table.ratings, np.zeros(table.nnz, np.int8))
    got = ctx.table_split(999, pcts, table.nnz)
    assert (got == want).all()
    cu = ctx.table_counts((1 << fe.TRAIN) | (1 << fe.VALIDATE) | (1 << fe.TEST), False)
    assert (cu == np.diff(table.user_ptr)).all()          # the uploaded table's column was rewritten too
    ctx.close()


def test_checkpoint_resume_is_bitwise(tmp_path):
    """EmfManager mirror on the device path: save after one iteration (files written from the device replicas),
    load into a fresh master and continue — the second iteration equals an uninterrupted run bit for bit."""
    from you_can_not_recommend_b200.emf_manager import EmfManager
    prob = make_problem("ml-100k", k=20)
    opts = {"factorsCount": 20, "seed": prob["seed"], "checkpointEveryIter": True, "gpu": {"bulk": True}}
    a = EmfMaster(prob["table"], opts)
    a.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
    ha = a.train(2)
    Ua, Va = a.userFactors.copy(), a.itemFactors.copy()
    a.endTrain()
    b = EmfMaster(prob["table"], opts)
    b.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
    EmfManager(b, str(tmp_path / "ml")).train(1)
    b.endTrain()
    c = EmfMaster(prob["table"], opts)
    c.splitDataForTrain()
    U1, V1, ci = EmfManager(c, str(tmp_path / "ml")).loadCalcResults()
    assert ci["calcCnt"] == 1 and ci["factorsCount"] == 20      # one train run (EmfLord.js:904), however many checkpoints
    c.prepareToTrain(U1, V1)
    hc = c.trainIter()
    c.syncFactorsToHost()
    assert hc == ha[1]
    assert (c.userFactors == Ua).all() and (c.itemFactors == Va).all()
    c.endTrain()


def test_device_ingest_rank_slices_match_host_slices():
    """A rank's slice of the plan built on the device (first_row + its portion bounds) equals the same slice of the
    host front end's row list, for a 3-way cut of every step (what world_size = 3 would hand out)."""
    from you_can_not_recommend_b200 import dist as ydist
    from you_can_not_recommend_b200.emf_master import STEP_MASK
    table = fe.synth_table("ml-100k", seed=5)
    m = EmfMaster(table, {"factorsCount": 8, "ratingsInPortionForAls": {"byUser": 3000, "byItem": 3000},
                          "ratingsInPortionForRmse": 700})
    m.splitDataForTrain()
    ctx = native.Context(8, table.users, table.items)
    ctx.table_upload(table.user_ptr, table.item_ids, table.ratings, table.dataset_type)
    for step in ("byUser", "byItem", "rmseValidate", "rmseTest"):
        csr, rl = _host_rowset_arrays(m, step)
        pto = np.asarray(m.portionsRowIdTo[step], np.int32)
        cnt = ctx.table_counts(STEP_MASK[step], step == "byItem")
        assert (cnt == np.diff(csr.ptr)).all()
        cuts = ydist.balanced_cuts(np.cumsum(cnt, dtype=np.int64)[pto.astype(np.int64) - 1], 3)
        for r in range(3):
            lo, hi = int(cuts[r]), int(cuts[r + 1])
            if hi == lo:
                continue
            first_row = 0 if lo == 0 else int(pto[lo - 1])
            rid = ctx.rowset_from_table(native.STEP_TYPES[step], STEP_MASK[step], pto[lo:hi], first_row)
            got = ctx.rowset_read(rid)
            r0, r1 = int(rl.portion_first[lo]), int(rl.portion_first[hi])
            assert (got["row_ids"] == rl.row_ids[r0:r1]).all() and (got["row_len"] == rl.row_len[r0:r1]).all()
            assert (got["row_start"] == rl.row_start[r0:r1]).all()
            assert (got["portion_first"] == rl.portion_first[lo:hi + 1] - r0).all()
            ctx.rowset_destroy(rid)
    ctx.close()
