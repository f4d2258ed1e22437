"""GPU parity on the BASELINE configs the first round left to properties: the CUDA path (through the C ABI)
against the CPU oracle — O32 over OpenBLAS (the reference's routine sequence sgemm/sgemv/sgesv/sdot) and the
fp64 referee O64 — on portions sampled from the full-size shapes at the reference's DEFAULT portion size
(ratingsInPortionForAls = ratingsInPortionForRmse = 10 000, EmfBase.js:97-103).

Bars (north star): factors <= 1e-3 relative per matrix (Frobenius) and per row, RMSE sums to fp32-dot noise
(1e-5 relative per portion), RMSE <= 1e-4 per iteration; the row ids / ratings a portion covers are the
front end's (bit-exact, tests/test_front_end.py) and are shared by both sides.
"""
import os
import socket

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import make_problem, oracle_portions, rel_fro, worst_row_rel
from you_can_not_recommend_b200 import front_end as fe
from you_can_not_recommend_b200 import native
from you_can_not_recommend_b200.emf_master import EmfMaster

pytestmark = pytest.mark.gpu
FACTOR_TOL = 1e-3
RMSE_TOL = 1e-4


def portion_buffers(m, step, p):
    csr = m._csr(step)
    pto = m.portionsRowIdTo[step]
    row_from = 0 if p == 0 else int(pto[p - 1])
    rows, indx, vals, _ = fe.build_portion(csr, row_from, int(pto[p]), m.maxRowsInPortion[step] + 1,
                                           max(1, m.maxRatingsInPortion[step]))
    return rows, indx, vals


def compact(rows, indx, fixed):
    """Same portion over a compact copy of the rows of `fixed` it reads and a local solved matrix: the oracle
    gathers by index, so the arithmetic is unchanged while the fp64 copies stay small."""
    R = int(rows[0])
    ids = rows[1:1 + 2 * R:2].copy()
    lens = rows[2:2 + 2 * R:2]
    nnz = int(lens.sum(dtype=np.int64))
    uniq, inv = np.unique(indx[:nnz], return_inverse=True)
    rows_l = rows.copy()
    rows_l[1:1 + 2 * R:2] = np.arange(R, dtype=np.int32)
    indx_l = indx.copy()
    indx_l[:nnz] = inv.astype(np.int32)
    return ids, rows_l, indx_l, np.ascontiguousarray(fixed[uniq])


def check_sampled_half_step(m, step, fixed, solved_gpu, solved_before, portions, lam=0.05, with_o64=True):
    """Rows of the sampled portions: GPU result vs O32 (BLAS) and O64 on the same inputs."""
    k = fixed.shape[1]
    got, w32, w64 = [], [], []
    for p in portions:
        rows, indx, vals = portion_buffers(m, step, p)
        R = int(rows[0])
        if R == 0:
            continue
        ids, rows_l, indx_l, fx = compact(rows, indx, fixed)
        keep = rows[2:2 + 2 * R:2] > 0                      # Q2: (A, 0) rows are skipped on both sides
        s32 = np.ascontiguousarray(solved_before[ids])
        oracle.als_portion(rows_l, indx_l, vals, fx, s32, lam, use_blas=True)
        got.append(solved_gpu[ids][keep])
        w32.append(s32[keep])
        if with_o64:
            s64 = solved_before[ids].astype(np.float64)
            oracle.als_portion(rows_l, indx_l, vals, fx.astype(np.float64), s64, lam, use_blas=True)
            w64.append(s64[keep])
    got, w32 = np.concatenate(got), np.concatenate(w32)
    assert got.shape[0] > 0 and got.shape[1] == k
    assert worst_row_rel(got, w32) < FACTOR_TOL, (step, "O32 worst row")
    assert rel_fro(got, w32) < FACTOR_TOL / 4, (step, "O32 frobenius")
    if with_o64:
        w64 = np.concatenate(w64)
        assert worst_row_rel(got, w64) < FACTOR_TOL, (step, "O64 worst row")
        assert rel_fro(got, w64) < FACTOR_TOL / 4, (step, "O64 frobenius")
    return got.shape[0]


def check_sampled_rmse(m, step, U, V, shift, portions, psums):
    for p in portions:
        rows, indx, vals = portion_buffers(m, step, p)
        want = oracle.rmse_portion(rows, indx, vals, U, V, shift, use_blas=True)     # fp32 sdot, fp64 sums
        got = psums[p]
        assert got[1] == want[1], (step, p)
        assert abs(got[0] - want[0]) <= 1e-5 * max(want[0], 1e-30), (step, p, got[0], want[0])
        assert abs(got[2] - want[2]) <= 1e-5 * abs(want[2]) + 1e-6, (step, p)


def heaviest_portions(m, step, n):
    csr = m._csr(step)
    pto = np.asarray(m.portionsRowIdTo[step], np.int64)
    ends = csr.ptr[pto]
    sizes = np.diff(np.concatenate([[0], ends]))
    return [int(p) for p in np.argsort(sizes)[::-1][:n]]


def spread_portions(m, step, n):
    P = len(m.portionsRowIdTo[step])
    return sorted(set(int(x) for x in np.linspace(0, P - 1, min(n, P))))


def run_sampled_shape(shape, k, n_user, n_item, n_rmse, with_o64=True, **synth_kw):
    """One iteration in bulk mode at the reference's default portion sizes, compared on sampled portions."""
    assert oracle.set_blas(threads=1), "OpenBLAS (scipy) is needed for the O32 routine path"
    prob = make_problem(shape, k=k, options={"gpu": {"bulk": True}}, **synth_kw)
    m = EmfMaster(prob["table"], dict(prob["options"]))
    assert m.options["ratingsInPortionForAls"] == {"byUser": 10000, "byItem": 10000}
    assert m.options["ratingsInPortionForRmse"] == 10000
    U0, V0 = prob["U0"], prob["V0"]
    m.prepareToTrain(U0.copy(), V0.copy())
    m.alsTrainStep("byUser")
    m.ctx.download_factors(native.USER_FACTORS)
    U1 = m.userFactors.copy()
    n = check_sampled_half_step(m, "byUser", V0, U1, U0, spread_portions(m, "byUser", n_user), with_o64=with_o64)
    assert n >= n_user                                     # at least one row per sampled portion
    m.alsTrainStep("byItem")
    m.ctx.download_factors(native.ITEM_FACTORS)
    V1 = m.itemFactors.copy()
    check_sampled_half_step(m, "byItem", U1, V1, V0, heaviest_portions(m, "byItem", n_item) +
                            spread_portions(m, "byItem", 4), with_o64=with_o64)
    # the three RMSE passes' per-portion sums (Q7 needs them per portion)
    for step, shift in (("rmseValidate", 0.0), ("rmseTest", 0.0), ("rmseTest", 0.3125)):
        P = len(m.portionsRowIdTo[step])
        tot, ps = m.ctx.rmse_rowset(m.rowsets[step], shift, P)
        assert abs(tot[0] - ps[:, 0].sum()) <= 1e-9 * tot[0] and tot[1] == ps[:, 1].sum()
        check_sampled_rmse(m, step, U1, V1, shift, spread_portions(m, step, n_rmse), ps)
    m.endTrain()


def test_mal_default_portions_sampled_vs_oracle():
    """BASELINE configs[2] (the headline shape) at ratingsInPortionForAls = 10 000: >= 64 evenly spread byUser
    portions, the 8 heaviest + 4 spread byItem portions, 32 RMSE portions per pass."""
    run_sampled_shape("mal", 100, n_user=64, n_item=8, n_rmse=32)


def test_netflix_default_portions_sampled_vs_oracle():
    """BASELINE configs[3]: Netflix shape (480K x 17.8K, 100M ratings, k=100), sampled the same way."""
    run_sampled_shape("netflix", 100, n_user=64, n_item=6, n_rmse=24, with_o64=False)


@pytest.mark.parametrize("k", [32, 64, 128, 192, 256])
def test_k_sweep_mal_subsample_vs_oracle(k):
    """BASELINE configs[4] (k sweep 32/64/128/256) on a MAL-shaped subsample: 120 K users x 12.7 K items, 8.3 M
    ratings, same power-law row lengths; all three kernel classes (dual, tensor-core Gram — one pass up to k = 124,
    one pass per pair of column blocks above —, k x k solve) at every k."""
    run_sampled_shape("mal", k, n_user=48, n_item=6, n_rmse=12, users=120_000, items=12_700, ratings=8_300_000)


@pytest.mark.parametrize("k", [128, 132, 160, 192, 224, 256])
def test_wide_systems_adversarial_rows_vs_oracle(k):
    """k > 124 (the reference allocates A for any factorsCount, EmfWorker.js:200-211; its author recommends 300-400
    for MAL, EmfBase.js:90): dual rows, rows around k, multi-slice rows, both step types, vs O32 and O64."""
    assert oracle.set_blas(threads=1)
    rng = np.random.default_rng(200 + k)
    n_fixed, n_solved = 9000, 48
    F = rng.normal(0, 0.3, (n_fixed, k)).astype(np.float32)
    lens = [1, 5, 95, 96, 97, 127, 128, 129, k - 1, k, k + 1, 2 * k + 3, 4095, 4096, 4097, 8999]
    ids = list(range(2, 2 + 2 * len(lens), 2))
    cols = [np.sort(rng.choice(n_fixed, n, replace=False)).tolist() for n in lens]
    vals = [rng.integers(1, 11, n).astype(float).tolist() for n in lens]
    from tests.helpers import portion_from_rows
    rows, indx, v = portion_from_rows(ids, cols, vals)
    S0 = rng.normal(0, 0.1, (n_solved, k)).astype(np.float32)
    S64 = S0.astype(np.float64)
    oracle.als_portion(rows, indx, v, F.astype(np.float64), S64, 0.05, use_blas=True)
    S32 = S0.copy()
    oracle.als_portion(rows, indx, v, F, S32, 0.05, use_blas=True)
    for step in (native.BY_USER, native.BY_ITEM):
        S = S0.copy()
        if step == native.BY_USER:
            ctx = native.Context(k, n_solved, n_fixed, 0.05, 0.05, profile=True)
            ctx.attach_factors(S, F)
        else:
            ctx = native.Context(k, n_fixed, n_solved, 0.05, 0.05, profile=True)
            ctx.attach_factors(F, S)
        ctx.start_train_step(step)
        ctx.als_portion(rows, indx, v)
        ctx.end_train_step()
        prof = ctx.profile_read()
        nb = (k + 59) // 60
        assert prof["gram_tc"]["launches"] == nb * (nb - 1) // 2 and prof["reduce_solve"]["rows"] == len(lens) - 4
        untouched = np.setdiff1d(np.arange(n_solved), ids)
        assert (S[untouched] == S0[untouched]).all()
        assert worst_row_rel(S[ids], S64[ids]) < FACTOR_TOL and worst_row_rel(S[ids], S32[ids]) < FACTOR_TOL
        assert rel_fro(S[ids], S64[ids]) < 2e-4
        ctx.close()


def test_wide_system_option_errors():
    for kw, pat in ((dict(factors_count=130), "% 4"), (dict(factors_count=256, gram_path=native.GRAM_FFMA), "tensor-core"),
                    (dict(factors_count=260), "1..256")):
        with pytest.raises(RuntimeError, match=pat):
            native.Context(kw.pop("factors_count"), 10, 10, **kw)


@pytest.mark.parametrize("bulk", [False, True])
def test_c2_ten_iterations_default_portions(bulk):
    """BASELINE configs[1]: ML-1M shape, k=100, ALL 10 iterations at the default 10 000-rating portions, through
    the worker messages and in bulk mode, against O32 over OpenBLAS (RMSE per iteration, factors at the end)."""
    assert oracle.set_blas(threads=1)
    prob = make_problem("ml-1m", k=100)
    ref = oracle.OracleTrainer(prob["U0"], prob["V0"], oracle_portions(prob), 0.05, 0.05,
                               prob["total_ratings_avg"], np.float32, use_blas=True)
    m = EmfMaster(prob["table"], dict(prob["options"], gpu={"bulk": bulk}))
    assert m.options["ratingsInPortionForAls"]["byUser"] == 10000
    m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
    for it in range(10):
        got, want = m.trainIter(), ref.train_iter()
        for key in ("rmseValidate", "rmseTest", "rmseTestShift", "globalAvgShift"):
            assert abs(got[key] - want[key]) < RMSE_TOL, (it, key, got[key], want[key])
    if bulk:
        m.syncFactorsToHost()
    assert rel_fro(m.userFactors, ref.U) < FACTOR_TOL and rel_fro(m.itemFactors, ref.V) < FACTOR_TOL
    m.endTrain()


@pytest.mark.parametrize("k", [20, 100, 7])
def test_rmse_portion_vs_o32_and_o64(k):
    """mw_calcRmsePortion against BOTH oracles: O32 (fp32 sdot through BLAS, what the reference computes,
    EmfBase.js:825-827) and O64."""
    assert oracle.set_blas(threads=1)
    rng = np.random.default_rng(31)
    U = rng.normal(0, 0.4, (60, k)).astype(np.float32)
    V = rng.normal(0, 0.4, (900, k)).astype(np.float32)
    lens = [0, 1, 2, 3, 7, 33, 64, 257, 800]
    R = len(lens)
    rows = np.zeros(2 * R + 1, np.int32)
    rows[0] = R
    rows[1::2] = np.arange(5, 5 + R)
    rows[2::2] = lens
    nnz = sum(lens)
    indx = np.concatenate([np.sort(rng.choice(900, n, replace=False)) for n in lens] + [[0]]).astype(np.int32)
    vals = np.concatenate([rng.integers(1, 11, nnz), [0]]).astype(np.float32)
    ctx = native.Context(k, 60, 900)
    ctx.attach_factors(U, V)
    for shift in (0.0, -0.41):
        w32 = oracle.rmse_portion(rows, indx, vals, U, V, shift, use_blas=True)
        w32c = oracle.rmse_portion(rows, indx, vals, U, V, shift, use_blas=False)
        w64 = oracle.rmse_portion(rows, indx, vals, U.astype(np.float64), V.astype(np.float64), shift)
        ctx.start_calc_rmse(native.RMSE_VALIDATE, shift)
        info = ctx.rmse_portion(rows, indx, vals)
        for want in (w32, w32c, w64):
            assert info.r_cnt == want[1] == nnz
            assert abs(info.r_sum_diff2 - want[0]) <= 1e-5 * want[0]
            assert abs(info.r_sum - want[2]) <= 1e-5 * abs(want[2]) + 1e-6
    ctx.close()


# ---- two ranks on two GPUs: NCCL / NVLink exchange -------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _two_rank_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from you_can_not_recommend_b200 import dist as ydist
    ydist.init_from_env("nccl")
    torch.cuda.set_device(rank)
    prob = make_problem("ml-1m", k=100)
    out = {}
    for mode in ("single", "fused", "nccl"):
        if mode == "single" and rank != 0:
            continue
        w = 1 if mode == "single" else world
        m = EmfMaster(prob["table"], dict(prob["options"], gpu={"bulk": True, "device": rank}), rank=rank if w > 1 else 0,
                      world=w)
        m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
        if mode == "fused":
            m.connectPeers()
        h = m.trainIter()
        m.syncFactorsToHost()
        out[mode] = (h, m.userFactors.copy(), m.itemFactors.copy())
        m.endTrain()
        if w > 1:
            dist.barrier()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_single_gpu_bitwise():
    """World size 2 (one process per GPU, NCCL): after one iteration both replicas on both ranks equal the 1-GPU
    run bit for bit, for the fused NVLink peer stores and for the NCCL all-gather exchange; RMSE values equal
    to summation order (1e-12)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    h1, U1, V1 = res[0]["single"]
    for rank in (0, 1):
        for mode in ("fused", "nccl"):
            h, U, V = res[rank][mode]
            assert (U == U1).all() and (V == V1).all(), (rank, mode)
            for key in ("rmseValidate", "rmseTest", "rmseTestShift", "globalAvgShift"):
                assert abs(h[key] - h1[key]) < 1e-12, (rank, mode, key)
