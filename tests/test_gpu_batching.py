"""Small portions (the reference default: 10 000 ratings per message) are queued inside the library and launched in
batches.  Whatever the batching, the bytes are the same: worker messages with refilled work buffers, cached
page-locked portions, the native call loop and the device-resident bulk path agree bit for bit."""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import make_problem, oracle_portions
from you_can_not_recommend_b200 import native
from you_can_not_recommend_b200.emf_master import EmfMaster

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flush", [0, 25_000, 300_000])
def test_batched_portion_paths_agree_bitwise(flush, monkeypatch):
    """ML-1M shape, k = 100, default 10k portions: ~100 portions per half-step, flushed every `flush` ratings
    (0 = library default: one batch per half-step here)."""
    if flush:
        monkeypatch.setenv("YCNR_BATCH_FLUSH", str(flush))
    prob = make_problem("ml-1m", k=100)
    res = []
    for gpu in ({"bulk": True}, {"bulk": False}, {"bulk": False, "cachePortions": True},
                {"bulk": False, "cachePortions": True, "nativeLoop": True}):
        m = EmfMaster(prob["table"], dict(prob["options"], gpu=gpu))
        m.prepareToTrain(prob["U0"].copy(), prob["V0"].copy())
        h = [m.trainIter() for _ in range(2)]
        if gpu["bulk"]:
            m.syncFactorsToHost()
        res.append((m.userFactors.copy(), m.itemFactors.copy(), h, m.completedPortions))
        m.endTrain()
    for U, V, h, _ in res[1:]:
        assert (U == res[0][0]).all() and (V == res[0][1]).all()
        for it in range(2):
            for key in ("rmseValidate", "rmseTest", "rmseTestShift", "globalAvgShift"):
                assert abs(h[it][key] - res[0][2][it][key]) < 1e-12, (it, key)
    assert res[1][3] == res[3][3] > 0          # same number of 'completedPortion' replies in the last pass


def test_async_rmse_portions_complete_in_queue_order(monkeypatch):
    """ycnr_rmse_portion_async / ycnr_rmse_poll: tags come back in the order they were queued, across several
    batches, each with the sums the synchronous call returns; a large portion in between takes the single-portion
    path and keeps its place in the order."""
    monkeypatch.setenv("YCNR_BATCH_FLUSH", "20000")
    prob = make_problem("ml-1m", k=100)
    m = prob["master"]
    por = oracle_portions(prob, ("rmseValidate",))["rmseValidate"]
    t = prob["table"]
    ctx = native.Context(100, t.users, t.items)
    ctx.attach_factors(prob["U0"], prob["V0"])
    ctx.start_calc_rmse(native.RMSE_VALIDATE, 0.25)
    sync = [ctx.rmse_portion(*p) for p in por]
    ctx.start_calc_rmse(native.RMSE_VALIDATE, 0.25)
    got = []
    for i, p in enumerate(por):
        ctx.rmse_portion_async(p[0], p[1], p[2], 1000 + i)
        got.extend(ctx.rmse_poll(False))
    got.extend(ctx.rmse_poll(True))
    assert [tag for tag, _ in got] == [1000 + i for i in range(len(por))]
    for (tag, a), b in zip(got, sync):
        assert (a.rows_from, a.rows_cnt, a.ratings_in_portion) == (b.rows_from, b.rows_cnt, b.ratings_in_portion)
        assert a.r_cnt == b.r_cnt and abs(a.r_sum_diff2 - b.r_sum_diff2) <= 1e-12 * b.r_sum_diff2
        assert abs(a.r_sum - b.r_sum) <= 1e-12 * abs(b.r_sum)
    assert ctx.rmse_poll(True) == []
    ctx.close()
    del m


def test_rmse_pass_with_another_shift_is_derived_without_kernels():
    """EmfLord.js:896-898: the third RMSE pass of an iteration sends the SAME test portions again, only the shift
    differs and the factors are untouched — the worker answers from the sums of the previous pass
    (sum (r-p-d)^2 = sum (r-p)^2 - 2d (sum r - sum p) + n d^2).  Same numbers as a fresh pass to 1e-12, no launch;
    after an ALS step the cache is stale and the kernels run again."""
    prob = make_problem("ml-100k", k=20)
    por = oracle_portions(prob, ("rmseTest",))["rmseTest"]
    t = prob["table"]
    U, V = prob["U0"].copy(), prob["V0"].copy()
    ctx = native.Context(20, t.users, t.items, profile=True)
    ctx.attach_factors(U, V)

    def one_pass(shift):
        ctx.start_calc_rmse(native.RMSE_TEST, shift)
        for i, p in enumerate(por):
            ctx.rmse_portion_async(p[0], p[1], p[2], i)
        return ctx.rmse_poll(True)

    ctx.start_calc_rmse(native.RMSE_TEST, 0.4)
    fresh = [ctx.rmse_portion(*p) for p in por]            # synchronous reference at the second shift
    a = one_pass(0.0)
    n0 = ctx.profile_read()["total_launches"]
    b = one_pass(0.4)
    assert ctx.profile_read()["total_launches"] == n0      # derived on the host
    assert [tag for tag, _ in b] == list(range(len(por)))
    for (_, got), want, (_, first) in zip(b, fresh, a):
        assert got.r_cnt == want.r_cnt == first.r_cnt and got.rows_cnt == want.rows_cnt
        assert abs(got.r_sum_diff2 - want.r_sum_diff2) <= 1e-12 * want.r_sum_diff2
        assert abs(got.r_sum - want.r_sum) <= 1e-12 * abs(want.r_sum)
    # factors change -> stale
    als = oracle_portions(prob, ("byUser",))["byUser"]
    ctx.start_train_step(native.BY_USER)
    ctx.als_portion(*als[0])
    ctx.end_train_step()
    n1 = ctx.profile_read()["total_launches"]
    c = one_pass(0.1)
    assert ctx.profile_read()["total_launches"] > n1 and len(c) == len(por)
    ctx.close()
