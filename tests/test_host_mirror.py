"""Host-side mirror of the reference interface: option names/defaults, merging, message table."""
import numpy as np
import pytest

from you_can_not_recommend_b200 import emf_base, front_end as fe
from you_can_not_recommend_b200.emf_master import EmfMaster
from you_can_not_recommend_b200.emf_worker import EmfProcess, EmfWorker


def test_default_options_match_reference():
    o = emf_base.default_options()                       # EmfBase.js:52-140
    assert o["factorsCount"] == 100 and o["trainIters"] == 10 and o["alg"] == "als"
    assert o["als"] == {"userFactReg": 0.05, "itemFactReg": 0.05, "initFirstFactorAsAvgRating": False}
    assert o["dataSetDistr"] == [85, 10, 5]
    assert o["ratingsInPortionForAls"] == {"byUser": 10000, "byItem": 10000} and o["ratingsInPortionForRmse"] == 10000
    assert o["useDoublePrecision"] is False and o["lowmem"] is False and o["maxRating"] == {"mal": 10, "ml": 5}


def test_deepmerge_semantics():
    m = emf_base.deepmerge({"a": {"x": 1, "y": 2}, "b": [1, 2]}, {"a": {"y": 3}, "b": [9]})
    assert m == {"a": {"x": 1, "y": 3}, "b": [9]}
    b = EmfMaster(fe.synth_table("custom", users=20, items=10, ratings=100), {"als": {"userFactReg": 0.1}})
    assert b.options["als"] == {"userFactReg": 0.1, "itemFactReg": 0.05, "initFirstFactorAsAvgRating": False}


def test_worker_message_table_and_rejections():
    w = EmfWorker(0, EmfProcess())
    for msg in ("prepareToTrain", "startTrain", "endTrain", "startTrainStep", "startCalcRmse",
                "calcTrainAlsPortion", "calcTrainSgdPortion", "calcRmsePortion"):        # EmfWorker.js:43-51
        assert msg in w.process._handlers
    with pytest.raises(NotImplementedError):
        w.process.onMessage("calcTrainSgdPortion", {})
    with pytest.raises(ValueError):
        EmfWorker(0, EmfProcess(), {"alg": "sgd"})
    for bad in ({"useDoublePrecision": True}, {"lowmem": True}):
        b = EmfWorker(0, EmfProcess(), bad)
        b.totalUsersCount = b.totalItemsCount = 4
        b.createSharedFactors()
        with pytest.raises(ValueError):
            b.openDevice()


def test_master_plan_and_factor_layout():
    t = fe.synth_table("custom", users=50, items=30, ratings=600)
    m = EmfMaster(t, {"factorsCount": 8, "ratingsInPortionForAls": {"byUser": 100, "byItem": 100},
                      "ratingsInPortionForRmse": 20})
    m.splitDataForTrain()
    assert m.totalUsersCount == 50 and m.totalItemsCount == 30
    for step in ("byUser", "byItem", "rmseValidate", "rmseTest"):
        assert m.portionsCount[step] == len(m.portionsRowIdTo[step]) >= 1
    assert m.portionsRowIdTo["byUser"][-1] == 50
    m.createSharedFactors()
    m.initSharedFactorsRandom()
    assert m.userFactors.shape == (50, 8) and m.userFactors.dtype == np.float32 and m.userFactors.flags.c_contiguous
    assert abs(float(m.userFactors.std()) - 1 / 8) < 0.03          # randomNormal(1/k), EmfBase.js:486
    assert m._alsPredict(m.userFactors[0], m.itemFactors[0]) == float(np.dot(m.userFactors[0], m.itemFactors[0]))
    # the sliced ranges of a 2-rank run tile the id space
    m2 = [EmfMaster(t, m.options, rank=r, world=2) for r in range(2)]
    for x in m2:
        x.splitDataForTrain()
        x.my_portions = {s: x._slice_portions(s) for s in ("byUser", "byItem")}
    for s in ("byUser", "byItem"):
        a0, b0 = m2[0]._solved_range(s)
        a1, b1 = m2[1]._solved_range(s)
        assert a0 == 0 and b0 == a1 and b1 == (50 if s == "byUser" else m.portionsRowIdTo[s][-1])


def test_manager_persistence_roundtrip_cpu(tmp_path):
    """EmfManager mirror (EmfManager.js:158-191, 324-397, 463-570): raw Float32 row-major factor files +
    calc_info.json, commit by renaming factors_temp -> factors_ready, reuse rules, grown tables.
    No GPU involved: a stand-in master carries the fields the manager reads."""
    import json
    import numpy as np
    from you_can_not_recommend_b200.emf_base import default_options
    from you_can_not_recommend_b200.emf_manager import EmfManager

    class M:
        pass
    m = M()
    m.options = default_options()
    m.options["gpu"]["bulk"] = False
    m.ctx = None
    m.factorsCount, m.totalUsersCount, m.totalItemsCount, m.globalAvgShift = 4, 5, 3, 0.25
    rng = np.random.default_rng(0)
    m.userFactors = rng.normal(size=(5, 4)).astype(np.float32)
    m.itemFactors = rng.normal(size=(3, 4)).astype(np.float32)
    mgr = EmfManager(m, str(tmp_path / "ml"))
    assert mgr.loadCalcResults() is None
    info = mgr.saveCalcResults()
    ready = tmp_path / "ml_factors_ready"
    assert sorted(p.name for p in ready.iterdir()) == ["calc_info.json", "item_factors", "user_factors"]
    assert not (tmp_path / "ml_factors_temp").exists()                       # renamed, not copied
    assert (ready / "user_factors").stat().st_size == 5 * 4 * 4             # EmfBase.js:694-697: no header
    assert np.fromfile(ready / "item_factors", np.float32).tobytes() == m.itemFactors.tobytes()
    ci = json.load(open(ready / "calc_info.json"))
    assert ci == info and set(ci) == {"alg", "algOptions", "useDoublePrecision", "factorsCount", "dataSetDistr",
                                      "totalUsersCount", "totalItemsCount", "dbType", "calcDate", "calcCnt",
                                      "globalAvgShift", "globalBias"}
    m.globalAvgShift = 0.0
    U, V, ci2 = mgr.loadCalcResults()
    assert (U == m.userFactors).all() and (V == m.itemFactors).all() and m.globalAvgShift == 0.25
    # more users than when the results were stored: old rows kept, new rows from the random init
    m.totalUsersCount = 7
    U2, _, _ = mgr.loadCalcResults()
    assert U2.shape == (7, 4) and (U2[:5] == m.userFactors).all() and np.abs(U2[5:]).max() > 0
    # different rank: nothing to reuse
    m.factorsCount = 8
    assert mgr.loadCalcResults() is None


def test_launch_plan_classes_cpu():
    """The library's host-side launch planning (csrc/ycnr_als.cu plan_count / plan_fill), no GPU: rows with
    n <= 0 are skipped (Q2 degenerate rows), n <= 88 (k = 100) go to the dual bin of their tile-row count ceil(n/4), longer
    rows are cut into <= 4096-rating slices for the tensor-core Gram, and the slices are handed out longest first."""
    import numpy as np
    from you_can_not_recommend_b200 import native
    lens = np.asarray([0, 1, 4, 5, 88, 89, 4096, 4097, 10000, -1, 33, 300], np.int32)
    p = native.debug_plan(lens, factors_count=100)
    assert [list(b) for b in p["dual"] if len(b)] == [[1, 2], [3], [10], [4]]         # mt = 1, 2, 9, 22
    assert list(p["dual"][0]) == [1, 2] and list(p["dual"][1]) == [3] and list(p["dual"][8]) == [10] and list(p["dual"][21]) == [4]
    assert list(p["fused"]) == [] and list(p["multi"]) == [5, 6, 7, 8, 11]
    items = list(zip(p["item_row"].tolist(), p["item_off"].tolist()))
    assert items == [(5, 0), (6, 0), (7, 0), (7, 4096), (8, 0), (8, 4096), (8, 8192), (11, 0)]
    slice_len = [min(4096, int(lens[r]) - o) for r, o in items]
    order = p["item_order"].tolist()
    assert sorted(order) == list(range(len(items)))
    assert [slice_len[i] // 32 for i in order] == sorted((l // 32 for l in slice_len), reverse=True)
    # the measured dual / primal crossover: 88 ratings on the tensor-core path at k = 100, 96 for wide systems
    assert len(native.debug_plan(np.asarray([96], np.int32), factors_count=100)["multi"]) == 1
    assert len(native.debug_plan(np.asarray([96], np.int32), factors_count=256)["dual"][23]) == 1
    # FFMA path (k % 4 != 0): rows up to split_cols are fused rows, longest first; k = 7 -> dual rows up to 4 ratings
    q = native.debug_plan(lens, factors_count=7)
    assert list(q["dual"][0]) == [1, 2] and sum(len(b) for b in q["dual"]) == 2
    assert list(q["fused"]) == [6, 11, 5, 4, 10, 3] and list(q["multi"]) == [7, 8]


def test_init_first_factor_as_avg_rating_cpu():
    """als.initFirstFactorAsAvgRating (EmfBase.js:500-511): the first factor of every row that has ratings is its
    average rating over sets 1,2,3 (EmfLord.getStats, EmfLord.js:67-119); rows without ratings keep the random
    draw (`avg !== undefined`, not `avg > 0`)."""
    import numpy as np
    from you_can_not_recommend_b200 import front_end as fe
    from you_can_not_recommend_b200.emf_master import EmfMaster
    # user 2 and item 3 have no ratings at all
    t = fe.table_from_triples(4, 5, [0, 0, 1, 3, 3, 3], [0, 1, 1, 0, 2, 4], [5, 3, 4, 1, 2, 3])
    m = EmfMaster(t, {"factorsCount": 4, "als": {"initFirstFactorAsAvgRating": True}})
    m.splitDataForTrain()
    m.createSharedFactors()
    m.initSharedFactorsRandom()
    plain = EmfMaster(t, {"factorsCount": 4})
    plain.splitDataForTrain()
    plain.createSharedFactors()
    plain.initSharedFactorsRandom()
    assert np.allclose(m.userFactors[[0, 1, 3], 0], [4.0, 4.0, 2.0])
    assert np.allclose(m.itemFactors[[0, 1, 2, 4], 0], [3.0, 3.5, 2.0, 3.0])
    assert m.userFactors[2, 0] == plain.userFactors[2, 0] and m.itemFactors[3, 0] == plain.itemFactors[3, 0]
    assert (m.userFactors[:, 1:] == plain.userFactors[:, 1:]).all()
    assert "ratingsAvgPerUser" in m.stats and plain.stats["ratingsAvgPerUser"] is None


def test_manager_calc_cnt_once_per_run_cpu(tmp_path):
    """calcCnt goes up once per train run (EmfLord.js:904) however many per-iteration checkpoints are written,
    the last checkpoint is not written twice, and only rank 0 touches the directories."""
    import json
    import numpy as np
    from you_can_not_recommend_b200.emf_base import default_options
    from you_can_not_recommend_b200.emf_manager import EmfManager

    class M:
        def __init__(self, rank):
            self.options = default_options()
            self.options["checkpointEveryIter"] = True
            self.ctx, self.rank, self.history, self.saves = None, rank, [], 0
            self.factorsCount, self.totalUsersCount, self.totalItemsCount, self.globalAvgShift = 2, 3, 2, 0.0
            self.userFactors = np.ones((3, 2), np.float32)
            self.itemFactors = np.ones((2, 2), np.float32)

        def trainIter(self):
            self.userFactors += 1
            self.history.append({})

    m = M(0)
    mgr = EmfManager(m, str(tmp_path / "a"))
    orig = mgr.saveCalcResults
    count = []
    mgr.saveCalcResults = lambda *a: (count.append(1), orig(*a))[1]
    mgr.train(3)
    assert len(count) == 3                                                    # 3 checkpoints, no extra final save
    ci = json.load(open(tmp_path / "a_factors_ready" / "calc_info.json"))
    assert ci["calcCnt"] == 1
    assert np.fromfile(tmp_path / "a_factors_ready" / "user_factors", np.float32)[0] == 4.0
    mgr.train(1)
    assert json.load(open(tmp_path / "a_factors_ready" / "calc_info.json"))["calcCnt"] == 2
    m1 = M(1)
    EmfManager(m1, str(tmp_path / "b")).train(2)
    assert not (tmp_path / "b_factors_ready").exists() and not (tmp_path / "b_factors_temp").exists()


def test_check_portion_header_against_array_lengths_cpu():
    """ycnr_check_portion (host only): the header must fit the arrays that carry it — what the N-API binding
    checks before it hands typed arrays to the portion calls (EmfWorker.js:176-219 reads them unchecked)."""
    import ctypes as C
    import numpy as np
    from you_can_not_recommend_b200 import native
    L = native.lib()
    rows = np.asarray([2, 5, 3, 9, 4], np.int32)

    def chk(r, ni, nv):
        return L.ycnr_check_portion(r.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(len(r)), C.c_int64(ni), C.c_int64(nv))
    assert chk(rows, 7, 7) == 0 and chk(rows, 8, 9) == 0
    assert chk(rows, 6, 7) != 0 and b"do not fit the indx/vals" in L.ycnr_last_error()
    assert chk(rows[:4], 7, 7) != 0 and b"do not fit the rows array" in L.ycnr_last_error()
    assert chk(np.asarray([-1], np.int32), 0, 0) != 0
    assert chk(np.asarray([1, 0, -2], np.int32), 5, 5) != 0 and b"negative cols" in L.ycnr_last_error()
    assert chk(np.asarray([0], np.int32), 0, 0) == 0
