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
