"""EmfBase — options, factor store and row helpers of the reference, host side.

Mirror of lib/emf/EmfBase.js for the ALS hot path only: same option names and
defaults (EmfBase.js:52-140), same factor layout (dense row-major Float32
[total x factorsCount], EmfBase.js:399-450), same method names for the pieces the
worker uses.  The arithmetic behind them is the CUDA library (native.py); options the
GPU path cannot honour (useDoublePrecision, lowmem) are rejected, not emulated.
"""
import copy
import os

import numpy as np

from . import front_end, native


def default_options():
    """EmfBase.DefaultOptions (EmfBase.js:52-140), ALS-relevant subset + the GPU knobs."""
    return {
        "dbType": "ml",
        "maxRating": {"mal": 10, "ml": 5},
        "als": {"userFactReg": 0.05, "itemFactReg": 0.05, "initFirstFactorAsAvgRating": False},
        "factorsCount": 100,
        "trainIters": 10,
        "alg": "als",
        "dataSetDistr": [85, 10, 5],
        "ratingsInPortionForRmse": 10 * 1000,
        "ratingsInPortionForAls": {"byUser": 10 * 1000, "byItem": 10 * 1000},
        # README.md:17-18: one worker, the math library is parallel itself
        "numThreadsForTrain": {"als": 1, "sgd": 1},
        "numThreadsForRmse": 1,
        "useDoublePrecision": False,
        "usePortionsCache": True,
        "lowmem": False,
        "keepFactorsOpened": True,
        "useClustering": True,
        "shared": {"userFactorsShmKey": -1, "itemFactorsShmKey": -1, "portionBufferShmKeys": {}},
        # --- additions of the B200 path (not in the reference) ---
        "gpu": {"device": 0, "gramPath": "auto", "dualMaxCols": -1, "splitCols": 0, "profile": False,
                "bulk": False, "cachePortions": False, "tcMinCols": 0, "solveChunks": 0, "deviceIngest": False},
        "seed": front_end.DEFAULT_SEED,
    }


def deepmerge(a, b):
    """deepmerge.all of the reference (EmfBase.js:284-287): dicts merge, everything else replaces."""
    out = copy.deepcopy(a)
    for k, v in (b or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = deepmerge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


_GRAM = {"auto": native.GRAM_AUTO, "ffma": native.GRAM_FFMA, "tc": native.GRAM_TC3XTF32}


class EmfBase:
    def __init__(self, options=None):
        self.options = deepmerge(default_options(), options or {})
        if self.options["alg"] != "als":
            raise ValueError("only alg='als' is on the B200 path (SGD is deprecated upstream, README.md:13)")
        self.factorsCount = int(self.options["factorsCount"])
        self.TypedArrayClass = np.float32
        self.TypedArraySize1 = 4
        self.totalUsersCount = 0
        self.totalItemsCount = 0
        self.userFactors = None
        self.itemFactors = None
        self.globalAvgShift = 0.0
        self.stats = {}
        self.ctx = None

    # -- factor store ---------------------------------------------------------------
    def createSharedFactors(self):
        """EmfBase.createSharedFactors (EmfBase.js:399-425); numpy arrays stand in for shm segments."""
        k = self.factorsCount
        self.userFactors = np.zeros((self.totalUsersCount, k), np.float32)
        self.itemFactors = np.zeros((self.totalItemsCount, k), np.float32)

    def openSharedFactors(self, userFactors, itemFactors):
        """EmfBase.openSharedFactors (EmfBase.js:430-450): adopt existing segments (no copy)."""
        k = self.factorsCount
        assert userFactors.dtype == np.float32 and userFactors.shape == (self.totalUsersCount, k)
        assert itemFactors.dtype == np.float32 and itemFactors.shape == (self.totalItemsCount, k)
        self.userFactors, self.itemFactors = userFactors, itemFactors

    def initSharedFactorsRandom(self, seed=None):
        """EmfBase.initSharedFactorsRandom (EmfBase.js:457-513): randomNormal(1/k) per element;
        optional first factor = average rating of the row (als.initFirstFactorAsAvgRating)."""
        seed = self.options["seed"] + 2 if seed is None else seed
        k = self.factorsCount
        self.userFactors[...] = front_end.init_factors(self.totalUsersCount, k, 0, seed)
        self.itemFactors[...] = front_end.init_factors(self.totalItemsCount, k, 1, seed)
        if self.options["als"]["initFirstFactorAsAvgRating"]:       # EmfBase.js:500-511
            for mat, key in ((self.userFactors, "ratingsAvgPerUser"), (self.itemFactors, "ratingsAvgPerItem")):
                avg = self.stats.get(key)
                if avg is None:
                    raise ValueError("als.initFirstFactorAsAvgRating needs stats.%s (run splitDataForTrain first)" % key)
                has = ~np.isnan(np.asarray(avg, np.float64))       # `avg !== undefined`
                mat[has, 0] = np.asarray(avg, np.float32)[has]

    def openDevice(self):
        """Create the GPU context for this worker and mirror the factor segments on it."""
        o = self.options
        if o["useDoublePrecision"]:
            raise ValueError("useDoublePrecision=true is not supported by the B200 path (float32 only)")
        if o["lowmem"]:
            raise ValueError("lowmem=true (file-backed factors) is not supported by the B200 path")
        g = o["gpu"]
        dev = g["device"]
        if "LOCAL_RANK" in os.environ and g.get("deviceFromLocalRank", True):
            dev = int(os.environ["LOCAL_RANK"])
        self.ctx = native.Context(
            self.factorsCount, self.totalUsersCount, self.totalItemsCount,
            o["als"]["userFactReg"], o["als"]["itemFactReg"], False, False, dev,
            _GRAM[g["gramPath"]], g["dualMaxCols"], g["splitCols"], g["profile"], g.get("tcMinCols", 0),
            solve_chunks=g.get("solveChunks", 0))
        self.ctx.attach_factors(self.userFactors, self.itemFactors)
        return self.ctx

    def closeDevice(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None

    # -- row helpers used by the worker ------------------------------------------------
    def getLatentFactorsPartData(self, type_, factorsBuffer, firstRowId, rowId):
        """View of the output factor row (EmfBase.js:518-532), no copy."""
        latent = self.userFactors if type_ == "byUser" else self.itemFactors
        return latent[rowId]

    def copySubFixedFactors(self, type_, subFixedFactorsData, indx):
        """EmfBase.copySubFixedFactors (EmfBase.js:537-555) through the addon's gather
        (the slot of cpp_utils sAlsBuildSubFixedFacts)."""
        fixed = self.itemFactors if type_ == "byUser" else self.userFactors
        indx = np.ascontiguousarray(indx, np.int32)
        sub = subFixedFactorsData.reshape(-1)[: len(indx) * self.factorsCount].reshape(len(indx), self.factorsCount)
        self.ctx.build_sub_fixed_facts(sub, fixed, indx)

    def getFactorsRowSync(self, type_, rowId):
        """EmfBase.getFactorsRowSync (EmfBase.js:702-718), shm branch."""
        return (self.userFactors if type_ == "byUser" else self.itemFactors)[rowId]

    def _alsPredict(self, uF, iF):
        """EmfBase._alsPredict (EmfBase.js:825-827): fp32 dot, then + globalAvgShift in double."""
        return float(np.dot(uF, iF)) + self.globalAvgShift

    def alsPredictSync(self, userId, itemId, uF=None):
        if uF is None:
            uF = self.getFactorsRowSync("byUser", userId)
        return self._alsPredict(uF, self.getFactorsRowSync("byItem", itemId))

    predictSync = alsPredictSync
