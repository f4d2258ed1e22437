"""EmfManager — persistence of calc results, host side.

Mirror of the parts of lib/emf/EmfManager.js that survive on the B200 path:

  getCalcInfo ............ EmfManager.js:158-176   (same fields, same JSON file name)
  _canReuseCalcResults ... EmfManager.js:179-191
  saveCalcResults ........ EmfManager.js:463-495 + _saveCalcResultsToRecommender 501-570: the factor matrices go
                           to <dir>_factors_temp as raw row-major Float32 files `user_factors` / `item_factors`
                           (EmfBase.js:289-299, byte length EmfBase.js:694-697) next to `calc_info.json`, then
                           the directory is renamed to <dir>_factors_ready — the rename is the commit point
  loadCalcResults ........ EmfManager.js:324-397: read them back when alg / dbType / factorsCount /
                           precision match; a grown table keeps the stored rows and leaves the new rows to the
                           random init (EmfManager.js:432-443, EmfMaster.js:347-358)
  checkpoint ............. the author's todo "saveCalcResults every iter!" (YcnrController.js:288): the same
                           commit after every iteration when options.checkpointEveryIter is set

The bytes come from the device replicas: a bulk-mode master first copies both matrices into the host
segments (EmfMaster.syncFactorsToHost), per-portion mode already has them there after every half-step.
"""
import json
import os
import shutil
import time

import numpy as np

USER_FACTORS_FILENAME = "user_factors"      # EmfBase.js:289-293
ITEM_FACTORS_FILENAME = "item_factors"
CALC_INFO_FILENAME = "calc_info.json"


class EmfManager:
    def __init__(self, master, factorsDir):
        """master: an EmfMaster (options, factor segments, counts); factorsDir: the `data/<db>` prefix —
        results live in <factorsDir>_factors_temp / <factorsDir>_factors_ready (EmfBase.js factorsPath)."""
        self.m = master
        self.factorsTempPath = factorsDir + "_factors_temp"
        self.factorsReadyPath = factorsDir + "_factors_ready"
        self.calcDate = None
        self.calcCnt = 0
        self.lastCalcInfo = None

    # -- calc info ------------------------------------------------------------------------------
    def getCalcInfo(self):
        m, o = self.m, self.m.options
        return {
            "alg": o["alg"], "algOptions": o[o["alg"]], "useDoublePrecision": o["useDoublePrecision"],
            "factorsCount": m.factorsCount, "dataSetDistr": o["dataSetDistr"],
            "totalUsersCount": m.totalUsersCount, "totalItemsCount": m.totalItemsCount, "dbType": o["dbType"],
            "calcDate": self.calcDate, "calcCnt": self.calcCnt,
            "globalAvgShift": m.globalAvgShift, "globalBias": 0,
        }

    def _canReuseCalcResults(self, ci1):
        ci2 = self.getCalcInfo()
        return (ci1 is not None and ci1["alg"] == ci2["alg"] and ci1["dbType"] == ci2["dbType"]
                and ci1["factorsCount"] == ci2["factorsCount"]
                and ci1["useDoublePrecision"] == ci2["useDoublePrecision"])

    # -- save -----------------------------------------------------------------------------------
    def saveCalcResults(self, calcInfo=None):
        """Only rank 0 touches the directories (every rank holds the same replicas after a half-step): with
        several ranks the rmtree/rename below would race.  calcCnt is the caller's business — upstream bumps it
        once per train run (EmfLord.js:904), not per save."""
        m = self.m
        if m.options["gpu"]["bulk"] and m.ctx is not None:
            m.syncFactorsToHost()
        self.calcDate = time.strftime("%Y-%m-%dT%H:%M:%S")
        info = calcInfo or self.getCalcInfo()
        if getattr(m, "rank", 0) != 0:
            self.lastCalcInfo = info
            return info
        if os.path.isdir(self.factorsTempPath):
            shutil.rmtree(self.factorsTempPath)
        os.makedirs(self.factorsTempPath)
        for name, mat in ((USER_FACTORS_FILENAME, m.userFactors), (ITEM_FACTORS_FILENAME, m.itemFactors)):
            with open(os.path.join(self.factorsTempPath, name), "wb") as f:
                f.write(np.ascontiguousarray(mat, np.float32).tobytes())     # raw row-major, no header
                f.flush()
                os.fsync(f.fileno())
        with open(os.path.join(self.factorsTempPath, CALC_INFO_FILENAME), "w") as f:
            json.dump(info, f, indent=2)
        # critical section: move /factors_temp to /factors_ready (EmfManager.js:548-553)
        if os.path.isdir(self.factorsReadyPath):
            shutil.rmtree(self.factorsReadyPath)
        os.rename(self.factorsTempPath, self.factorsReadyPath)
        self.lastCalcInfo = info
        return info

    # -- load -----------------------------------------------------------------------------------
    def loadCalcResults(self):
        """Returns (userFactors, itemFactors, calcInfo) ready for EmfMaster.prepareToTrain(U, V), or None when
        nothing reusable is stored.  Matrices are sized for the CURRENT counts: stored rows are kept, rows of
        users/items that did not exist then are drawn from the random init."""
        path = os.path.join(self.factorsReadyPath, CALC_INFO_FILENAME)
        if not os.path.exists(path):
            return None
        with open(path) as f:
            ci = json.load(f)
        if not self._canReuseCalcResults(ci):
            return None
        m, k = self.m, self.m.factorsCount
        out = []
        for name, rows_old, rows_new, which in ((USER_FACTORS_FILENAME, ci["totalUsersCount"], m.totalUsersCount, 0),
                                                (ITEM_FACTORS_FILENAME, ci["totalItemsCount"], m.totalItemsCount, 1)):
            raw = np.fromfile(os.path.join(self.factorsReadyPath, name), np.float32)
            if raw.size != rows_old * k:
                raise ValueError("%s: %d floats on disk, calc_info says %d x %d" % (name, raw.size, rows_old, k))
            mat = raw.reshape(rows_old, k)
            if rows_new != rows_old:
                from . import front_end
                grown = front_end.init_factors(rows_new, k, which, m.options["seed"] + 2)
                keep = min(rows_old, rows_new)
                grown[:keep] = mat[:keep]
                mat = grown
            out.append(np.ascontiguousarray(mat))
        self.lastCalcInfo = ci
        self.calcCnt = int(ci.get("calcCnt") or 0)
        m.globalAvgShift = float(ci.get("globalAvgShift") or 0.0)
        return out[0], out[1], ci

    # -- train with checkpoints --------------------------------------------------------------------
    def train(self, iters=None):
        """EmfLord.train (EmfLord.js:864-926) + saveCalcResults at the end (921); with
        options.checkpointEveryIter also after every iteration."""
        m = self.m
        n = m.options["trainIters"] if iters is None else iters
        self.calcCnt += 1                                   # once per run (EmfLord.js:904)
        saved = False
        for _ in range(n):
            m.trainIter()
            saved = bool(m.options.get("checkpointEveryIter"))
            if saved:
                self.saveCalcResults()
        if not saved:                                       # the last iteration's checkpoint IS the final state
            self.saveCalcResults()
        return m.history
