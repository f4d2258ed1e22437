"""CPU oracle for the ALS hot path (TEST INFRASTRUCTURE ONLY — see als_oracle.c).

PARITY UNPINNED: upstream has no tests/fixtures and its BLAS/LAPACK layer is an
un-vendored dependency; this restates the reference's call sequence
(lib/emf/EmfWorker.js:169-315, lib/emf/EmfMaster.js:389-412,757-786,
lib/emf/EmfLord.js:892-902,954-958,1043-1081).

  O32 = float arithmetic, O64 = double arithmetic (the neutral referee).
  use_blas=True routes through the dlopen'ed OpenBLAS exactly as the reference's
  nblas calls would (sgemm T/N, sgemv, sgesv, sdot); otherwise portable C loops.
"""
import ctypes as C
import glob
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("als_oracle.c", "als_oracle_body.inc", "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def find_openblas():
    """The OpenBLAS bundled with scipy in this image."""
    try:
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        if cands:
            return os.path.abspath(cands[0])
    except Exception:
        pass
    return None


def set_blas(threads=0, path=None):
    """Enable the BLAS/LAPACK routine path. Returns True when the library was loaded."""
    path = path or find_openblas()
    if not path:
        return False
    return lib().oracle_set_blas(path.encode(), C.c_int(int(threads))) == 0


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def als_portion(buf_rows, buf_indx, buf_vals, fixed, solved, lam, use_blas=False):
    """mw_calcTrainAlsPortion on one portion; `solved` is updated in place. dtype picks O32/O64."""
    assert fixed.dtype == solved.dtype and fixed.flags.c_contiguous and solved.flags.c_contiguous
    k = fixed.shape[1]
    nr = C.c_int64(0)
    if fixed.dtype == np.float32:
        fn, ct = lib().oracle_als_portion_f32, C.c_float
    else:
        fn, ct = lib().oracle_als_portion_f64, C.c_double
    rc = fn(_p(buf_rows, C.c_int32), _p(buf_indx, C.c_int32), _p(buf_vals, C.c_float),
            _p(fixed, ct), _p(solved, ct), C.c_int(k), C.c_double(lam), C.c_int(int(use_blas)), C.byref(nr))
    if rc != 0:
        raise FloatingPointError("oracle: singular system in portion row %d" % (rc - 100))
    return nr.value


def rmse_portion(buf_rows, buf_indx, buf_vals, U, V, shift, use_blas=False):
    """mw_calcRmsePortion on one portion -> (rSumDiff2, rCnt, rSum)."""
    k = U.shape[1]
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    if U.dtype == np.float32:
        fn, ct = lib().oracle_rmse_portion_f32, C.c_float
    else:
        fn, ct = lib().oracle_rmse_portion_f64, C.c_double
    fn(_p(buf_rows, C.c_int32), _p(buf_indx, C.c_int32), _p(buf_vals, C.c_float), _p(U, ct), _p(V, ct),
       C.c_int(k), C.c_double(shift), C.c_int(int(use_blas)), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


class OracleTrainer:
    """The reference train loop on the CPU.

    portions: dict stepType -> list of (bufRows, bufIndx, bufVals) in upstream wire format,
              stepType in byUser, byItem, rmseValidate, rmseTest.
    """

    def __init__(self, U0, V0, portions, user_reg=0.05, item_reg=0.05, total_ratings_avg=0.0,
                 dtype=np.float32, use_blas=False):
        self.U = np.ascontiguousarray(U0, dtype=dtype).copy()
        self.V = np.ascontiguousarray(V0, dtype=dtype).copy()
        self.portions = portions
        self.user_reg, self.item_reg = user_reg, item_reg
        self.total_ratings_avg = total_ratings_avg
        self.use_blas = use_blas
        self.global_avg_shift = 0.0
        self.history = []

    def als_train_step(self, step_type):            # EmfLord.alsTrainStep 963-984
        by_user = step_type == "byUser"
        fixed, solved = (self.V, self.U) if by_user else (self.U, self.V)
        lam = self.user_reg if by_user else self.item_reg
        n = 0
        for rows, indx, vals in self.portions[step_type]:
            n += als_portion(rows, indx, vals, fixed, solved, lam, self.use_blas)
        return n

    def calc_rmse(self, step_type, use_global_avg_shift):   # EmfLord.calcRmse 1043-1081
        calc_shift = not use_global_avg_shift               # EmfMaster._startCalcRmse 389-402
        if calc_shift:
            self.global_avg_shift = 0.0
        r_sum_diff2 = r_cnt = r_sum = 0.0
        last = (0.0, 0.0, 0.0)
        for rows, indx, vals in self.portions[step_type]:
            last = rmse_portion(rows, indx, vals, self.U, self.V, self.global_avg_shift, self.use_blas)
            r_sum_diff2 += last[0]
            r_cnt += last[1]
            r_sum += last[2]
        rmse = math.sqrt(1.0 * r_sum_diff2 / r_cnt) if r_cnt else float("nan")
        # Q7: predAvg from the LAST completed portion only (EmfMaster.js:779 uses msg., not this.)
        pred_avg = last[2] / last[1] if last[1] else float("nan")
        if calc_shift:
            self.global_avg_shift = self.total_ratings_avg - pred_avg
        return rmse

    def train_iter(self):                            # EmfLord.train 892-902
        self.als_train_step("byUser")
        self.als_train_step("byItem")
        out = {}
        out["rmseValidate"] = self.calc_rmse("rmseValidate", False)
        out["rmseTest"] = self.calc_rmse("rmseTest", False)
        out["rmseTestShift"] = self.calc_rmse("rmseTest", True)
        out["globalAvgShift"] = self.global_avg_shift
        self.history.append(out)
        return out

    def train(self, iters):
        for _ in range(iters):
            self.train_iter()
        return self.history


def recommend_items_for_user(U, V, user_id0, skip_item_ids0, limit=20, min_recommend_rating=0.0,
                             global_avg_shift=0.0):
    """Literal restatement of YcnrController.recommendItemsForUser (lib/YcnrController.js:227-284), 0-based ids.

    The loop is upstream's: push, sort by predict descending (a stable sort, as in current V8), raise
    minRatingInSelection, and pop the last entry whenever the list has reached `limit` (281-282) — which is
    why at most limit-1 items come back.  predict = fp32 dot + shift in double (EmfBase.js:815-827).
    Returns [(item_id0, predict)], best first."""
    skip = set(int(i) for i in skip_item_ids0)                 # skipItemIds, 244-251
    uf = np.asarray(U[user_id0], np.float32)
    rec = []
    min_in_sel = 0.0
    for item in range(V.shape[0]):                             # 264
        if item in skip:
            continue
        predict = float(np.dot(uf, np.asarray(V[item], np.float32))) + global_avg_shift
        if predict >= min_recommend_rating and (len(rec) < limit or predict > min_in_sel):   # 271-272
            rec.append((item, predict))
            rec.sort(key=lambda t: -t[1])                      # 274 (list.sort is stable)
            if predict > min_in_sel:
                min_in_sel = predict                           # 275-276
            if len(rec) >= limit:
                rec.pop()                                      # 277-278
    return rec
