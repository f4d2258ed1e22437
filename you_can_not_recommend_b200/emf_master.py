"""EmfMaster — single-box driver of the worker interface (the caller of the hot path).

This is the minimum of lib/emf/EmfMaster.js + lib/emf/EmfLord.js needed to drive the
workers exactly as the reference does, with the PostgreSQL side replaced by an
in-memory RatingsTable (front_end.py):

  prepareToTrain ........ EmfLord.js:617-653, 39-43 (split to sets, stats, portions),
                          EmfMaster.js:138-142,156-234 (worker + portion buffers)
  alsTrainStep .......... EmfLord.js:963-984, EmfMaster.js:364-383,645-689
  calcRmse .............. EmfLord.js:1043-1081, EmfMaster.js:389-412,757-786 (Q7)
  train ................. EmfLord.js:892-902 (Q8 order)

Two ways to run a step: the drop-in per-portion messages (reference wire format), or
`bulk` row sets that keep every portion of a step resident on the GPU.
With world > 1 (one process per GPU, torch.distributed) portions are split into
contiguous nnz-balanced slices per rank and the solved slices are exchanged after every
half-step (dist.py), replacing 'alsSaveCalcedFactors' (EmfMaster.js:711-723).
"""
import math
import time

import numpy as np

from . import dist as ydist
from . import front_end as fe
from . import native
from .emf_base import EmfBase
from .emf_worker import EmfProcess, EmfWorker

STEP_MASK = {"byUser": fe.MASK_TRAIN, "byItem": fe.MASK_TRAIN,
             "rmseValidate": fe.MASK_VALIDATE, "rmseTest": fe.MASK_TEST}


class EmfMaster(EmfBase):
    def __init__(self, table, options=None, rank=0, world=1, group=None):
        super().__init__(options)
        self.table = table
        self.rank, self.world, self.group = rank, world, group
        self.portionsRowIdTo = {}
        self.portionsCount = {}
        self.maxRatingsInPortion = {}
        self.maxRowsInPortion = {}
        self.rowlists = {}
        self.rowsets = {}
        self.my_portions = {}
        self.workers = []
        self.rmse = float("nan")
        self.predAvg = float("nan")
        self.history = []
        self.completedPortions = 0
        self.rSumDiff2 = self.rCnt = self.rSum = 0.0
        self._lastRmseMsg = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.fusedPeers = False
        self.sharedHost = False
        self._ranges = {}
        self.phase_ms = {}
        self._workerMU = []
        self._prepared = {}
        self._rmseGlobal = {}
        self._alsSteps = 0

    # ---- prepare ---------------------------------------------------------------------------
    def splitDataForTrain(self):
        """EmfLord.splitDataForTrain (EmfLord.js:39-43): splitToSets -> getStats -> splitToPortions."""
        t, o = self.table, self.options
        if not (t.dataset_type != 0).any():
            fe.split_sets(t, o["dataSetDistr"], seed=o["seed"] + 1)
        self.totalUsersCount, self.totalItemsCount = t.users, t.items   # max(id), Q5
        cu, ci = t.counts_per_user(), t.counts_per_item()
        # avg_rating of EmfLord.getStats (EmfLord.js:67-78, 95-119); only als.initFirstFactorAsAvgRating reads it
        au, ai = (t.avg_per_user(), t.avg_per_item()) if o["als"]["initFirstFactorAsAvgRating"] else (None, None)
        self.stats = {
            "ratingsCntPerUser": cu, "ratingsCntPerItem": ci,
            "ratingsAvgPerUser": au, "ratingsAvgPerItem": ai,   # NaN = upstream's `undefined` (row.cnt == 0)
            "maxRatingsPerUser": int(cu.max()), "maxRatingsPerItem": int(ci.max()),
            "trainUsersRatingsCount": int(cu.sum(dtype=np.int64)),
            "trainItemsRatingsCount": int(ci.sum(dtype=np.int64)),
            "totalRatingsAvg": t.total_ratings_avg(),
        }
        nthr = o["numThreadsForTrain"]["als"]
        d = o["dataSetDistr"]
        for step, cnt, rip, pct in (
                ("byUser", cu, o["ratingsInPortionForAls"]["byUser"], 0),
                ("byItem", ci, o["ratingsInPortionForAls"]["byItem"], 0),
                ("rmseValidate", cu, o["ratingsInPortionForRmse"], d[1] + 1),
                ("rmseTest", cu, o["ratingsInPortionForRmse"], d[2] + 1)):
            pto, mr, mrows = fe.split_to_portions(cnt, rip, nthr, pct)
            self.portionsRowIdTo[step] = pto
            self.portionsCount[step] = len(pto)
            self.maxRatingsInPortion[step] = mr
            self.maxRowsInPortion[step] = mrows

    def _csr(self, step):
        if step == "byItem":
            return self.table.csr_by_item(STEP_MASK[step])
        return self.table.csr_by_user(STEP_MASK[step])

    def _slice_portions(self, step):
        """Contiguous, nnz-balanced slice of the portion list for this rank."""
        pto = self.portionsRowIdTo[step]
        n = len(pto)
        if self.world == 1 or n == 0:
            return 0, n
        ptr = self._csr(step).ptr
        if step in ("byUser", "byItem"):                 # estimated solve cost per row, not raw nnz
            ends = ydist.cost_ends(np.diff(ptr), pto, self.factorsCount)
        else:
            ends = ptr[np.asarray(pto, np.int64)]        # cumulative ratings at each portion end
        cuts = ydist.balanced_cuts(ends, self.world)
        return int(cuts[self.rank]), int(cuts[self.rank + 1])

    def prepareToTrain(self, userFactors=None, itemFactors=None):
        o = self.options
        self.splitDataForTrain()
        if userFactors is not None:
            self.openSharedFactors(userFactors, itemFactors)      # warm start (EmfManager.js:405-457)
        elif self.world > 1 and not o["gpu"]["bulk"]:
            # upstream, every worker of a node maps the SAME SysV segments (EmfBase.js:403-412, 430-450) and
            # writes its own rows into them: one /dev/shm mapping shared by all ranks of the box
            self._createNodeSharedFactors()
        else:
            self.createSharedFactors()
            self.initSharedFactorsRandom()
        device_front_end = o["gpu"]["bulk"] and o["gpu"].get("deviceIngest", False)
        if not device_front_end:                 # (the device front end cuts the slices from its own counts)
            for step in ("byUser", "byItem", "rmseValidate", "rmseTest"):
                self.my_portions[step] = self._slice_portions(step)
        # worker + portion buffers (EmfMaster.js:156-234: Int32[2*maxRows+1], Int32/Float32[maxRatings])
        mra = max(self.maxRatingsInPortion["byUser"], self.maxRatingsInPortion["byItem"])
        mrow = max(self.maxRowsInPortion["byUser"], self.maxRowsInPortion["byItem"])
        mrr = max(self.maxRatingsInPortion["rmseValidate"], self.maxRatingsInPortion["rmseTest"])
        mrowr = max(self.maxRowsInPortion["rmseValidate"], self.maxRowsInPortion["rmseTest"])
        pb = {
            "alsRows": np.zeros(2 * mrow + 1, np.int32), "alsIndx": np.zeros(mra, np.int32),
            "alsVals": np.zeros(mra, np.float32),
            "rmseRows": np.zeros(2 * mrowr + 1, np.int32), "rmseIndx": np.zeros(mrr, np.int32),
            "rmseVals": np.zeros(mrr, np.float32),
        }
        wp, mp = EmfProcess(), EmfProcess()
        wp.peer, mp.peer = mp, wp
        mp.on("completedPortion", self.wm_completedPortion)
        mp.on("preparedToTrain", lambda d: None)
        mp.on("setMemoryUsage", lambda d: self._workerMU.append(d["mu"]))
        mp.on("endedTrainStep", lambda d: None)
        w = EmfWorker(0, wp, o)
        w.master_side = mp
        self.workers = [w]
        mp.emit("prepareToTrain", {
            "stats": {"totalRatingsAvg": self.stats["totalRatingsAvg"]},
            "options": o, "totalUsersCount": self.totalUsersCount, "totalItemsCount": self.totalItemsCount,
            "shared": {"userFactors": self.userFactors, "itemFactors": self.itemFactors, "portionBuffer": pb},
        })
        self.ctx = w.ctx
        mp.emit("startTrain")
        if o["gpu"]["bulk"]:
            self.prepareBulk()
        elif o["usePortionsCache"] and o["gpu"].get("cachePortions", False):
            self.preparePortionsCache()
        return self

    def preparePortionsCache(self):
        """usePortionsCache taken to its limit: every portion of every step stays converted in
        page-locked host memory, so a step only hands buffers to the worker (no fetch/convert per
        iteration; upstream keeps one portion ahead, EmfMaster.js:434-494,656-658).  A portion's
        indx/vals are a contiguous slice of the step's fetch, so the cache is the fetch itself plus
        one header array per portion."""
        self.portionCache = {}
        for step in ("byUser", "byItem", "rmseValidate", "rmseTest"):
            csr = self._csr(step)
            pto = np.asarray(self.portionsRowIdTo[step], np.int32)
            rl = fe.build_rowlist(csr, pto)
            if csr.nnz:
                self.ctx.host_register(csr.idx)
                self.ctx.host_register(csr.vals)
            prefix = "rmse" if step.startswith("rmse") else "als"
            bufs = []
            for p in range(len(pto)):
                r0, r1 = int(rl.portion_first[p]), int(rl.portion_first[p + 1])
                hdr = np.empty(2 * (r1 - r0) + 1, np.int32)
                hdr[0] = r1 - r0
                hdr[1::2] = rl.row_ids[r0:r1]
                hdr[2::2] = rl.row_len[r0:r1]
                if hdr.nbytes >= (1 << 16):          # big headers are DMA'd straight from the cache too
                    self.ctx.host_register(hdr)
                a = int(csr.ptr[0 if p == 0 else pto[p - 1]])
                b = max(int(csr.ptr[pto[p]]), a + 1) if csr.nnz else a
                bufs.append({prefix + "Rows": hdr, prefix + "Indx": csr.idx[a:b], prefix + "Vals": csr.vals[a:b],
                             "fetched": int(csr.ptr[pto[p]]) - a})
            self.portionCache[step] = bufs

    def _preparedPortions(self, step, prefix, lo, hi):
        key = (step, lo, hi)
        if key not in self._prepared:
            bufs = self.portionCache[step][lo:hi]
            prep = self.ctx.prepare_portions([(b[prefix + "Rows"], b[prefix + "Indx"], b[prefix + "Vals"]) for b in bufs],
                                             tags=list(range(lo, hi)))
            prep["h2d"] = sum(8 * b["fetched"] + 4 * len(b[prefix + "Rows"]) for b in bufs)
            self._prepared[key] = prep
        return self._prepared[key]

    def prepareBulkOnDevice(self):
        """gpu.deviceIngest: upload the ratings table once and let the device build every step's fetch and
        portion headers (ycnr_table_upload / ycnr_rowset_from_table) — the master's SQL fetch + per-rating
        conversion loop (EmfMaster.js:501-614) for whole steps, bit-identical to the host front end."""
        t = self.table
        self.ctx.table_upload(t.user_ptr, t.item_ids, t.ratings, t.dataset_type)
        for step in ("byUser", "byItem", "rmseValidate", "rmseTest"):
            pto = np.asarray(self.portionsRowIdTo[step], np.int32)
            if self.world == 1 or len(pto) == 0:
                self.my_portions[step] = (0, len(pto))
            else:                                                # nnz-balanced contiguous slice, from the device counts
                cnt = self.ctx.table_counts(STEP_MASK[step], step == "byItem")
                if step in ("byUser", "byItem"):
                    ends = ydist.cost_ends(cnt, pto, self.factorsCount)
                else:
                    ends = np.cumsum(cnt, dtype=np.int64)[pto.astype(np.int64) - 1]
                cuts = ydist.balanced_cuts(ends, self.world)
                self.my_portions[step] = (int(cuts[self.rank]), int(cuts[self.rank + 1]))
            lo, hi = self.my_portions[step]                      # this rank's contiguous slice of the plan
            first_row = 0 if lo == 0 else int(pto[lo - 1])
            if hi > lo:
                rid = self.ctx.rowset_from_table(native.STEP_TYPES[step], STEP_MASK[step], pto[lo:hi], first_row)
            else:                                                # a rank without portions: an empty row set
                rid = self.ctx.rowset_from_table(native.STEP_TYPES[step], STEP_MASK[step],
                                                 np.asarray([first_row], np.int32), first_row)
            self.rowsets[step] = rid
            self.rowlists[step] = None

    def prepareBulk(self):
        """Upload every step's portions once as a device-resident row set."""
        if self.options["gpu"].get("deviceIngest", False):
            return self.prepareBulkOnDevice()
        for step in ("byUser", "byItem", "rmseValidate", "rmseTest"):
            csr = self._csr(step)
            lo, hi = self.my_portions[step]
            pto = np.asarray(self.portionsRowIdTo[step], np.int32)
            rl = fe.build_rowlist(csr, pto)
            r0, r1 = int(rl.portion_first[lo]), int(rl.portion_first[hi])
            ids, start, ln = rl.row_ids[r0:r1], rl.row_start[r0:r1], rl.row_len[r0:r1]
            # upload only this rank's span of the ratings arrays
            if len(ids):
                s0 = int(start[0])
                s1 = int((start + ln).max())
            else:
                s0 = s1 = 0
            pf = (rl.portion_first[lo:hi + 1] - r0).astype(np.int32)
            rid = self.ctx.rowset_create(native.STEP_TYPES[step], np.ascontiguousarray(ids),
                                         np.ascontiguousarray(start - s0), np.ascontiguousarray(ln),
                                         csr.idx[s0:s1], csr.vals[s0:s1], pf if len(pf) > 1 else None)
            self.rowlists[step] = (ids, ln, pf)
            self.rowsets[step] = rid

    def _createNodeSharedFactors(self):
        import os
        k = self.factorsCount
        tag = "%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid())

        def init(mats):
            self.userFactors, self.itemFactors = mats
            self.initSharedFactorsRandom()

        self.userFactors, self.itemFactors = ydist.node_shared_matrices(
            tag, [(self.totalUsersCount, k), (self.totalItemsCount, k)], self.rank, init, self.group)
        self.sharedHost = True

    def getMemoryUsage(self):
        """EmfBase.getMemoryUsage (EmfBase.js:880-900): [factor segments ("shm"), this process' rss, one rss per
        worker]; the workers answer 'getMemoryUsage' with 'setMemoryUsage' (EmfWorker.js:43,109-113)."""
        import resource
        self._workerMU = []
        self._prepared = {}
        self._rmseGlobal = {}
        self._alsSteps = 0
        for w in self.workers:
            w.master_side.emit("getMemoryUsage")
        shm = sum(int(a.nbytes) for a in (self.userFactors, self.itemFactors) if a is not None)
        return [shm, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss * 1024] + [mu["rss"] for mu in self._workerMU]

    def endTrain(self):
        for w in self.workers:
            w.master_side.emit("endTrain")
        self.workers = []
        self.ctx = None

    # ---- ALS half-step ------------------------------------------------------------------------
    def _solved_range(self, step):
        """Row-id range [lo, hi) this rank solves in `step` (contiguous by construction)."""
        lo, hi = self.my_portions[step]
        pto = self.portionsRowIdTo[step]
        a = 0 if lo == 0 else int(pto[lo - 1])
        b = a if hi == lo else int(pto[hi - 1])
        return a, b

    def alsTrainStep(self, stepType):
        """EmfLord.alsTrainStep (EmfLord.js:963-984)."""
        self._alsSteps += 1
        mp = self.workers[0].master_side
        if self.options["gpu"]["bulk"]:
            self.ctx.als_rowset(self.rowsets[stepType])
        else:
            csr = self._csr(stepType)
            pto = self.portionsRowIdTo[stepType]
            lo, hi = self.my_portions[stepType]
            pb = self.workers[0].portionBuffer
            self.completedPortions = 0
            mp.emit("startTrainStep", {"stepType": stepType})
            cache = getattr(self, "portionCache", None)
            if cache is not None and self.options["gpu"].get("nativeLoop", False):
                # the per-portion calls of the whole half-step issued by native code (ycnr_als_portions): what an
                # N-API binding costs per message is microseconds, a Python message round trip is ~20 us
                prep = self._preparedPortions(stepType, "als", lo, hi)
                self.ctx.als_portions(prep)
                self.completedPortions += prep["n"]
                self.h2d_bytes += prep["h2d"]
                lo = hi
            for p in range(lo, hi):
                if cache is not None:
                    cb = cache[stepType][p]
                    self.h2d_bytes += 8 * cb["fetched"] + 4 * len(cb["alsRows"])
                    mp.emit("calcTrainAlsPortion", {"portionNo": p, "portionBuffer": cb})
                    continue
                row_from = 0 if p == 0 else int(pto[p - 1])
                # fill the worker's buffer (EmfMaster.js:571-614, 656-658)
                fetched = fe.build_portion_into(csr, row_from, int(pto[p]), pb["alsRows"], pb["alsIndx"], pb["alsVals"])
                self.h2d_bytes += 8 * fetched + 4 * (2 * int(pb["alsRows"][0]) + 1)
                mp.emit("calcTrainAlsPortion", {"portionNo": p})
            mp.emit("endTrainStep")
            a, b = self._solved_range(stepType)
            self.d2h_bytes += (b - a) * self.factorsCount * 4
        if self.world > 1:
            self._refresh_replicas(stepType)

    def connectPeers(self):
        """Fused all-gather ('alsSaveCalcedFactors' without a separate exchange step): the solve kernels
        store every solved row into all replicas over NVLink; a half-step then ends with one barrier."""
        for which in (native.USER_FACTORS, native.ITEM_FACTORS):
            ydist.connect_peers(self.ctx, which, self.rank, self.world, self.group)
        self.fusedPeers = True

    def _refresh_replicas(self, stepType):
        which = native.USER_FACTORS if stepType == "byUser" else native.ITEM_FACTORS
        if self.fusedPeers and self.options["gpu"]["bulk"]:
            # my peer stores are complete when my kernels are, everybody else's into my replica when theirs are:
            # a barrier in stream order, the host moves on
            ydist.stream_barrier(self.ctx, self.group)
            return
        if self.fusedPeers and self.sharedHost:
            self.ctx.synchronize()
            ydist.barrier(self.group)
            return
        if stepType not in self._ranges:    # static per step: the portion plan does not change between iterations
            self._ranges[stepType] = ydist.all_ranges(self._solved_range(stepType), self.world, self.group)
        ranges = self._ranges[stepType]
        # per-portion mode: the host segment is the truth.  With one segment shared by all ranks every rank
        # has already written its own rows (ycnr_end_train_step); private segments need the peers' rows too.
        private_host = not self.options["gpu"]["bulk"] and not self.sharedHost
        ydist.refresh_replicas(self.ctx, which, self.factorsCount, ranges, self.rank, self.group,
                               host=(self.userFactors if which == 0 else self.itemFactors) if private_host else None)

    def alsTrainIter(self):
        """EmfLord.alsTrainIter (EmfLord.js:954-958)."""
        self.alsTrainStep("byUser")
        self.alsTrainStep("byItem")

    # ---- RMSE --------------------------------------------------------------------------------
    def wm_completedPortion(self, msg):
        """EmfMaster.m_completedPortion (EmfMaster.js:757-786), accumulation part."""
        self.completedPortions += 1
        if "rSumDiff2" in msg:
            self.rSumDiff2 += msg["rSumDiff2"]
            self.rCnt += msg["rCnt"]
            self.rSum += msg["rSum"]
            self._lastRmseMsg = msg

    def wm_completedPortions(self, tags, infos):
        """m_completedPortion for all replies of a pass at once (the native multi-portion path): the same additions
        in the same (queue) order — np.cumsum adds sequentially — and the same 'last message' (quirk Q7)."""
        n = len(tags)
        if n == 0:
            return
        self.completedPortions += n
        self.rSumDiff2 = float(np.cumsum(np.concatenate(([self.rSumDiff2], infos["r_sum_diff2"])))[-1])
        self.rCnt = float(np.cumsum(np.concatenate(([self.rCnt], infos["r_cnt"])))[-1])
        self.rSum = float(np.cumsum(np.concatenate(([self.rSum], infos["r_sum"])))[-1])
        self._lastRmseMsg = {"portionNo": int(tags[-1]), "rSumDiff2": float(infos["r_sum_diff2"][-1]),
                             "rCnt": float(infos["r_cnt"][-1]), "rSum": float(infos["r_sum"][-1])}

    def _gatherRmseShift0(self):
        """world > 1, bulk: 'rmseSaveCalcs' (EmfMaster.js:726-736) for the validate AND the test pass in one
        all-gather — sums at shift 0 plus the sum of the ratings, with which every rank derives the pass at any
        other shift (the third pass) without another exchange."""
        d = self.options["dataSetDistr"]
        vec = []
        for step, pct in (("rmseValidate", d[1]), ("rmseTest", d[2])):
            lo, hi = self.my_portions[step]
            if pct and hi > lo:
                tot, ps = self.ctx.rmse_rowset(self.rowsets[step], 0.0, hi - lo)
                rsum, rlast = self.ctx.rmse_rowset_ratings(self.rowsets[step])
                vec += [tot[0], tot[1], tot[2], rsum, ps[-1, 2], ps[-1, 1], 1.0]
            else:
                vec += [0.0] * 7
        rows = ydist.all_gather_doubles(vec, self.group)
        for si, step in enumerate(("rmseValidate", "rmseTest")):
            g = {"D": 0.0, "C": 0.0, "P": 0.0, "R": 0.0, "lastP": 0.0, "lastC": 0.0, "ver": self._alsSteps}
            for r in rows:                               # rank order: deterministic sums; last portion = highest rank
                v = r[7 * si:7 * si + 7]
                g["D"] += v[0]
                g["C"] += v[1]
                g["P"] += v[2]
                g["R"] += v[3]
                if v[6] != 0.0:
                    g["lastP"], g["lastC"] = v[4], v[5]
            self._rmseGlobal[step] = g

    def calcRmse(self, stepType, useGlobalAvgShift):
        """EmfLord.calcRmse (EmfLord.js:1043-1081) + EmfMaster._startCalcRmse (389-412)."""
        d = self.options["dataSetDistr"]
        if (d[1] == 0 and stepType == "rmseValidate") or (d[2] == 0 and stepType == "rmseTest"):
            return None
        calcGlobalAvgShift = not useGlobalAvgShift
        if calcGlobalAvgShift:
            self.globalAvgShift = 0.0
        g = self._rmseGlobal.get(stepType)
        if self.world > 1 and self.options["gpu"]["bulk"] and g and g["ver"] == self._alsSteps:
            s_ = self.globalAvgShift                     # every rating's prediction moves by the shift
            self.rSumDiff2 = g["D"] - 2.0 * s_ * (g["R"] - g["P"]) + g["C"] * s_ * s_
            self.rCnt, self.rSum = g["C"], g["P"] + g["C"] * s_
            self.rmse = math.sqrt(self.rSumDiff2 / self.rCnt) if self.rCnt else float("nan")
            self.predAvg = (g["lastP"] + g["lastC"] * s_) / g["lastC"] if g["lastC"] else float("nan")
            if calcGlobalAvgShift:
                self.globalAvgShift = self.stats["totalRatingsAvg"] - self.predAvg
            return self.rmse
        self.rSum = self.rSumDiff2 = self.rCnt = 0.0
        self._lastRmseMsg = None
        lo, hi = self.my_portions[stepType]
        if self.options["gpu"]["bulk"]:
            if hi > lo:
                tot, ps = self.ctx.rmse_rowset(self.rowsets[stepType], self.globalAvgShift, hi - lo)
                self.rSumDiff2, self.rCnt, self.rSum = tot
                self._lastRmseMsg = {"rSum": ps[-1, 2], "rCnt": ps[-1, 1]}
        else:
            mp = self.workers[0].master_side
            csr = self._csr(stepType)
            pto = self.portionsRowIdTo[stepType]
            pb = self.workers[0].portionBuffer
            mp.emit("startCalcRmse", {"stepType": stepType, "globalAvgShift": self.globalAvgShift})
            cache = getattr(self, "portionCache", None)
            if cache is not None and self.options["gpu"].get("nativeLoop", False):
                prep = self._preparedPortions(stepType, "rmse", lo, hi)
                self.ctx.rmse_portions_async(prep)
                self.wm_completedPortions(*self.ctx.rmse_poll_arrays(True))
                self.h2d_bytes += prep["h2d"]
                self.d2h_bytes += 24 * prep["n"]
                lo = hi
            for p in range(lo, hi):
                if cache is not None:
                    cb = cache[stepType][p]
                    self.h2d_bytes += 8 * cb["fetched"] + 4 * len(cb["rmseRows"])
                    mp.emit("calcRmsePortion", {"portionNo": p, "portionBuffer": cb})
                    self.d2h_bytes += 24
                    continue
                row_from = 0 if p == 0 else int(pto[p - 1])
                fetched = fe.build_portion_into(csr, row_from, int(pto[p]), pb["rmseRows"], pb["rmseIndx"], pb["rmseVals"])
                self.h2d_bytes += 8 * fetched + 4 * (2 * int(pb["rmseRows"][0]) + 1)
                mp.emit("calcRmsePortion", {"portionNo": p})
                self.d2h_bytes += 24
            mp.emit("endCalcRmse")
        last = self._lastRmseMsg
        if self.world > 1:   # 'rmseSaveCalcs' reduce-to-root + the last portion's partials (Q7)
            (self.rSumDiff2, self.rCnt, self.rSum), last = ydist.reduce_rmse(
                (self.rSumDiff2, self.rCnt, self.rSum), last, self.group)
        self.rmse = math.sqrt(1.0 * self.rSumDiff2 / self.rCnt) if self.rCnt else float("nan")
        # Q7: predAvg from the LAST completed portion's message (EmfMaster.js:779)
        self.predAvg = last["rSum"] / last["rCnt"] if last and last["rCnt"] else float("nan")
        if calcGlobalAvgShift:
            self.globalAvgShift = self.stats["totalRatingsAvg"] - self.predAvg
        return self.rmse

    # ---- train loop ------------------------------------------------------------------------------
    def _timed(self, name, fn, *args):
        """Per-portion mode only (every phase ends host-synchronous there): wall time per phase, the
        analogue of the reference's per-step console timings (EmfLord.js:1075-1077)."""
        if self.options["gpu"]["bulk"]:
            return fn(*args)
        t0 = time.perf_counter()
        out = fn(*args)
        self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    def trainIter(self):
        """One pass of the loop body of EmfLord.train (EmfLord.js:892-902)."""
        self._timed("byUser", self.alsTrainStep, "byUser")
        self._timed("byItem", self.alsTrainStep, "byItem")
        if self.options["gpu"]["bulk"]:
            # the validate and the test pass both run with shift 0 (EmfMaster.js:389-402): queue them together, the
            # calcRmse calls below only collect; the third pass is derived from the second on the host
            d = self.options["dataSetDistr"]
            for step, pct in (("rmseValidate", d[1]), ("rmseTest", d[2])):
                lo, hi = self.my_portions[step]
                if pct and hi > lo:
                    self.ctx.rmse_rowset_begin(self.rowsets[step], 0.0)
            if self.world > 1:
                self._gatherRmseShift0()
        out = {
            "rmseValidate": self._timed("rmseValidate", self.calcRmse, "rmseValidate", False),
            "rmseTest": self._timed("rmseTest", self.calcRmse, "rmseTest", False),
            "rmseTestShift": self._timed("rmseTestShift", self.calcRmse, "rmseTest", True),
            "globalAvgShift": self.globalAvgShift,
        }
        self.history.append(out)
        return out

    def train(self, iters=None):
        for _ in range(self.options["trainIters"] if iters is None else iters):
            self.trainIter()
        if self.options["gpu"]["bulk"]:
            self.syncFactorsToHost()
        return self.history

    # ---- serving (YcnrController.recommendItemsForUser, lib/YcnrController.js:227-284) -----------------
    def recommendItemsForUsers(self, users, limit=20, minRecommendRating=0.0):
        """Batched form of the controller's brute-force top-N.  users: dicts {'list_id': 1-based id,
        'unrated_items': [1-based item ids]} as upstream; the rated items come from the ratings table
        (the SQL at 233-239).  Returns per user a list of {'predict', 'id' (1-based)}, best first —
        at most limit-1 entries, like upstream (281-282)."""
        t = self.table
        uids, skips = [], []
        for u in users:
            u0 = int(u["list_id"]) - 1
            rated = t.item_ids[t.user_ptr[u0]:t.user_ptr[u0 + 1]]
            unrated = np.asarray(u.get("unrated_items") or [], np.int32) - 1
            uids.append(u0)
            skips.append(np.concatenate([rated, unrated]).astype(np.int32))
        out = self.ctx.recommend_batch(uids, skips, limit, minRecommendRating, self.globalAvgShift)
        return [[{"predict": p, "id": i + 1} for i, p in rec] for rec in out]

    def recommendItemsForUser(self, user, limit=20, minRecommendRating=0.0):
        return self.recommendItemsForUsers([user], limit, minRecommendRating)[0]

    def syncFactorsToHost(self):
        """Bulk mode keeps the factors on the device; copy both matrices into the host segments."""
        self.ctx.download_factors(native.USER_FACTORS)
        self.ctx.download_factors(native.ITEM_FACTORS)
        self.d2h_bytes += (self.totalUsersCount + self.totalItemsCount) * self.factorsCount * 4
