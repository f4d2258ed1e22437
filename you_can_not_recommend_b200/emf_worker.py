"""EmfWorker — the reference worker's message interface over the B200 library.

Mirror of lib/emf/EmfWorker.js: same handler names, same message fields, same portion
buffer triple (alsRows/alsIndx/alsVals, rmseRows/rmseIndx/rmseVals — SURVEY.md §5.4),
same 'completedPortion' reply.  Only the bodies of mw_calcTrainAlsPortion /
mw_calcRmsePortion changed: they hand the buffers to the CUDA library instead of
looping over rows with BLAS/LAPACK calls.

One addition the reference does not have: 'endTrainStep' (the master sends it when
all portions of a half-step completed, EmfMaster.js:776-785) makes the solved rows
visible in the host factor segment — portions are queued asynchronously on the GPU.
"""
import numpy as np

from . import native
from .emf_base import EmfBase


class EmfProcess:
    """EmfProcess (lib/emf/EmfProcess.js:38-62): {msg, data} packets between master and worker.
    In-process stand-in for child_process IPC: emit() calls the peer's registered handler."""

    def __init__(self):
        self._handlers = {}
        self.peer = None

    def on(self, msg, fn):
        self._handlers[msg] = fn

    def emit(self, msg, data=None):
        if self.peer is not None:
            self.peer.onMessage(msg, data or {})

    def onMessage(self, msg, data):
        fn = self._handlers.get(msg)
        if fn is None:
            raise KeyError("unhandled message '%s'" % msg)
        return fn(data)


class EmfWorker(EmfBase):
    def __init__(self, workerId=0, workerProcess=None, options=None):
        super().__init__(options)
        assert workerId != -1
        self.workerId = workerId
        self.process = workerProcess or EmfProcess()
        self.portionBuffer = None
        self.workType = None
        self.stepType = None
        self._status = "ready"
        p = self.process
        # handler table of EmfWorker.init (EmfWorker.js:43-51)
        p.on("getMemoryUsage", self.mw_getMemoryUsage)
        p.on("prepareToTrain", self.mw_prepareToTrain)
        p.on("startTrain", self.mw_startTrain)
        p.on("endTrain", self.mw_endTrain)
        p.on("startTrainStep", self.mw_startTrainStep)
        p.on("endTrainStep", self.mw_endTrainStep)
        p.on("startCalcRmse", self.mw_startCalcRmse)
        p.on("endCalcRmse", self.mw_endCalcRmse)
        p.on("calcTrainAlsPortion", self.mw_calcTrainAlsPortion)
        p.on("calcTrainSgdPortion", self.mw_calcTrainSgdPortion)
        p.on("calcRmsePortion", self.mw_calcRmsePortion)

    def mw_getMemoryUsage(self, data=None):
        """EmfWorker.mw_getMemoryUsage (EmfWorker.js:109-113): process.memoryUsage() of the worker, here the
        resident set of this process plus what the context holds on the device and in page-locked memory."""
        import resource
        mu = {"rss": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss * 1024, "heapTotal": 0, "heapUsed": 0}
        if self.ctx is not None:
            mu.update(self.ctx.memory_usage())
        self.process.emit("setMemoryUsage", {"mu": mu})

    # -- lifecycle (EmfWorker.js:119-164) -------------------------------------------------
    def mw_prepareToTrain(self, data):
        self._status = "preparing"
        self.stats = data["stats"]
        self.options = data["options"]
        self.factorsCount = int(self.options["factorsCount"])
        self.totalUsersCount = data["totalUsersCount"]
        self.totalItemsCount = data["totalItemsCount"]
        self.openSharedFactors(data["shared"]["userFactors"], data["shared"]["itemFactors"])
        self.openWorkPortionBuffers(data["shared"]["portionBuffer"])
        self.openDevice()
        self._status = "ready"
        self.process.emit("preparedToTrain")

    def openWorkPortionBuffers(self, pb):
        """EmfWorker.openWorkPortionBuffers (EmfWorker.js:66-89): adopt the master's buffers."""
        for name, dt in (("alsRows", np.int32), ("alsIndx", np.int32), ("alsVals", np.float32),
                         ("rmseRows", np.int32), ("rmseIndx", np.int32), ("rmseVals", np.float32)):
            assert pb[name].dtype == dt and pb[name].flags.c_contiguous, name
        self.portionBuffer = pb

    def mw_startTrain(self, data=None):
        self._status = "training"

    def mw_endTrain(self, data=None):
        self.portionBuffer = None
        self.closeDevice()
        self._status = "ready"

    def mw_startTrainStep(self, msg):
        self.workType = "train"
        self.stepType = msg["stepType"]
        self.ctx.start_train_step(native.STEP_TYPES[self.stepType])

    def mw_endTrainStep(self, msg=None):
        self.ctx.end_train_step()
        self.process.emit("endedTrainStep", {"stepType": self.stepType})

    def mw_startCalcRmse(self, msg):
        self.workType = "rmse"
        self.stepType = msg["stepType"]
        self.globalAvgShift = msg["globalAvgShift"]
        self.ctx.start_calc_rmse(native.STEP_TYPES[self.stepType], self.globalAvgShift)

    # -- the hot path ---------------------------------------------------------------------
    def mw_calcTrainAlsPortion(self, msg):
        """EmfWorker.mw_calcTrainAlsPortion (EmfWorker.js:169-261)."""
        pb = msg.get("portionBuffer") or self.portionBuffer     # cached portion (usePortionsCache) or the work buffer
        info = self.ctx.als_portion(pb["alsRows"], pb["alsIndx"], pb["alsVals"])
        self.process.emit("completedPortion", {
            "portionNo": msg["portionNo"],
            "rowsRange": {"from": info.rows_from if info.rows_cnt > 0 else None, "cnt": info.rows_cnt},
            "ratingsInPortion": info.ratings_in_portion,
            "time": info.time_ms,
            "memoryUsage": None,
        })

    def mw_calcRmsePortion(self, msg):
        """EmfWorker.mw_calcRmsePortion (EmfWorker.js:266-315).  The portion is queued on the GPU (small portions
        are launched in batches); its 'completedPortion' reply — an asynchronous message upstream as well
        (EmfWorker.js:304-314) — is emitted as soon as its sums are back, always in portion order."""
        pb = msg.get("portionBuffer") or self.portionBuffer
        self.ctx.rmse_portion_async(pb["rmseRows"], pb["rmseIndx"], pb["rmseVals"], msg["portionNo"])
        self._rmseQueued = getattr(self, "_rmseQueued", 0) + 1
        if self._rmseQueued % 8 == 0:            # replies are asynchronous anyway: look for finished portions now and then
            self._emitCompletedRmse(False)

    def mw_endCalcRmse(self, msg=None):
        """Addition (like endTrainStep): the master has handed out the last portion of the pass — flush the
        queue and send the remaining 'completedPortion' replies."""
        self._emitCompletedRmse(True)

    def _emitCompletedRmse(self, wait):
        for tag, info in self.ctx.rmse_poll(wait):
            self.process.emit("completedPortion", {
                "portionNo": int(tag),
                "rowsRange": {"from": info.rows_from if info.rows_cnt > 0 else None, "cnt": info.rows_cnt},
                "ratingsInPortion": info.ratings_in_portion,
                "time": info.time_ms,
                "memoryUsage": None,
                "rSumDiff2": info.r_sum_diff2,
                "rCnt": info.r_cnt,
                "rSum": info.r_sum,
            })

    def mw_calcTrainSgdPortion(self, msg):
        raise NotImplementedError("SGD is deprecated upstream (README.md:13) and not on the B200 path")
