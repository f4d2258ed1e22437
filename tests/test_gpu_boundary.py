"""Boundary hardening on the GPU: real SysV shared-memory factor segments, header / id validation of the
per-portion path, the upstream-arity gather export, the memory-usage message, two worker contexts on one GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import make_problem, oracle_portions, portion_from_rows, rel_fro, worst_row_rel
from you_can_not_recommend_b200 import native
from you_can_not_recommend_b200.emf_master import EmfMaster

pytestmark = pytest.mark.gpu

IPC_PRIVATE, IPC_CREAT, IPC_RMID = 0, 0o1000, 0


class SysVSegment:
    """shm-typed-array's create()/get() (EmfBase.js:403-412, 430-450) through libc: shmget + shmat."""

    def __init__(self, nbytes=None, shmid=None):
        self.libc = C.CDLL("libc.so.6", use_errno=True)
        self.libc.shmat.restype = C.c_void_p
        self.libc.shmat.argtypes = [C.c_int, C.c_void_p, C.c_int]
        self.libc.shmdt.argtypes = [C.c_void_p]
        self.owner = shmid is None
        if shmid is None:
            shmid = self.libc.shmget(IPC_PRIVATE, C.c_size_t(nbytes), IPC_CREAT | 0o600)
            if shmid < 0:
                raise OSError(C.get_errno(), "shmget")
        self.shmid = shmid
        self.addr = self.libc.shmat(shmid, None, 0)
        if self.addr in (None, C.c_void_p(-1).value):
            raise OSError(C.get_errno(), "shmat")

    def array(self, shape):
        n = int(np.prod(shape))
        buf = (C.c_float * n).from_address(self.addr)
        return np.frombuffer(buf, np.float32).reshape(shape)

    def close(self):
        self.libc.shmdt(C.c_void_p(self.addr))
        if self.owner:
            self.libc.shmctl(self.shmid, IPC_RMID, None)


def test_factors_in_real_sysv_segments():
    """The factor store as upstream has it: two SysV segments created by the master, attached by the worker
    (page-locked by ycnr_attach_factors), trained into in place, read through a SECOND attachment of the same
    segment (what the master / recommender processes see, EmfManager.js:479-489)."""
    prob = make_problem("ml-100k", k=20)
    t = prob["table"]
    segU, segV = SysVSegment(t.users * 20 * 4), SysVSegment(t.items * 20 * 4)
    U, V = segU.array((t.users, 20)), segV.array((t.items, 20))
    U[...] = prob["U0"]
    V[...] = prob["V0"]
    ref = oracle.OracleTrainer(prob["U0"], prob["V0"], oracle_portions(prob), 0.05, 0.05, prob["total_ratings_avg"], np.float32)
    want = ref.train_iter()
    m = EmfMaster(t, {"factorsCount": 20, "seed": prob["seed"]})
    m.prepareToTrain(U, V)                       # openSharedFactors: adopts the segments, no copy
    assert m.userFactors.ctypes.data == segU.addr
    got = m.trainIter()
    mu = m.getMemoryUsage()
    assert mu[0] == (t.users + t.items) * 20 * 4 and len(mu) == 3 and mu[2] > 0
    m.endTrain()
    for key in ("rmseValidate", "rmseTest", "rmseTestShift"):
        assert abs(got[key] - want[key]) < 1e-4
    # a second attachment of the same segments (another process would do exactly this)
    seg2U, seg2V = SysVSegment(shmid=segU.shmid), SysVSegment(shmid=segV.shmid)
    assert seg2U.addr != segU.addr
    U2, V2 = seg2U.array((t.users, 20)), seg2V.array((t.items, 20))
    assert rel_fro(U2, ref.U) < 1e-3 and rel_fro(V2, ref.V) < 1e-3
    assert (U2 == U).all() and not (U2 == prob["U0"]).all()
    for s in (seg2U, seg2V, segU, segV):
        s.close()


def test_portion_ids_are_validated():
    """A stale / corrupt portion must not become out-of-bounds device writes: row ids are checked on the host
    (range, strictly ascending), column ids on the device; the error surfaces at the step barrier and the
    context stays usable."""
    k = 20
    rng = np.random.default_rng(1)
    F = rng.normal(0, 0.3, (50, k)).astype(np.float32)
    S = np.full((8, k), 2.0, np.float32)
    ctx = native.Context(k, 8, 50)
    ctx.attach_factors(S, F)
    good = portion_from_rows([1, 4], [[3, 7, 9], [0, 49]], [[1, 2, 3], [4, 5]])
    for ids, msg in (([1, 8], "outside the factor matrix"), ([-1, 2], "outside the factor matrix"),
                     ([4, 1], "ascend"), ([3, 3], "ascend")):
        bad = portion_from_rows(ids, [[3, 7, 9], [0, 49]], [[1, 2, 3], [4, 5]])
        ctx.start_train_step(native.BY_USER)
        with pytest.raises(RuntimeError, match=msg):
            ctx.als_portion(*bad)
        ctx.end_train_step()
    assert (S == 2.0).all()
    for cols in ([[3, 7, 50], [0, 1]], [[3, -2, 9], [0, 1]]):
        bad = portion_from_rows([1, 4], cols, [[1, 2, 3], [4, 5]])
        ctx.start_train_step(native.BY_USER)
        ctx.als_portion(*bad)                          # queued; the device check raises the flag
        with pytest.raises(RuntimeError, match="column id outside"):
            ctx.end_train_step()
        assert (S == 2.0).all()                        # nothing was written
        ctx.start_calc_rmse(native.RMSE_TEST, 0.0)
        with pytest.raises(RuntimeError, match="item id outside"):
            ctx.rmse_portion(*bad)
    with pytest.raises(RuntimeError, match="do not fit"):
        ctx.start_train_step(native.BY_USER)
        ctx.als_portion(good[0], good[1][:3], good[2])
    ctx.end_train_step()
    # the context still works
    ctx.start_train_step(native.BY_USER)
    ctx.als_portion(*good)
    ctx.end_train_step()
    want = S.copy()
    want[:] = 2.0
    oracle.als_portion(good[0], good[1], good[2], F, want, 0.05)
    assert worst_row_rel(S, want) < 1e-3 and not (S[1] == 2.0).all() and (S[[0, 2, 3, 5, 6, 7]] == 2.0).all()
    ctx.close()


def test_gather_export_upstream_arity_and_attached_fast_path():
    """sAlsBuildSubFixedFacts(sub, fixed, indx, cols, k) with upstream's five arguments (cpp_utils.js:15-19) on the
    process's current context; when `fixed` is the attached matrix the rows come from the device replica (only
    the ids are uploaded), otherwise row by row from the host matrix — bit-exact both ways."""
    rng = np.random.default_rng(4)
    k = 100
    U = rng.normal(0, 1, (300, k)).astype(np.float32)
    V = rng.normal(0, 1, (2000, k)).astype(np.float32)
    ctx = native.Context(k, 300, 2000, profile=True)
    ctx.attach_factors(U, V)
    indx = rng.integers(0, 2000, 333).astype(np.int32)
    sub = np.zeros((333, k), np.float32)
    ctx.profile_reset()
    native.build_sub_fixed_facts_noctx(sub, V, indx, 333, k)
    assert (sub == V[indx]).all() and ctx.profile_read()["gather"]["launches"] == 1     # replica gather kernel
    other = rng.normal(0, 1, (500, k)).astype(np.float32)
    sub2 = np.zeros((77, k), np.float32)
    ctx.profile_reset()
    ctx.build_sub_fixed_facts(sub2, other, indx[:77] % 500)
    assert (sub2 == other[indx[:77] % 500]).all() and ctx.profile_read()["gather"]["launches"] == 0
    ctx.close()
    with pytest.raises(RuntimeError, match="no live context"):
        native.build_sub_fixed_facts_noctx(sub, V, indx, 333, k)


def test_two_worker_contexts_on_one_gpu():
    """numThreadsForTrain.als = 2 (upstream default is numCPUs workers, EmfBase.js:110): two worker contexts
    share one GPU and the factor segments, each solving every other portion of a half-step; same result as one."""
    prob = make_problem("ml-100k", k=20)
    t = prob["table"]
    por = oracle_portions(prob, ("byUser",))["byUser"]
    U1, V = prob["U0"].copy(), prob["V0"].copy()
    one = native.Context(20, t.users, t.items)
    one.attach_factors(U1, V)
    one.start_train_step(native.BY_USER)
    for p in por:
        one.als_portion(*p)
    one.end_train_step()
    one.close()
    U2 = prob["U0"].copy()
    a, b = native.Context(20, t.users, t.items), native.Context(20, t.users, t.items)
    for c in (a, b):
        c.attach_factors(U2, V)
        c.start_train_step(native.BY_USER)
    for i, p in enumerate(por):
        (a if i % 2 == 0 else b).als_portion(*p)
    for c in (a, b):
        c.end_train_step()
    assert (U2 == U1).all() and not (U2 == prob["U0"]).all()
    mu = a.memory_usage()
    assert mu["device"] >= (t.users + t.items) * 20 * 4 and mu["deviceTotal"] > mu["deviceFree"] > 0
    a.close()
    b.close()
