"""Per-class / per-dual-bin device times of one bulk iteration (tuning experiments; not the bench of record).
  YCNR_NVCC_FLAGS="-DYCNR_DUAL_TPT=4" python scripts/quick_bench.py [workload] [k] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from you_can_not_recommend_b200 import build  # noqa: E402

build.build_cuda(force=bool(os.environ.get("YCNR_NVCC_FLAGS")))
from you_can_not_recommend_b200 import front_end as fe  # noqa: E402
from you_can_not_recommend_b200.emf_master import EmfMaster  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "mal"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
table = fe.synth_table(workload)
m = EmfMaster(table, {"factorsCount": k, "gpu": {"bulk": True, "profile": True}})
m.prepareToTrain()
for _ in range(2):
    m.trainIter()
m.ctx.synchronize()
m.ctx.profile_reset()
import time  # noqa: E402
t0 = time.perf_counter()
for _ in range(steps):
    out = m.trainIter()
m.ctx.synchronize()
wall = (time.perf_counter() - t0) / steps * 1e3
prof = m.ctx.profile_read()
bins = m.ctx.profile_dual_bins()
res = {"flags": os.environ.get("YCNR_NVCC_FLAGS", ""), "workload": workload, "k": k, "wall_ms_per_step": wall,
       "classes": {c: round(v["ms"] / steps, 3) for c, v in prof.items() if isinstance(v, dict) and v["launches"]},
       "dual_bins": {mt: [round(ms / steps, 3), rows // steps] for mt, ms, rows in bins}, "rmse": out}
print(json.dumps(res))
m.endTrain()
