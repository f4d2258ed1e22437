"""Per-class device times of an e2e (per-portion, native loop) iteration: where the GPU time of the batched path goes.
  python scripts/e2e_classes.py [workload] [k] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from you_can_not_recommend_b200 import build  # noqa: E402

build.build_cuda(force=bool(os.environ.get("YCNR_NVCC_FLAGS")))
from you_can_not_recommend_b200 import front_end as fe  # noqa: E402
from you_can_not_recommend_b200.emf_master import EmfMaster  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "mal"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
table = fe.synth_table(workload)
m = EmfMaster(table, {"factorsCount": k, "gpu": {"bulk": False, "profile": True, "cachePortions": True, "nativeLoop": True
}})
m.prepareToTrain()
for _ in range(3):
    m.trainIter()
m.ctx.synchronize()
res = {}
for phase in ("byUser", "byItem"):
    m.ctx.profile_reset()
    t0 = time.perf_counter()
    for _ in range(steps):
        m.alsTrainStep(phase)
    m.ctx.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    prof = m.ctx.profile_read()
    res[phase] = {"wall_ms": round(wall, 2),
                  "classes": {c: [round(v["ms"] / steps, 3), v["launches"] // steps] for c, v in prof.items() if isinstance(v, dict) and v["launches"]}}
print(json.dumps(res))
m.endTrain()
