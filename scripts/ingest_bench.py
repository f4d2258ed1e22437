"""Front-end timing on the MAL shape: host front end (CSR build + portion headers + upload through
ycnr_rowset_create) against the device-side ingest (ycnr_table_upload + ycnr_rowset_from_table), and a check
that one training iteration gives identical RMSE values either way.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from you_can_not_recommend_b200 import front_end as fe
from you_can_not_recommend_b200.emf_master import EmfMaster

workload = sys.argv[1] if len(sys.argv) > 1 else "mal"
table = fe.synth_table(workload)
k = fe.SHAPES[workload]["factors"]
out = {"workload": workload, "ratings": table.nnz}
hist = {}
for dev in (False, True):
    m = EmfMaster(table, {"factorsCount": k, "gpu": {"bulk": False, "deviceIngest": dev}})
    m.options["gpu"]["bulk"] = False
    m.prepareToTrain()                      # split, stats, plan, factors, worker, device context (common to both)
    table._cache.clear()                    # the host CSR caches: time the fetch itself
    m.options["gpu"]["bulk"] = True
    m.ctx.synchronize()
    t0 = time.perf_counter()
    m.prepareBulk()
    m.ctx.synchronize()
    out["device_ingest_s" if dev else "host_front_end_s"] = time.perf_counter() - t0
    hist[dev] = m.trainIter()
    m.endTrain()
# first-time split (EmfLord.doSplitToSets; README.md:127: 1 h 05 min on MAL upstream): host threads vs device
import numpy as np
from you_can_not_recommend_b200 import native
want = table.dataset_type.copy()
t0 = time.perf_counter()
fe.split_sets(table, (85, 10, 5), seed=fe.DEFAULT_SEED + 1)
out["host_split_s"] = time.perf_counter() - t0
ctx = native.Context(k, table.users, table.items)
ctx.table_upload(table.user_ptr, table.item_ids, table.ratings, np.zeros(table.nnz, np.int8))
t0 = time.perf_counter()
got = ctx.table_split(fe.DEFAULT_SEED + 1, (85, 10, 5), table.nnz)
out["device_split_s"] = time.perf_counter() - t0
out["same_split"] = bool((got == table.dataset_type).all() and (got == want).all())
ctx.close()
out["same_rmse"] = hist[False] == hist[True]
out["rmse"] = hist[True]
print(json.dumps(out))
