"""Top-N serving throughput: ycnr_recommend_batch on the MAL shape, random factors.
  python scripts/recommend_bench.py [users_in_batch] [limit]  -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from you_can_not_recommend_b200 import native  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 20
users, items, k = 1_750_000, 12_700, 100
rng = np.random.default_rng(0)
U = rng.normal(0.3, 0.2, (users, k)).astype(np.float32)
V = rng.normal(0.3, 0.2, (items, k)).astype(np.float32)
ctx = native.Context(k, users, items)
ctx.attach_factors(U, V)
uids = rng.choice(users, batch, replace=False).astype(np.int32)
skips = [rng.choice(items, 66, replace=False).astype(np.int32) for _ in range(batch)]     # a MAL user's list
ctx.recommend_batch(uids[:64], skips[:64], limit, 0.0, 0.1)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    out = ctx.recommend_batch(uids, skips, limit, 0.0, 0.1)
    best = min(best, time.perf_counter() - t0)
pairs = batch * items
print(json.dumps({"metric": "recommend_users_per_sec", "value": batch / best, "batch": batch, "limit": limit,
                  "items": items, "factors": k, "ms": best * 1e3, "user_item_pairs_per_sec": pairs / best,
                  "algorithmic_gbs": pairs * k * 4 / best / 1e9,
                  "note": "wall time of ycnr_recommend_batch incl. H2D of ids/skip lists and D2H of the lists; V (5 MB) is L2-resident, "
                          "so the algorithmic bytes (items x k x 4 per user) are L2 traffic, not HBM"}))
ctx.close()
