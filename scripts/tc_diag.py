"""Diagnostic: tcgen05 Gram path vs FFMA path vs float64 on adversarial rows (run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests.helpers import portion_from_rows, rel_fro, worst_row_rel
from you_can_not_recommend_b200 import native

rng = np.random.default_rng(1)
k, n_fixed = 100, 6000
F = rng.normal(0, 0.3, (n_fixed, k)).astype(np.float32)
lens = [97, 100, 128, 129, 200, 333, 1000, 4096, 4097, 5999]
ids = list(range(len(lens)))
cols = [np.sort(rng.choice(n_fixed, n, replace=False)).tolist() for n in lens]
vals = [rng.integers(1, 11, n).astype(float).tolist() for n in lens]
rows, indx, v = portion_from_rows(ids, cols, vals)
ref = np.zeros((len(lens), k))
for r, n in enumerate(lens):
    Y = F[cols[r]].astype(np.float64)
    ref[r] = np.linalg.solve(Y.T @ Y + 0.05 * n * np.eye(k), Y.T @ np.asarray(vals[r]))
for name, kw in (("ffma", dict(gram_path=native.GRAM_FFMA)),
                 ("tc v0", dict(gram_path=native.GRAM_TC3XTF32, tc_variant=0)),
                 ("tc raw-head", dict(gram_path=native.GRAM_TC3XTF32, tc_variant=16))):
    S = np.zeros((len(lens), k), np.float32)
    try:
        ctx = native.Context(k, len(lens), n_fixed, 0.05, 0.05, profile=True, **kw)
        ctx.attach_factors(S, F)
        ctx.start_train_step(native.BY_USER)
        ctx.als_portion(rows, indx, v)
        ctx.end_train_step()
        prof = ctx.profile_read()
        err = [float(np.linalg.norm(S[r] - ref[r]) / np.linalg.norm(ref[r])) for r in range(len(lens))]
        print(name, "worst", max(err), ["%.1e" % e for e in err], {c: prof[c]["launches"] for c in ("gram_tc", "gram_partial", "primal_fused", "reduce_solve")}, flush=True)
        ctx.close()
    except Exception as ex:
        print(name, "FAILED", ex, flush=True)
