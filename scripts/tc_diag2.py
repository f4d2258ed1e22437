"""Diagnostic: dump the tile partials of the tcgen05 Gram and compare with numpy (GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests.helpers import portion_from_rows
from you_can_not_recommend_b200 import native

np.set_printoptions(linewidth=200, precision=4, suppress=True)
rng = np.random.default_rng(1)
k, n_fixed = 100, 3000
KT = 25
F = rng.normal(0, 0.3, (n_fixed, k)).astype(np.float32)
lens = [128, 200]
cols = [np.sort(rng.choice(n_fixed, n, replace=False)) for n in lens]
vals = [rng.integers(1, 11, n).astype(np.float32) for n in lens]
rows, indx, v = portion_from_rows([0, 1], [c.tolist() for c in cols], [x.tolist() for x in vals])
ntiles = KT * (KT + 1) // 2 + KT


def tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def expect(r):
    Y = np.zeros((lens[r], 128), np.float32)
    Y[:, :k] = F[cols[r]]
    Y[:, k] = vals[r]
    H = tf32(Y)
    L2 = tf32((2 * (Y - H)).astype(np.float32))
    H64, L64 = H.astype(np.float64), L2.astype(np.float64)
    return H64.T @ H64, H64.T @ L64, Y.astype(np.float64).T @ Y.astype(np.float64)


def untile(p):
    X = np.full((104, 104), np.nan)
    t = 0
    for I in range(KT):
        for L in range(I + 1):
            X[4 * I:4 * I + 4, 4 * L:4 * L + 4] = p[t]
            t += 1
    for L in range(KT):
        X[100:104, 4 * L:4 * L + 4] = p[t]
        t += 1
    return X


def run(**kw):
    S = np.zeros((2, k), np.float32)
    ctx = native.Context(k, 2, n_fixed, 0.05, 0.05, **kw)
    ctx.attach_factors(S, F)
    ctx.start_train_step(native.BY_USER)
    ctx.als_portion(rows, indx, v)
    ctx.end_train_step()
    p = ctx.debug_read_partials(2, ntiles)
    ctx.close()
    return p, S


pf, Sf = run(gram_path=native.GRAM_FFMA, split_cols=32 * 1, dual_max_cols=0)
print("ffma partial items per row:", "n/a (split 32)")
for variant, label in ((0, "sym"), (2, "raw H^T H"), (4, "raw H^T 2L")):
    for lbo_swap in (0, 1):
        try:
            pt, St = run(gram_path=native.GRAM_TC3XTF32, tc_variant=variant | lbo_swap, dual_max_cols=0)
        except Exception as ex:
            print(label, lbo_swap, "FAILED", ex)
            continue
        for r in range(2):
            HH, HL, YY = expect(r)
            X = untile(pt[r])
            want = {0: YY, 2: HH, 4: HL}[variant][:104, :104]
            mask = ~np.isnan(X)
            if variant == 0:
                pass
            diff = np.abs(np.where(mask, X - want, 0))
            scale = np.abs(want).max()
            print("%-12s swap=%d row=%d  max|X|=%.4g  max|want|=%.4g  maxdiff/scale=%.3g  nonzero frac=%.3f" %
                  (label, lbo_swap, r, np.nanmax(np.abs(X)), scale, diff.max() / scale, float((np.nan_to_num(X) != 0).mean())), flush=True)
            if r == 0 and diff.max() / scale > 1e-3 and variant == 2 and lbo_swap == 0:
                print(" got  X[0:8,0:8]:\n", X[0:8, 0:8])
                print(" want X[0:8,0:8]:\n", want[0:8, 0:8])
                print(" got  X[96:104,0:8]:\n", X[96:104, 0:8])
                print(" want X[96:104,0:8]:\n", want[96:104, 0:8])
                print(" got  X[32:40,28:36]:\n", X[32:40, 28:36])
                print(" want X[32:40,28:36]:\n", want[32:40, 28:36])
