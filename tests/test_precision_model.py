"""Numerical model of the tensor-core Gram (csrc/gram_tc.cuh) on the CPU: why the path is 3xTF32 and not plain TF32.

kind::tf32 reads only the upper 19 bits of an fp32 operand (sign, 8 exponent, 10 mantissa bits: the low 13 mantissa
bits are ignored) and accumulates in fp32.  The kernel splits y = h + l with h = those 19 bits and l = y - h (exact in
fp32) and accumulates H^T H + H^T L + L^T H (the L^T L term, ~2^-22 relative, is dropped).  This test replays that
arithmetic in numpy on ALS normal equations of the reference's shape (EmfWorker.js:231-247: A = Y^T Y + lambda*n*I,
b = Y^T r, x = A^-1 b) and checks it against fp64: the split keeps the solved factors within the 1e-3 bar of the
north star with two orders of magnitude to spare, single-pass TF32 does not."""
import numpy as np


def tf32_head(y):
    return (y.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def solve(A, b):
    return np.linalg.solve(A.astype(np.float64), b.astype(np.float64))


def rows(seed, n_rows=30, k=100, items=3000, rank=8, noise=0.03):
    """Fixed-side factors the way a few ALS iterations leave them: a dominant low-rank part (the synthetic ratings
    are rank-8 + noise) plus small residual directions — Y^T Y is then badly conditioned up to the
    lambda*n ridge, which is what amplifies operand rounding."""
    rng = np.random.default_rng(seed)
    V = (rng.normal(0, 1, (items, rank)) @ rng.normal(0, 0.6 / np.sqrt(rank), (rank, k))
         + rng.normal(0, noise, (items, k))).astype(np.float32)
    for _ in range(n_rows):
        n = int(rng.integers(100, 1500))
        Y = V[rng.choice(items, n, replace=False)]
        r = rng.integers(1, 11, n).astype(np.float32)
        yield Y, r, np.float32(0.05 * n)


def test_split_tf32_meets_the_bar_and_plain_tf32_does_not():
    worst3, worst1 = 0.0, 0.0
    for Y, r, lam in rows(0):
        k = Y.shape[1]
        Yr = np.concatenate([Y, r[:, None]], axis=1)                     # the ratings ride as one more operand column
        H = tf32_head(Yr)
        L = (Yr - H).astype(np.float32)                                  # exact: both share the exponent
        assert (H.astype(np.float64) + L.astype(np.float64) == Yr.astype(np.float64)).all()
        HH = (H.T.astype(np.float32) @ H).astype(np.float32)             # fp32 accumulation
        HL = (H.T.astype(np.float32) @ L).astype(np.float32)
        X3 = HH + HL + HL.T                                              # the kernel's D1 + 2*D2, symmetrised
        ref = Yr.astype(np.float64).T @ Yr.astype(np.float64)
        eye = np.eye(k)
        x_ref = solve(ref[:k, :k] + float(lam) * eye, ref[:k, k])
        x3 = solve(X3[:k, :k].astype(np.float64) + float(lam) * eye, X3[:k, k])
        x1 = solve(HH[:k, :k].astype(np.float64) + float(lam) * eye, HH[:k, k])
        worst3 = max(worst3, np.linalg.norm(x3 - x_ref) / np.linalg.norm(x_ref))
        worst1 = max(worst1, np.linalg.norm(x1 - x_ref) / np.linalg.norm(x_ref))
    assert worst3 < 1e-4, worst3            # measured on the GPU against fp64: 1.6e-5 worst row
    assert worst1 > 2e-3, worst1            # single-pass TF32 breaks the 1e-3 bar (≈5e-3 here): not an option


def test_dual_form_is_the_same_solution():
    """als_dual_kernel solves x = Y^T (Y Y^T + lambda*n I)^-1 r for rows with fewer ratings than factors —
    algebraically (Y^T Y + lambda*n I)^-1 Y^T r (EmfWorker.js:231-247), an n x n instead of a k x k system."""
    rng = np.random.default_rng(1)
    for n in (1, 7, 33, 96):
        Y = rng.normal(0, 0.4, (n, 100))
        r = rng.integers(1, 6, n).astype(np.float64)
        lam = 0.05 * n
        primal = np.linalg.solve(Y.T @ Y + lam * np.eye(100), Y.T @ r)
        dual = Y.T @ np.linalg.solve(Y @ Y.T + lam * np.eye(n), r)
        assert np.linalg.norm(primal - dual) <= 1e-9 * np.linalg.norm(primal)
