"""Generates the committed fixtures in this directory.

Upstream ships no golden vectors (package.json:30), and node is not installed, so the
fixtures are minted here, from two independent sources:
  * tiny_als.json  — a 3-user x 4-item, k=2 problem whose normal equations are solved in
                     exact rational arithmetic (fractions.Fraction), not by the oracle;
  * q2_cases.json  — hand-simulated outputs of the portion-conversion loop
                     (EmfMaster.js:582-609) for the patterns of SURVEY.md Q2;
  * c1_trajectory.json — frozen O64 RMSE trajectory of the ML-100k-shaped config
                     (regression pin of the oracle itself).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys
from fractions import Fraction

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def exact_solve(A, b):
    n = len(b)
    M = [[Fraction(x) for x in row] + [Fraction(y)] for row, y in zip(A, b)]
    for c in range(n):
        p = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        for r in range(n):
            if r != c:
                f = M[r][c] / M[c][c]
                M[r] = [x - f * y for x, y in zip(M[r], M[c])]
    return [M[i][n] / M[i][i] for i in range(n)]


def tiny_als():
    # item factors (4 x 2) with exactly representable binary fractions, lambda = 1/16
    V = [[0.5, -0.25], [1.0, 0.75], [-0.5, 0.125], [0.25, 1.5]]
    lam = Fraction(1, 16)
    users = {0: [(0, 4.0), (1, 5.0), (3, 2.0)], 1: [(2, 1.0)], 2: [(0, 3.0), (1, 3.0), (2, 4.0), (3, 5.0)]}
    out = {"k": 2, "lambda": float(lam), "V": V, "users": {}, "rows": []}
    for u, rs in users.items():
        n = len(rs)
        A = [[sum(Fraction(V[i][a]) * Fraction(V[i][b]) for i, _ in rs) + (lam * n if a == b else 0)
              for b in range(2)] for a in range(2)]
        bb = [sum(Fraction(V[i][a]) * Fraction(r) for i, r in rs) for a in range(2)]
        x = exact_solve(A, bb)
        out["users"][str(u)] = {"ratings": rs, "x": [float(v) for v in x],
                                "x_frac": ["%d/%d" % (v.numerator, v.denominator) for v in x]}
    return out


def q2_cases():
    # rows given as letters; expected = list of [rowLetter, cols] exactly as the upstream loop emits
    return [
        {"rows": "AABB", "expect": [["A", 2], ["B", 1]]},
        {"rows": "AAB", "expect": [["A", 2]]},
        {"rows": "A", "expect": [["A", 0]]},
        {"rows": "AB", "expect": [["A", 1]]},
        {"rows": "AAAA", "expect": [["A", 3]]},
        {"rows": "ABBC", "expect": [["A", 1], ["B", 2]]},
        {"rows": "ABCD", "expect": [["A", 1], ["B", 1], ["C", 1]]},
        {"rows": "", "expect": []},
    ]


def c1_trajectory():
    from oracle import oracle
    from tests.helpers import make_problem, oracle_portions
    prob = make_problem("ml-100k", k=20)
    tr = oracle.OracleTrainer(prob["U0"], prob["V0"], oracle_portions(prob), 0.05, 0.05,
                              prob["total_ratings_avg"], dtype=np.float64)
    hist = tr.train(10)
    return {"shape": "ml-100k", "k": 20, "seed": prob["seed"], "history": hist,
            "U_checksum": float(np.abs(tr.U).sum()), "V_checksum": float(np.abs(tr.V).sum())}


if __name__ == "__main__":
    for name, fn in (("tiny_als", tiny_als), ("q2_cases", q2_cases), ("c1_trajectory", c1_trajectory)):
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
