"""Product front end (libycnr_host.so) vs the literal Python restatement in oracle/front_end.py:
bit-exact splits, portion plans and portion buffers (north star: "bit-exact ratings indexing and splits")."""
import json
import os

import numpy as np
import pytest

from oracle import front_end as ofe
from you_can_not_recommend_b200 import front_end as fe

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def small_table(seed=7, users=60, items=40, ratings=700):
    return fe.synth_table("custom", seed=seed, users=users, items=items, ratings=ratings, max_rating=5)


def test_prng_matches_oracle():
    L = fe.lib()
    for seed, a, b in [(0, 0, 0), (20261017, 5, 9), (2**63 + 11, 2**40, 3), (1, 2**64 - 1, 7)]:
        h = L.ycnr_mix64(seed, a, b)
        assert h == ofe.mix64(seed, a, b)
        assert L.ycnr_u01(h) == ofe.u01(h)


def test_synth_table_is_well_formed():
    t = small_table()
    assert t.nnz == 700 and t.user_ptr[0] == 0
    for u in range(t.users):
        ids = t.item_ids[t.user_ptr[u]:t.user_ptr[u + 1]]
        assert len(ids) >= 1 and (np.diff(ids) > 0).all()          # unique + ascending (ORDER BY user, item)
        assert ids.min() >= 0 and ids.max() < t.items
    assert set(np.unique(t.ratings)).issubset({1.0, 2.0, 3.0, 4.0, 5.0})
    t2 = small_table()
    assert (t.item_ids == t2.item_ids).all() and (t.ratings == t2.ratings).all()   # deterministic


@pytest.mark.parametrize("pcts", [(85, 10, 5), (70, 20, 10), (100, 0, 0), (50, 50, 0)])
def test_split_sets_bit_exact(pcts):
    t = small_table(seed=11, users=80, items=50, ratings=1500)
    fe.split_sets(t, pcts, seed=123)
    want = ofe.split_sets(123, t.user_ptr, pcts)
    assert (t.dataset_type == want).all()
    n = np.diff(t.user_ptr)
    # rule Q9 counts per user
    for u in range(t.users):
        d = t.dataset_type[t.user_ptr[u]:t.user_ptr[u + 1]]
        t0 = int(np.ceil(n[u] * pcts[0] / 100))
        assert (d == 1).sum() == t0 or sum(pcts[1:]) == 0


def test_split_sets_single_rating_users():
    t = fe.table_from_triples(5, 3, [0, 1, 2, 3, 4], [0, 1, 2, 0, 1], [1, 2, 3, 4, 5])
    fe.split_sets(t, (85, 10, 5), seed=5)
    assert (t.dataset_type == 1).all()          # ceil(0.85) = 1 -> train
    assert (t.dataset_type == ofe.split_sets(5, t.user_ptr, (85, 10, 5))).all()


@pytest.mark.parametrize("rip,nthr,pct", [(100, 1, 0), (37, 1, 0), (10, 4, 0), (10000, 3, 0), (50, 1, 11), (50, 2, 6), (1, 1, 0)])
def test_split_to_portions_matches_oracle(rip, nthr, pct):
    rng = np.random.default_rng(3)
    cnt = rng.integers(0, 40, 200).astype(np.int32)
    cnt[rng.integers(0, 200, 30)] = 0               # id holes (Q5)
    got, mr, mrows = fe.split_to_portions(cnt, rip, nthr, pct)
    want, wmr, wmrows = ofe.split_to_portions(cnt, rip, nthr, pct)
    assert list(got) == want and mr == wmr and mrows == wmrows
    assert (np.diff(got) > 0).all()


def test_q2_golden_cases():
    """Hand-simulated outputs of the upstream conversion loop."""
    cases = json.load(open(os.path.join(GOLD, "q2_cases.json")))
    for case in cases:
        letters = case["rows"]
        names = sorted(set(letters))
        rid = {c: i for i, c in enumerate(names)}
        data = [{"r": rid[c] + 1, "c": j + 1, "rating": 1.0 + j} for j, c in enumerate(letters)]
        # oracle literal
        rows, indx, vals = ofe.convert_portion(data, 8, 16)
        got = [[names[rows[1 + 2 * r]], int(rows[2 + 2 * r])] for r in range(rows[0])]
        assert got == case["expect"], (letters, got)
        # product
        if letters:
            u = [rid[c] for c in letters]
            t = fe.table_from_triples(len(names), len(letters), u, list(range(len(letters))), [1.0 + j for j in range(len(letters))])
            t.dataset_type[:] = 1
            csr = t.csr_by_user(fe.MASK_TRAIN)
            prow, pindx, pvals, fetched = fe.build_portion(csr, 0, len(names), 8, 16)
            assert fetched == len(letters)
            assert (prow[:2 * rows[0] + 1] == rows[:2 * rows[0] + 1]).all()
            assert (pindx[:fetched] == indx[:fetched]).all() and (pvals[:fetched] == vals[:fetched]).all()


@pytest.mark.parametrize("step", ["byUser", "byItem", "rmseValidate", "rmseTest"])
def test_portions_bit_exact_vs_oracle(step):
    t = small_table(seed=21, users=70, items=45, ratings=1200)
    fe.split_sets(t, (85, 10, 5), seed=9)
    cnt = t.counts_per_item() if step == "byItem" else t.counts_per_user()
    pct = {"rmseValidate": 11, "rmseTest": 6}.get(step, 0)
    pto, mr, mrows = fe.split_to_portions(cnt, 150, 1, pct)
    mask = {"rmseValidate": fe.MASK_VALIDATE, "rmseTest": fe.MASK_TEST}.get(step, fe.MASK_TRAIN)
    csr = t.csr_by_item(mask) if step == "byItem" else t.csr_by_user(mask)
    rl = fe.build_rowlist(csr, pto)
    r = 0
    for p in range(len(pto)):
        row_from = 0 if p == 0 else int(pto[p - 1])
        data = ofe.fetch(t.user_ptr, t.item_ids, t.ratings, t.dataset_type, step, row_from, int(pto[p]))
        want = ofe.convert_portion(data, mrows + 1, mr)
        got = fe.build_portion(csr, row_from, int(pto[p]), mrows + 1, mr)
        assert got[3] == len(data) <= mr                     # planner keeps portions inside the buffers
        R = want[0][0]
        assert (got[0][:2 * R + 1] == want[0][:2 * R + 1]).all()
        assert (got[1][:len(data)] == want[1][:len(data)]).all()
        assert (got[2][:len(data)] == want[2][:len(data)]).all()
        # bulk row list = concatenated headers, addressing the same ratings
        assert rl.portion_first[p] == r
        off = 0
        for q in range(R):
            assert rl.row_ids[r] == want[0][1 + 2 * q] and rl.row_len[r] == want[0][2 + 2 * q]
            seg = slice(int(rl.row_start[r]), int(rl.row_start[r]) + int(rl.row_len[r]))
            assert (csr.idx[seg] == want[1][off:off + rl.row_len[r]]).all()
            off += int(rl.row_len[r])
            r += 1
    assert rl.portion_first[-1] == r == len(rl.row_ids)
    # Q2: exactly one rating dropped per non-empty portion
    nonempty = sum(1 for p in range(len(pto)) if csr.ptr[pto[p]] > csr.ptr[0 if p == 0 else pto[p - 1]])
    assert csr.nnz - rl.nnz == nonempty


def test_stats_counts():
    t = small_table(seed=4)
    fe.split_sets(t, (85, 10, 5), seed=1)
    assert (t.counts_per_user() == np.diff(t.user_ptr)).all()
    assert (t.counts_per_item() == np.bincount(t.item_ids, minlength=t.items)).all()
    assert abs(t.total_ratings_avg() - t.ratings.astype(np.float64).mean()) < 1e-12


def test_build_portion_overflow_is_an_error():
    t = small_table(seed=4)
    t.dataset_type[:] = 1
    csr = t.csr_by_user(fe.MASK_TRAIN)
    with pytest.raises(RuntimeError, match="exceed buffer"):
        fe.build_portion(csr, 0, t.users, t.users, 10)


def test_q2_closed_form_equals_conversion_loop():
    """The device ingest (csrc/ingest_kernels.cuh) does not replay the master's per-rating loop
    (EmfMaster.js:582-609); it uses its closed form: inside every portion all non-empty rows are emitted with
    their full count except the LAST non-empty one, which is emitted one short — and not at all when that
    made it empty, unless it is the portion's only row.  Checked against the loop restatement on random plans."""
    import numpy as np
    from you_can_not_recommend_b200 import front_end as fe
    rng = np.random.default_rng(0)
    for _ in range(300):
        rows = int(rng.integers(1, 30))
        cnt = rng.integers(0, 4, rows)
        ptr = np.zeros(rows + 1, np.int64)
        np.cumsum(cnt, out=ptr[1:])
        cuts = np.unique(np.concatenate([rng.integers(1, rows + 1, int(rng.integers(1, 6))), [rows]])).astype(np.int32)
        csr = fe.Csr(ptr, np.zeros(int(ptr[-1]), np.int32), np.zeros(int(ptr[-1]), np.float32))
        rl = fe.build_rowlist(csr, cuts)
        ids, lens, pf = [], [], [0]
        for p in range(len(cuts)):
            lo, hi = (0 if p == 0 else int(cuts[p - 1])), int(cuts[p])
            ne = [r for r in range(lo, hi) if cnt[r] > 0]
            for r in ne:
                if r == ne[-1]:
                    if cnt[r] == 1 and len(ne) > 1:
                        continue
                    ids.append(r)
                    lens.append(int(cnt[r]) - 1)
                else:
                    ids.append(r)
                    lens.append(int(cnt[r]))
            pf.append(len(ids))
        assert list(rl.row_ids) == ids and list(rl.row_len) == lens and list(rl.portion_first) == pf
        assert list(rl.row_start) == [int(ptr[r]) for r in ids]
