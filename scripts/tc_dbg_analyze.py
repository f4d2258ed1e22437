"""Analyse the (tag, clock64) event log gram_tc writes for CTA 0 when built with the diagnostic TcDbg hooks
(experimental branch only).  usage: python scripts/tc_dbg_analyze.py <dump> <stages>"""
import sys
import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(40, 4096, 2)
S = int(sys.argv[2])
def ev(w):
    a = raw[w]
    n = int((a[:, 0] != 0).sum())
    return a[:n, 0].astype(np.int64), a[:n, 1].astype(np.int64)
tag, clk = ev(4)
t5, t6 = clk[tag == 5], clk[tag == 6]
n = min(len(t5), len(t6))
print("MMA warp: stages logged", n, " period (wake->wake) mean %.0f  median %.0f clk;  wake->after issue mean %.0f" %
      (np.diff(t5[:n]).mean(), np.median(np.diff(t5[:n])), (t6[:n] - t5[:n]).mean()))
t9, t10 = clk[tag == 9], clk[tag == 10]
if len(t9) and len(t10):
    # align by following events: each 5 is followed by 9; 10 only in full stages
    i5 = np.nonzero(tag == 5)[0]
    seq = []
    for i in i5:
        j = i + 1
        rec = {5: clk[i]}
        while j < len(tag) and tag[j] != 5:
            rec[int(tag[j])] = clk[j]
            j += 1
        if all(k in rec for k in (9, 10, 6)):
            seq.append((rec[9] - rec[5], rec[10] - rec[9], rec[6] - rec[10]))
    seq = np.asarray(seq)
    print("MMA warp per full stage: wake->fence done %.0f, 4 MMAs issued %.0f, commit %.0f clk (each incl. ~20 clk of logging)" % tuple(seq.mean(0)))
t7, t8 = clk[tag == 7], clk[tag == 8]
if len(t7):
    m = min(len(t7), len(t8))
    print("MMA warp: items logged", m, " wait for a free accumulator mean %.0f clk" % (t8[:m] - t7[:m]).mean())
res = {k: [] for k in ("raw_handover", "split", "full_handover", "empty_handover", "loader_issue", "slot_cycle")}
for s in range(S):
    lt, lc = ev(5 + s)
    l1, l2 = lc[lt == 1], lc[lt == 2]
    sp = []
    for h in range(2):
        st, sc = ev(5 + S + 2 * s + h)
        sp.append((sc[st == 3], sc[st == 4]))
    uses = min(len(l1), len(l2), len(sp[0][0]), len(sp[0][1]), len(sp[1][0]), len(sp[1][1]))
    for u in range(2, uses - 1):
        stage = u * S + s
        if stage + S >= n:
            break
        res["loader_issue"].append(l2[u] - l1[u])
        res["raw_handover"].append(max(sp[0][0][u], sp[1][0][u]) - l2[u])
        res["split"].append(max(sp[0][1][u] - sp[0][0][u], sp[1][1][u] - sp[1][0][u]))
        res["full_handover"].append(t5[stage] - max(sp[0][1][u], sp[1][1][u]))
        res["empty_handover"].append(l1[u + 1] - t6[stage])
        res["slot_cycle"].append(l1[u + 1] - l1[u])
for k, v in res.items():
    v = np.asarray(v)
    if len(v):
        print("%-15s n=%5d  mean %7.0f  median %7.0f  p10 %7.0f  p90 %7.0f clk" % (k, len(v), v.mean(), np.median(v), np.percentile(v, 10), np.percentile(v, 90)))
