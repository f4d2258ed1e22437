"""Summarise an ncu --page source --csv export: top SASS lines by stall samples + stall-reason totals.
usage: python scripts/ncu_top.py <source.csv> [N] [section]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # n-th kernel section of the export
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[which]
end = his[which + 1] - 1 if which + 1 < len(his) else len(rows)
print(rows[hi - 1][:2])
hdr = rows[hi]
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
c = {h: i for i, h in enumerate(hdr)}
S = c["# Samples"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {h: sum(int(r[c[h]] or 0) for r in body) for h in stall_cols}
print("stall totals:", {k[6:]: "%.1f%%" % (100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[c[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print("%5d %6.2f%% exec=%-9s %-70s %s" % (i, 100.0 * int(r[S] or 0) / max(tot, 1), r[c["Instructions Executed"]], r[c["Source"]].strip()[:70],
                                            " ".join("%s:%d" % (h, v) for v, h in top if v)))
