"""Per-instruction shared-memory wavefronts of a kernel from an `ncu --page source --csv` export: every LDS / STS /
LDGSTS / UTMALDG line with its executed count, L1 shared wavefronts, ideal wavefronts and conflicts.
usage: python scripts/ncu_smem_lines.py <source.csv> [section]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[his[sec]]
c = {h: i for i, h in enumerate(hdr)}
end = his[sec + 1] - 1 if sec + 1 < len(his) else len(rows)
body = [r for r in rows[his[sec] + 1:end] if len(r) == len(hdr)]
cols = [h for h in hdr if "Wavefront" in h or "Conflict" in h or "wavefront" in h]
print("columns:", cols)
tot = {h: 0 for h in cols}
for r in body:
    src = r[c["Source"]].strip()
    if not any(op in src for op in ("LDS", "STS", "LDGSTS", "UTMALDG", "LDSM", "STSM")):
        continue
    vals = []
    for h in cols:
        try:
            v = int(float(r[c[h]] or 0))
        except ValueError:
            v = 0
        tot[h] += v
        vals.append(v)
    if any(vals):
        print("%-70s exec=%-10s %s" % (src[:70], r[c["Instructions Executed"]], vals))
print("totals:", tot)
