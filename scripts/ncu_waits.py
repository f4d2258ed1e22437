"""Warp-role view of an `ncu --page source --csv` export of gram_tc_kernel: stall samples aggregated by the
mbarrier each try_wait spins on (empty / raw / full / accumulator-full), plus the instruction mix.
usage: python scripts/ncu_waits.py <source.csv> [section] > profiles/<name>.txt"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[his[sec]]
c = {h: i for i, h in enumerate(hdr)}
end = his[sec + 1] - 1 if sec + 1 < len(his) else len(rows)
body = [r for r in rows[his[sec] + 1:end] if len(r) == len(hdr)]
tot = sum(int(r[c["# Samples"]] or 0) for r in body)
print(rows[his[sec] - 1][:2])
print("total samples", tot)
agg = {}
for i, r in enumerate(body):
    m = re.search(r"TRYWAIT P\d, \[R\d+\+URZ\+(0x[0-9a-f]+)\]", r[c["Source"]])
    if m and i + 1 < len(body):
        s = int(r[c["# Samples"]] or 0) + int(body[i + 1][c["# Samples"]] or 0)
        a = agg.setdefault(m.group(1), [0, 0])
        a[0] += s
        a[1] += int(r[c["Instructions Executed"]] or 0)
print("samples spent spinning on each mbarrier array (smem offset: samples, try_wait executions, share):")
for k, v in sorted(agg.items()):
    print("  %s  %9d  %12d  %5.1f%%" % (k, v[0], v[1], 100.0 * v[0] / max(tot, 1)))
cls = {}
for r in body:
    op = r[c["Source"]].strip().split()
    if not op:
        continue
    o = op[0] if not op[0].startswith("@") else (op[1] if len(op) > 1 else op[0])
    o = o.split(".")[0]
    a = cls.setdefault(o, [0, 0])
    a[0] += int(r[c["# Samples"]] or 0)
    a[1] += int(r[c["Instructions Executed"]] or 0)
print("opcode: samples share, warp-instructions")
for k, v in sorted(cls.items(), key=lambda kv: -kv[1][0])[:16]:
    print("  %-10s %5.1f%%  %12d" % (k, 100.0 * v[0] / max(tot, 1), v[1]))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {h[6:]: sum(int(r[c[h]] or 0) for r in body) for h in stall_cols}
print("stall reasons:", {k: "%.1f%%" % (100.0 * v / max(tot, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]})
