"""DRAM traffic per launch and per kernel class from `ncu --set full` reports (raw page):
usage: python scripts/ncu_traffic.py out.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, json, subprocess, sys
CLASS = [("gram_tc_kernel", "gram_tc"), ("als_dual", "dual_fused"), ("rmse_rows_kernel", "rmse_rows"),
         ("rmse_portion_reduce", "rmse_reduce"), ("als_primal_kernel", "reduce_solve"), ("als_solve_blocks", "reduce_solve")]
agg = {}
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    def val(r, name):
        i = h.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        cls = next((c for k, c in CLASS if k in name), None)
        if cls is None:
            continue
        a = agg.setdefault(cls, {"dram_bytes": 0.0, "launches": 0, "ncu_ms_total": 0.0})
        a["dram_bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        a["launches"] += 1
        a["ncu_ms_total"] += val(r, "gpu__time_duration.sum")
res = {c: {"dram_bytes_per_launch": a["dram_bytes"] / a["launches"], "launches": a["launches"], "ncu_ms_total": a["ncu_ms_total"]}
       for c, a in agg.items()}
res["_source"] = "ncu --set full --clock-control none, one timed MAL iteration; workload mal k=100, 1 GPU"
json.dump(res, open(sys.argv[1], "w"), indent=1)
print(json.dumps(res))
