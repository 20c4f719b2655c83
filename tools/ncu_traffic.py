"""GEMM DRAM traffic per launch from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list of `bench.py --profile-step` -> profiles/r01_gemm_traffic.json (read by bench.py)."""
import csv
import json
import sys

path = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/r01_gemm_traffic.json"
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
per = {}
for r in rows[hi + 1:]:
    if len(r) <= vi or "gemm_bf16_tcgen05" not in r[ki]:
        continue
    d = per.setdefault(r[ii], {})
    d[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
n = len(per)
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in per.values())
wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in per.values())
t = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
res = {"launches_profiled": n, "dram_bytes_per_launch": (rd + wr) / max(1, n), "dram_read_bytes": rd,
       "dram_write_bytes": wr, "gemm_time_us_under_ncu": t,
       "note": f"dram__bytes_read.sum + dram__bytes_write.sum over the {n} st5 GEMM launches of the profiled updates "
               "(ncu, --clock-control none), divided by the launch count"}
json.dump(res, open(out, "w"), indent=1)
print(res)
