"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (optionally joined with gemm_shapes.json)."""
import collections
import csv
import json
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
shapes_path = sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mi = hdr.index("Metric Name") if "Metric Name" in hdr else None  # (lists with several metrics per launch: durations only)
agg = collections.defaultdict(lambda: [0, 0.0])
gemm = []
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) <= vi or (mi is not None and r[mi] != "gpu__time_duration.sum"):
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
    name = re.sub(r"\(.*", "", r[ki])[:64]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
    if "gemm_bf16_tcgen05" in r[ki]:
        gemm.append((v, re.search(r"<(.*?)>", r[ki]).group(1)))
print(f"total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f}% n={n:4d} avg={t / n:7.1f}  {k}")
if shapes_path:
    shapes = json.load(open(shapes_path))
    assert len(gemm) % len(shapes) == 0, (len(shapes), len(gemm))
    print(f"({len(gemm) // len(shapes)} recorded steps; GEMM table = the last one)")
    gemm = gemm[-len(shapes):]
    a2 = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (t, tmpl), s in zip(gemm, shapes):
        key = (s["M"], s["N"], s["K"], s["nb"], s["a_mn"], s["b_mn"], s["c_fp32"], s["acc"], s.get("act", 0),
               s.get("drop", False), s.get("ag", False), tmpl)
        a2[key][0] += 1
        a2[key][1] += t
        a2[key][2] += 2.0 * s["M"] * s["N"] * s["K"] * s["nb"]
    gt = sum(a[1] for a in a2.values())
    print(f"GEMM total {gt:.0f} us")
    for k, (n, t, f) in sorted(a2.items(), key=lambda x: -x[1][1])[:30]:
        print(f"{t:8.1f}us n={n:3d} avg={t / n:7.1f} {f / t / 1e6:7.1f}TF/s M={k[0]} N={k[1]} K={k[2]} nb={k[3]} "
              f"mn={k[4]}{k[5]} f32={k[6]} acc={k[7]} act={k[8]} drop={int(k[9])} ag={int(k[10])} <{k[11]}>")
