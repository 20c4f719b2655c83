"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--launch-skip n --launch-count 1]`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.defaultdict(lambda: [0, "", collections.Counter(), 0])
hdr, cur_file = None, ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        ei = hdr.index("Instructions Executed")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and len(r) > si and r[si].strip().isdigit() and r[0].strip().isdigit():
        key = (cur_file, int(r[0]))
        a = agg[key]
        a[0] += int(r[si])
        a[1] = r[1].strip()[:100]
        a[3] += int(r[ei] or 0)
        for i in stall:
            if r[i]:
                a[2][hdr[i][6:]] += int(r[i])
tot = sum(a[0] for a in agg.values())
print("samples", tot)
for key, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = ", ".join(f"{k}:{v}" for k, v in a[2].most_common(2))
    print(f"{100 * a[0] / tot:5.1f}% {key[0]}:{key[1]:<4d} inst={a[3]:>8d} [{st}]  {a[1]}")
