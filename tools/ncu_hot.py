"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv [--launch-skip n --launch-count 1]`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si = hdr.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > si and r[si].strip().isdigit()]
tot = sum(int(r[si]) for r in body)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("kernel:", rows[0][1][:110], " samples", tot, " instrs", len(body))
agg = {}
for i in stall_cols:
    agg[hdr[i]] = sum(int(r[i] or 0) for r in body)
print("stall mix:", ", ".join(f"{k[6:]}={100 * v / max(1, tot):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:6d} {100 * int(r[si]) / tot:5.1f}%  {r[1][:90]:90s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
