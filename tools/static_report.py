"""Static evidence for profiles/: per-kernel registers / spills / static smem from `ptxas -v` (build.build(verbose=True)
output on stdin or a file) and the count of Blackwell mnemonics (tcgen05 = UTC*, TMA = UBLKCP / UTMALDG / UTMASTG) per
kernel from `cuobjdump -sass` of the built library. No GPU needed.
usage: python tools/static_report.py /tmp/ptxas_v.txt > profiles/rNN_static.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speecht5_b200", "lib",
                   "libspeecht5_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(sig):
    sig = re.sub(r"^void ", "", sig)
    depth, cut = 0, len(sig)
    for i, ch in enumerate(sig):  # drop the parameter list, keep template arguments
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return sig[:cut].replace("(anonymous namespace)::", "")


def ptxas(path):
    rows, cur = {}, None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            cur = m.group(1)
            rows[cur] = {"regs": None, "spill_st": 0, "spill_ld": 0, "stack": 0, "smem": 0}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            rows[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            rows[cur]["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            rows[cur]["smem"] = int(s.group(1)) if s else 0
    return rows


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.defaultdict(collections.Counter), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base.startswith("UTC") or base in ("UBLKCP", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "SYNCS",
                                               "UTMACCTL", "UTMACMDFLUSH"):
            counts[cur][base] += 1
        counts[cur]["_total"] += 1
    return counts


def main():
    rows = ptxas(sys.argv[1])
    counts = sass()
    names = demangle(sorted(set(rows) | set(counts)))
    print("# static report: ptxas -v (sm_100a) + cuobjdump -sass mnemonic counts per kernel")
    print(f"{'kernel':<78} {'regs':>4} {'stack':>5} {'sp_st':>5} {'sp_ld':>5} {'smem':>6} {'sass':>6}  blackwell mnemonics")
    for mangled in sorted(rows, key=lambda k: short(names[k])):
        r, c = rows[mangled], counts.get(mangled, {})
        bw = " ".join(f"{k}:{v}" for k, v in sorted(c.items()) if not k.startswith("_"))
        print(f"{short(names[mangled])[:78]:<78} {r['regs']:>4} {r['stack']:>5} {r['spill_st']:>5} {r['spill_ld']:>5} "
              f"{r['smem']:>6} {c.get('_total', 0):>6}  {bw}")


if __name__ == "__main__":
    main()
