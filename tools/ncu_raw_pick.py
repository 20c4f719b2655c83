"""Pick the headline metrics of every profiled launch out of `ncu -i X.ncu-rep --page raw --csv`: duration, DRAM bytes,
DRAM / L2 / tensor-pipe / issue utilisation, registers, achieved occupancy. usage: python tools/ncu_raw_pick.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
ki = hdr.index("Kernel Name")
units = rows[hi + 1]
print("kernel | " + " | ".join(f"{w} [{units[i]}]" for w, i in idx))
for r in rows[hi + 2:]:
    if len(r) <= ki:
        continue
    print(r[ki][:60] + " | " + " | ".join(r[i] for _, i in idx))
