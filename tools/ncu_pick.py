"""Pick metrics out of `ncu -i X.ncu-rep --page raw --csv`:  python tools/ncu_pick.py raw.csv [substr ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "pipe_fp64", "pipe_tensor",
                        "warps_active.avg.pct", "registers_per_thread", "issue_active.avg.pct", "dram_throughput", "bank_conflicts",
                        "warp_issue_stalled", "occupancy", "inst_executed.sum", "lts__t_sector_hit"]
for r in rows[2:]:
    print("==", r[4][:60])
    for h, u, v in zip(hdr, units, r):
        if any(w in h for w in want):
            print(f"  {h} [{u}] = {v}")
