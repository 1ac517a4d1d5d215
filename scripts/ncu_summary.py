"""ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python scripts/ncu_summary.py X_raw.csv [more.csv ...] > profiles/summary.txt"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def num(v):
    return float(v.replace(",", ""))


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== %s: %s" % (path.split("/")[-1], vals[head.index("Kernel Name")]))
        for w in WANT:
            if w in head:
                i = head.index(w)
                print("  %-66s %s %s" % (w, vals[i], units[i]))
        st = [(h, vals[i]) for i, h in enumerate(head) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
        st = sorted([(num(v), h) for h, v in st if v not in ("", "n/a")], reverse=True)[:8]
        for v, h in st:
            print("   stall samples %-50s %d" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), v))
