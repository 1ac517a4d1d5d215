#!/bin/bash
# First GPU call of round 2 (one GPU, ~16 min): validates what was written at the end of round 1 without GPU access and
# fills the measurement gaps of the default workload (cfg3).  Run as
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/r2_first_gpu_call.sh'
# Everything lands in gpurun_out/r2_first/.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"

# 1. the whole GPU suite, with the gated GPU-vs-model build test enabled (tests/test_gpu_build.py)
LB200_UNVALIDATED=1 timeout 600 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"

# 2. the default bench line (cfg3: 10M x d768 cosine, batch 4096) incl. the prefix-graph cpu_baseline / parity block
( time timeout 900 python bench.py ) > "$OUT/bench_cfg3.json" 2> "$OUT/bench_cfg3.err"
echo "bench rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/bench_cfg3.err" | tee -a "$OUT/summary.txt"

# 3. DRAM traffic of the search kernel on cfg3 for roofline.traffic (profiles/search_kernel_traffic.json, key "cfg3"):
#    two launches after three warm-up launches; the 10M build runs un-profiled (ncu only replays the selected kernel)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 2 \
    -o "$OUT/search_cfg3" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/ncu_cfg3.log" 2>&1
echo "ncu rc=$?" | tee -a "$OUT/summary.txt"
ncu -i "$OUT/search_cfg3.ncu-rep" --page raw --csv > "$OUT/search_cfg3_raw.csv" 2>/dev/null
python - "$OUT" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1] + "/search_cfg3_raw.csv")))
if len(rows) > 2:
    head = rows[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "lts__t_sector_hit_rate.pct", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
    with open(sys.argv[1] + "/ncu_cfg3_summary.txt", "w") as f:
        for w in want:
            if w in head:
                i = head.index(w)
                f.write("%-60s %s %s\n" % (w, rows[1][i], "  ".join(r[i] for r in rows[2:])))
PY
cat "$OUT/ncu_cfg3_summary.txt" 2>/dev/null | tee -a "$OUT/summary.txt"
