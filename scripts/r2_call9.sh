#!/bin/bash
# Round-2 GPU call 9 (one GPU): tcgen05 exhaustive search after the heap epilogue; ncu of the re-designed PQ search kernel.
set -u
OUT=gpurun_out/r2_call9
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_exact_tc.py tests/test_gpu_golden.py "tests/test_gpu_group.py::test_group_wide_lists" -q -s > "$OUT/pytest.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
LB200_EXACT_REPORT=1 timeout 600 python scripts/exp_exact.py 2000000 1024 768 > "$OUT/exp_exact.log" 2>&1
echo "exp exact rc=$?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/exp_exact.log" | tee -a "$OUT/summary.txt"
LB200_EXACT=tc timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_tc_filter -c 1 -o "$OUT/exact_tc" -f python scripts/exp_exact.py 500000 1024 768 > "$OUT/ncu_exact.log" 2>&1
echo "ncu exact rc=$?" | tee -a "$OUT/summary.txt"
ncu -i "$OUT/exact_tc.ncu-rep" --page raw --csv > "$OUT/exact_tc_raw.csv" 2>/dev/null
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o "$OUT/search_cfg4s" -f \
    python bench.py --workload cfg4s --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/ncu_cfg4s.log" 2>&1
echo "ncu cfg4s rc=$?" | tee -a "$OUT/summary.txt"
ncu -i "$OUT/search_cfg4s.ncu-rep" --page raw --csv > "$OUT/search_cfg4s_raw.csv" 2>/dev/null
timeout 900 python bench.py --workload cfg4s --steps 50 --warmup 5 --no-cpu-baseline --search-expand 2 > "$OUT/bench_cfg4s_expand2.json" 2> "$OUT/bench_cfg4s_expand2.err"
python -c "
import json
d=json.loads(open('$OUT/bench_cfg4s_expand2.json').read().strip().splitlines()[-1]); print('cfg4s expand 2:', round(d['value']), 'q/s kernel ms', d['roofline']['kernel_ms_per_step'], 'recall@k', d['recall_at_k'])" | tee -a "$OUT/summary.txt"
