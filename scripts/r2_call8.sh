#!/bin/bash
# Round-2 GPU call 8 (one GPU): group tests on one device after the allocation-order fix, exhaustive search timing + ncu.
set -u
OUT=gpurun_out/r2_call8
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_exact_tc.py tests/test_gpu_search.py -q -s > "$OUT/pytest.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
timeout 600 python scripts/exp_exact.py 2000000 1024 768 > "$OUT/exp_exact.log" 2>&1
echo "exp exact rc=$?" | tee -a "$OUT/summary.txt"; cat "$OUT/exp_exact.log" | tail -4 | tee -a "$OUT/summary.txt"
LB200_EXACT=tc timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_tc_filter -c 1 -o "$OUT/exact_tc" -f python scripts/exp_exact.py 500000 1024 768 > "$OUT/ncu_exact.log" 2>&1
echo "ncu rc=$?" | tee -a "$OUT/summary.txt"
ncu -i "$OUT/exact_tc.ncu-rep" --page raw --csv > "$OUT/exact_tc_raw.csv" 2>/dev/null
timeout 600 python scripts/exp_group.py 4 300000 cos 2048 > "$OUT/exp_group_4ranks_1dev.log" 2>&1
echo "exp group rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/exp_group_4ranks_1dev.log" | tee -a "$OUT/summary.txt"
