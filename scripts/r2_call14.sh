#!/bin/bash
# Round-2 GPU call 14 (two GPUs): the group kernel after limiting the L2 prefetch to short lists: 1 and 2 ranks, 1 M rows.
set -u
OUT=gpurun_out/r2_call14
mkdir -p "$OUT"
timeout 400 python scripts/exp_group.py 2 1000000 cos 4096 > "$OUT/exp_group.log" 2>&1
echo "rc=$?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/exp_group.log" | tee -a "$OUT/summary.txt"
timeout 300 python -m pytest tests/test_gpu_group.py -q > "$OUT/pytest_group.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest_group.log" | tee -a "$OUT/summary.txt"
