#!/bin/bash
# Round-2 GPU call 2 (one GPU): the row-sharded group's protocol with several ranks on one device, the whole GPU suite.
set -u
OUT=gpurun_out/r2_call2
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_group.py -q -x -s > "$OUT/pytest_group.log" 2>&1
echo "pytest group rc=$?" | tee -a "$OUT/summary.txt"; tail -15 "$OUT/pytest_group.log" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_group.py > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rest rc=$?" | tee -a "$OUT/summary.txt"; tail -8 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
timeout 600 python scripts/exp_group.py 2 200000 cos 4096 > "$OUT/exp_group_cos.log" 2>&1
echo "exp rc=$?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/exp_group_cos.log" | tee -a "$OUT/summary.txt"
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > "$OUT/sanitizer_racecheck.log" 2>&1
echo "racecheck rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/sanitizer_racecheck.log" | tee -a "$OUT/summary.txt"
