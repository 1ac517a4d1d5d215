#!/bin/bash
# Round-2 GPU call 16 (one GPU): the batching daemon serving from a row-sharded group (--devices).
set -u
OUT=gpurun_out/r2_call16
mkdir -p "$OUT"
timeout 150 python -m pytest tests/test_gpu_daemon.py -q > "$OUT/pytest.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
