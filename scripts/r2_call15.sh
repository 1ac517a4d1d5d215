#!/bin/bash
# Round-2 GPU call 15 (one GPU): last sanity pass of the final tree (group tests, smoke).
set -u
OUT=gpurun_out/r2_call15
mkdir -p "$OUT"
timeout 200 python -m pytest tests/test_gpu_group.py "tests/test_gpu_search.py::test_warp_per_query_kernel_equals_cta_kernel" -q > "$OUT/pytest.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
timeout 120 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -1 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
