#!/bin/bash
# Round-2 GPU call 12 (one GPU): the final tree through the whole GPU suite and smoke, exactly as the driver runs them.
set -u
OUT=gpurun_out/r2_call12
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -1 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --workload cfg3s --steps 30 --warmup 5 --no-cpu-baseline > "$OUT/bench_cfg3s.json" 2> "$OUT/bench_cfg3s.err"
echo "bench cfg3s rc=$?" | tee -a "$OUT/summary.txt"
python -c "
import json
d=json.loads(open('$OUT/bench_cfg3s.json').read().strip().splitlines()[-1]); print('cfg3s:', round(d['value']), 'q/s', d['roofline']['frac'], d['ground_truth'])" | tee -a "$OUT/summary.txt"
