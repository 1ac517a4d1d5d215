#!/bin/bash
# Round-2 GPU call 7 (one GPU): the whole GPU suite, sanitizer on the new kernels, narrow-row A/B, PQ, the metric's
# configuration with its CPU baseline, the reference arm.
set -u
OUT=gpurun_out/r2_call7
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q -s > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -1 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
for tool in memcheck racecheck; do
  LB200_GROUP_TIMEOUT_S=900 timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > "$OUT/sanitizer_$tool.log" 2>&1
  echo "sanitizer $tool rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/sanitizer_$tool.log" | tee -a "$OUT/summary.txt"
done
# narrow rows: warps per query / ring depth
for cfg in "4 12288" "2 12288" "2 6144"; do
  set -- $cfg
  LB200_SEARCH_WARPS=$1 LB200_RING_BYTES=$2 timeout 600 python bench.py --workload cfg5t --steps 100 --warmup 5 --no-cpu-baseline > "$OUT/bench_cfg5t_w$1_r$2.json" 2> "$OUT/bench_cfg5t_w$1_r$2.err"
  python -c "
import json,sys
d=json.loads(open('$OUT/bench_cfg5t_w$1_r$2.json').read().strip().splitlines()[-1]); print('cfg5t warps $1 ring $2:', round(d['value']), 'q/s  frac', round(d['roofline']['frac'],3), 'recall', d['recall_at_10'])" | tee -a "$OUT/summary.txt"
done
timeout 900 python bench.py --workload cfg4s --steps 50 --warmup 5 > "$OUT/bench_cfg4s.json" 2> "$OUT/bench_cfg4s.err"
echo "cfg4s rc=$?" | tee -a "$OUT/summary.txt"
python -c "
import json
d=json.loads(open('$OUT/bench_cfg4s.json').read().strip().splitlines()[-1]); print('cfg4s:', round(d['value']), 'q/s e2e', round(d['e2e']['value']), 'kernel ms', d['roofline']['kernel_ms_per_step'], 'recall@k', d['recall_at_k'], 'parity', d['parity'], 'cpu', d['cpu_baseline'])" | tee -a "$OUT/summary.txt"
timeout 1500 python bench.py > "$OUT/bench_cfg3.json" 2> "$OUT/bench_cfg3.err"
echo "cfg3 rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/bench_cfg3.err" | tee -a "$OUT/summary.txt"
python -c "
import json
d=json.loads(open('$OUT/bench_cfg3.json').read().strip().splitlines()[-1]); print('cfg3:', round(d['value']), 'q/s e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], d['roofline']['traffic'], 'recall', d['recall_at_10']); print('build', d['build']); print('gt', d['ground_truth']); print('cpu', d['cpu_baseline']); print('parity', d['parity'])" | tee -a "$OUT/summary.txt"
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > "$OUT/bench_cfg3_reference.json" 2> "$OUT/bench_cfg3_reference.err"
echo "reference rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/bench_cfg3_reference.err" | tee -a "$OUT/summary.txt"; tail -c 700 "$OUT/bench_cfg3_reference.json" | tee -a "$OUT/summary.txt"
