#!/bin/bash
# Round-2 GPU call 6 (eight GPUs): the group over 8 real NVLink peers: tests, then the 1/10-scale metric workload at N = 8 and 4.
set -u
OUT=gpurun_out/r2_call6
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_group.py -q -s > "$OUT/pytest_group.log" 2>&1
echo "pytest group rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/pytest_group.log" | tee -a "$OUT/summary.txt"
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2981$N bench.py \
    --gpus $N --workload cfg3s --steps 60 --warmup 5 > "$OUT/bench_cfg3s_n$N.json" 2> "$OUT/bench_cfg3s_n$N.err"
echo "bench n$N rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/bench_cfg3s_n$N.err" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_cfg3s_n$N.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "recall", d["recall_at_10"])
    print("same graph:", d["sharding"]["same_graph_as_1gpu"], "1gpu:", d["sharding"]["one_gpu_same_run"]["value"], "speedup", d["sharding"]["speedup_vs_one_gpu_same_run"])
    print("per rank:", [(r["rank"], round(r["rows_evaluated_per_query"]), round(r["kernel_ms"], 3), round(r["local_hbm_frac"], 3)) for r in d["sharding"]["per_rank"]])
    print("roofline", d["roofline"]["frac"], "rounds/q", d["roofline"]["rounds_per_query"])
except Exception as e:
    print("no bench line:", e)
PY
done
