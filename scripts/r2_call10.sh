#!/bin/bash
# Round-2 GPU call 10 (eight GPUs): the metric's configuration (cfg3, 10 M rows) on 8 GPUs, launched exactly as the driver does.
set -u
OUT=gpurun_out/r2_call10
mkdir -p "$OUT"
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29888 bench.py \
    --gpus 8 --steps 100 --warmup 5 ) > "$OUT/bench_cfg3_n8.json" 2> "$OUT/bench_cfg3_n8.err"
echo "bench n8 rc=$?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/bench_cfg3_n8.err" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_cfg3_n8.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "recall", d["recall_at_10"])
    print("same graph:", d["sharding"]["same_graph_as_1gpu"], "1gpu:", d["sharding"]["one_gpu_same_run"], "speedup", d["sharding"]["speedup_vs_one_gpu_same_run"])
    print("per rank:", [(r["rank"], round(r["rows_evaluated_per_query"]), round(r["kernel_ms"], 3), round(r["local_hbm_frac"], 3)) for r in d["sharding"]["per_rank"]])
    print("roofline", d["roofline"]["frac"], "rounds/q", d["roofline"]["rounds_per_query"], "distribute s", d["sharding"]["distribute_seconds"], "build", d["build"])
except Exception as e:
    print("no bench line:", e)
PY
