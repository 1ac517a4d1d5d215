#!/bin/bash
# Round-2 GPU call 11 (four GPUs): where an owner's time goes per expansion round (cycle counters), cfg3s on 4 GPUs.
set -u
OUT=gpurun_out/r2_call11
mkdir -p "$OUT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29844 bench.py \
    --gpus 4 --workload cfg3s --steps 40 --warmup 5 > "$OUT/bench_cfg3s_n4.json" 2> "$OUT/bench_cfg3s_n4.err"
echo "bench n4 rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/bench_cfg3s_n4.err" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_cfg3s_n4.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "speedup", d["sharding"]["speedup_vs_one_gpu_same_run"], d["sharding"]["same_graph_as_1gpu"])
    for r in d["sharding"]["per_rank"]:
        print(r)
    print("rounds/q", d["roofline"]["rounds_per_query"])
except Exception as e:
    print("no bench line:", e)
PY
