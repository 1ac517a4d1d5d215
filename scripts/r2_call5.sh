#!/bin/bash
# Round-2 GPU call 5 (two GPUs): group kernel v3 (owner / helper-pool roles).
set -u
OUT=gpurun_out/r2_call5
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_group.py "tests/test_gpu_search.py::test_warp_per_query_kernel_equals_cta_kernel" -q -s -x > "$OUT/pytest_group.log" 2>&1
echo "pytest group rc=$?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/pytest_group.log" | tee -a "$OUT/summary.txt"
for B in 4096 2048; do
timeout 600 python scripts/exp_group.py 2 1000000 cos $B > "$OUT/exp_group_2dev_B$B.log" 2>&1
echo "exp B=$B rc=$?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/exp_group_2dev_B$B.log" | tee -a "$OUT/summary.txt"
done
LB200_GROUP_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py \
    --gpus 2 --workload cfg3s --steps 40 --warmup 5 > "$OUT/bench_cfg3s_n2.json" 2> "$OUT/bench_cfg3s_n2.err"
echo "bench n2 rc=$?" | tee -a "$OUT/summary.txt"; grep "lb200 group rank" "$OUT/bench_cfg3s_n2.err" | tail -4 | tee -a "$OUT/summary.txt"
python - "$OUT/bench_cfg3s_n2.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "recall", d["recall_at_10"])
    print("same graph:", d["sharding"]["same_graph_as_1gpu"], "1gpu:", d["sharding"]["one_gpu_same_run"], "speedup", d["sharding"]["speedup_vs_one_gpu_same_run"])
    print("per rank:", d["sharding"]["per_rank"])
except Exception as e:
    print("no bench line:", e)
PY
