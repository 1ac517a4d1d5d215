#!/bin/bash
# Round-2 GPU call 3 (two GPUs): group over real NVLink peers (single-process and CUDA-IPC multi-process), tensor-core exact
# search and the new PQ evaluator against their references.
set -u
OUT=gpurun_out/r2_call3
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_group.py -q -s > "$OUT/pytest_group.log" 2>&1
echo "pytest group rc=$?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/pytest_group.log" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_exact_tc.py tests/test_gpu_pq.py tests/test_gpu_golden.py -q -s > "$OUT/pytest_tc_pq.log" 2>&1
echo "pytest tc/pq rc=$?" | tee -a "$OUT/summary.txt"; tail -12 "$OUT/pytest_tc_pq.log" | tee -a "$OUT/summary.txt"
timeout 600 python scripts/exp_group.py 2 1000000 cos 4096 > "$OUT/exp_group_2dev.log" 2>&1
echo "exp rc=$?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/exp_group_2dev.log" | tee -a "$OUT/summary.txt"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py \
    --gpus 2 --workload cfg3s --steps 40 --warmup 5 > "$OUT/bench_cfg3s_n2.json" 2> "$OUT/bench_cfg3s_n2.err"
echo "bench n2 rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench_cfg3s_n2.err" | tee -a "$OUT/summary.txt"; tail -c 1500 "$OUT/bench_cfg3s_n2.json" | tee -a "$OUT/summary.txt"
