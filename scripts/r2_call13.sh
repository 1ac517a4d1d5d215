#!/bin/bash
# Round-2 GPU call 13 (two GPUs): rows in flight per warp in the group kernel (ring depth 2 / 3 / 4), 1 M rows, 2 ranks.
set -u
OUT=gpurun_out/r2_call13
mkdir -p "$OUT"
for R in 2 3 4; do
LB200_GROUP_RING_SLOTS=$R timeout 400 python scripts/exp_group.py 2 1000000 cos 4096 > "$OUT/exp_group_ring$R.log" 2>&1
echo "ring $R rc=$?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/exp_group_ring$R.log" | tee -a "$OUT/summary.txt"
done
