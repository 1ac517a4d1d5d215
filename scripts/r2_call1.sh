#!/bin/bash
# Round-2 GPU call 1 (one GPU): evidence gaps named by VERDICT r1 "Next round" item 1.
set -u
OUT=gpurun_out/r2_call1
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/nproc.txt"; python -c "import os;print(len(os.sched_getaffinity(0)))" >> "$OUT/nproc.txt"; cat /sys/fs/cgroup/cpu.max >> "$OUT/nproc.txt" 2>&1

timeout 900 python -m pytest tests -m gpu -q -x -s > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"

# compute-sanitizer on smoke-size kernels
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_small.py > "$OUT/sanitizer_$tool.log" 2>&1
  echo "sanitizer $tool rc=$?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/sanitizer_$tool.log" | tee -a "$OUT/summary.txt"
done

# ncu --set full: hamming (cfg5t), PQ (cfg4s), then cfg3 (the benchmarked configuration; 10M build runs unprofiled)
prof() { # name workload extra-args
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 \
      -o "$OUT/search_$1" -f python bench.py --workload $2 --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/ncu_$1.log" 2>&1
  echo "ncu $1 rc=$?" | tee -a "$OUT/summary.txt"
  ncu -i "$OUT/search_$1.ncu-rep" --page raw --csv > "$OUT/search_$1_raw.csv" 2>/dev/null
}
prof cfg5t cfg5t
prof cfg4s cfg4s
prof cfg3 cfg3
ls -la "$OUT" >> "$OUT/summary.txt"
