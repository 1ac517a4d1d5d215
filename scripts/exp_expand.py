"""Dev experiment: relaxed-order search (--search-expand) on the latency-bound workloads."""
import json, subprocess, sys
for wl, steps, es in (("cfg5t", 60, (1, 2, 4)), ("cfg4s", 30, (1, 2, 4))):
    for e in es:
        out = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline",
                              "--search-expand", str(e)], capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(wl, "expand", e, "qps %.0f" % d["value"], "recall %.4f" % d["recall_at_k"], "roofline %.3f" % d["roofline"]["frac"],
                  "dist/q %.0f" % d["roofline"]["dist_evals_per_query"], flush=True)
        except Exception as ex:
            print(wl, e, "failed", ex, out.stderr[-500:])
