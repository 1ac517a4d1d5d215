"""Dev experiment: run bench workloads under different LB200_* tuning environments.
usage: python scripts/exp_env.py WORKLOAD STEPS 'ENVJSON' ['ENVJSON' ...]"""
import json, os, subprocess, sys
wl, steps = sys.argv[1], sys.argv[2]
for spec in sys.argv[3:]:
    env = dict(os.environ, **json.loads(spec))
    out = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", steps, "--warmup", "3", "--no-cpu-baseline"],
                         capture_output=True, text=True, env=env)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(wl, spec, "qps %.0f" % d["value"], "recall %.4f" % d["recall_at_k"], "roofline %.3f" % d["roofline"]["frac"],
              "build %.1fs" % d["build"]["seconds"], flush=True)
    except Exception as ex:
        print(wl, spec, "failed", ex, out.stderr[-500:])
