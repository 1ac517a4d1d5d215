"""Experiment: the row-sharded group on whatever devices the box has (ranks > devices: several ranks per device).
    python scripts/exp_group.py [ranks] [n] [workload-like: cos|l2sq] [batch]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lantern_b200 import api  # noqa: E402

ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
metric = sys.argv[3] if len(sys.argv) > 3 else "cos"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
dim, k, ef, M = 768, 10, 128 if metric == "cos" else 64, 32 if metric == "cos" else 16
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
kind, rowb = "f32", dim * 4
if metric == "hamming":  # cfg5-like: 768-byte binary rows
    dim, kind, rowb = 6144, "b1", 768
    X = bench.bits_torch(n, dim, 42, dev)
    Q = bench.bits_torch(4 * B, dim, 43, dev).cpu().numpy()
else:
    X = bench.structured_torch(n, dim, 42, dev)
    Q = bench.structured_torch(4 * B, dim, 43, dev).cpu().numpy()
idx = api.Index(dim, metric, kind, M=M, efc=128, ef=ef)
idx.reserve(n)
t0 = time.perf_counter()
idx.add_batch_device(np.arange(1, n + 1, dtype=np.uint64), X.data_ptr(), n, rowb, kind)
idx.build()
torch.cuda.synchronize()
print("build %.1f s" % (time.perf_counter() - t0), idx.last_build_stats(), flush=True)
k1, d1, c1 = idx.search_batch(Q[:B], k, ef)
st1 = idx.last_stats()
for _ in range(3):
    idx.search_batch(Q[:B], k, ef)
st1 = idx.last_stats()
print("1 GPU kernel (one CTA per query): %.3f ms  %.0f q/s  evals/q %.0f  alg GB/s %.0f" % (
    st1["kernel_ms"], B / st1["kernel_ms"] * 1e3, st1["computed_distances"] / B, st1["algorithmic_bytes"] / st1["kernel_ms"] / 1e6), flush=True)
idx.set_option("search_kernel", 2)
kw, dw, cw = idx.search_batch(Q[:B], k, ef)
for _ in range(3):
    idx.search_batch(Q[:B], k, ef)
stw = idx.last_stats()
print("1 GPU kernel (one warp per query): %.3f ms  %.0f q/s  identical rows %.4f  bit-equal dists %s  same counters %s  alg GB/s %.0f" % (
    stw["kernel_ms"], B / stw["kernel_ms"] * 1e3, float(np.mean(np.all(kw == k1, axis=1))), np.array_equal(dw.view(np.uint32), d1.view(np.uint32)),
    stw["computed_distances"] == st1["computed_distances"], stw["algorithmic_bytes"] / stw["kernel_ms"] / 1e6), flush=True)
idx.set_option("search_kernel", 1)
ndev = api.device_count()
for rk in sorted(set([1, ranks])):
    grp = api.Group.local([r % ndev for r in range(rk)])
    grp.distribute(idx, root=0, max_batch=B)
    for rep in range(4):
        t0 = time.perf_counter()
        kg, dg, cg = grp.search_batch(Q[(rep % 4) * B:(rep % 4 + 1) * B], k, ef)
        wall = time.perf_counter() - t0
    kg, dg, cg = grp.search_batch(Q[:B], k, ef)
    sts = [grp.last_stats(r) for r in range(rk)]
    same = float(np.mean(np.all(kg == k1, axis=1)))
    kms = max(s["kernel_ms"] for s in sts)
    print("group ranks=%d on %d device(s): kernel %.3f ms (max over ranks)  %.0f q/s  wall(host api) %.1f ms  identical rows %.4f  "
          "bit-equal dists %s  rows/query per rank %s  rounds/q %.1f" % (
              rk, ndev, kms, B / kms * 1e3, wall * 1e3, same, np.array_equal(dg.view(np.uint32), d1.view(np.uint32)),
              [round(s["local_rows_evaluated"] / B) for s in sts], sum(s["owner_rounds"] for s in sts) / B), flush=True)
    for s in sts:  # owner phases per round, microseconds at 1.965 GHz; own rows per round -> microseconds per row
        rounds = max(1, s["owner_rounds"])
        us = [s["owner_cycles_" + ph] / rounds / 1965.0 for ph in ("produce", "local", "wait", "consume")]
        own_rows = (s["owner_computed_distances"] / max(1, rk)) / rounds
        print("   rank %d: produce %.2f  local rows %.2f  wait %.2f  insert %.2f us/round;  ~%.2f us per own row" % (
            s["rank"], us[0], us[1], us[2], us[3], us[1] / max(own_rows, 1e-9)), flush=True)
    grp.close()
