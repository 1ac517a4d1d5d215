"""Smoke-size walk through the search, build, PQ and exact kernels for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
Sizes are tiny on purpose: the sanitizer slows kernels down by one to two orders of magnitude."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lantern_b200 import api  # noqa: E402

rng = np.random.default_rng(0)
n, d, nq, k = 600, 48, 24, 5
X = rng.standard_normal((n, d)).astype(np.float32)
Q = rng.standard_normal((nq, d)).astype(np.float32)
for metric, quant in (("l2sq", "f32"), ("cos", "f16"), ("l2sq", "i8")):
    g = api.Index(d, metric, quant, M=8, efc=32, ef=24)
    g.reserve(n)
    g.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    g.build()
    keys, dists, counts = g.search_batch(Q, k)
    assert (counts == k).all()
    g.close()
B = rng.integers(0, 256, (n, 32), dtype=np.uint8)
g = api.Index(256, "hamming", "b1", M=8, efc=32, ef=24)
g.reserve(n)
g.add_batch(np.arange(1, n + 1, dtype=np.uint64), B)
g.build()
g.search_batch(B[:nq], k)
g.close()
cb, rounds = api.train_pq(X, 12, 16, "l2sq", 5, 3)
g = api.Index(d, "l2sq", "f32", M=8, efc=32, ef=24, pq=True, num_centroids=16, num_subvectors=12, codebook=cb)
g.reserve(n)
g.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
g.build()
g.search_batch(Q, k)
g.close()
api.exact_search(X, Q, k, "l2sq")
api.exact_search(X, Q, k, "cos")
print("sanitize_small: done, %d kernel launches" % api.kernel_launches())
