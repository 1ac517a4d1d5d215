"""Dev experiment: recall of the batched GPU build vs build_ratio / build_batch (needs a GPU)."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from lantern_b200 import api
from oracle import reflib
from util import structured, recall

n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 50000, int(sys.argv[2]) if len(sys.argv) > 2 else 64
X = structured(n, d, seed=3); Q = structured(1000, d, seed=4)
truth, _ = api.exact_search(X, Q, 10, "l2sq")
keys = np.arange(1, n + 1, dtype=np.uint64)
t = time.time()
r = reflib.RefIndex(d, "l2sq", M=16, efc=128, ef=64, threads=reflib.lib().refx_hardware_threads()); r.reserve(n)
r.add_batch(keys, X)
tb = time.time() - t
rk, _, _, comp, _ = r.search_batch(Q, 10)
print("reference multi-thread build %.1fs recall %.4f dist/q %.0f" % (tb, recall(rk - 1, truth), comp / len(Q)))
import os
combos = eval(os.environ.get('COMBOS', '[(64,592),(64,1024),(128,592),(128,1184),(256,592)]'))
for ratio, batch in combos:
    g = api.Index(d, "l2sq", "f32", M=16, efc=128, ef=64)
    g.set_option("build_ratio", ratio); g.set_option("build_batch", batch)
    g.reserve(n); g.add_batch(keys, X)
    t = time.time(); g.build(); tb = time.time() - t
    gk, _, _ = g.search_batch(Q, 10)
    st = g.last_stats()
    print("ratio %3d batch %5d: build %.2fs recall %.4f dist/q %.0f" % (ratio, batch, tb, recall(gk - 1, truth), st["computed_distances"] / len(Q)))
