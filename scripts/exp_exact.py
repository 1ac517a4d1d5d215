"""Experiment: exhaustive search, tensor-core path vs SIMT path.  python scripts/exp_exact.py [n] [nq] [dim]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lantern_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 768
dev = torch.device("cuda", 0)
X = bench.structured_torch(n, dim, 42, dev)
Q = bench.structured_torch(nq, dim, 43, dev)
k = 10
out = {}
stream = torch.cuda.current_stream()
for mode in ("tc", "simt"):
    os.environ["LB200_EXACT"] = mode
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dists = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        api.exact_search_device(X.data_ptr(), n, dim * 4, Q.data_ptr(), nq, dim * 4, k, keys.data_ptr(), dists.data_ptr(), "cos", "f32", dim,
                                stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[mode] = (keys.cpu().numpy(), dists.cpu().numpy())
    print("%s: %.2f ms  %.1f TFLOP/s (2*n*nq*d%s)" % (mode, ms, (3 if mode == "tc" else 1) * 2.0 * n * nq * dim / ms / 1e9, " x 3 tf32 MMAs" if mode == "tc" else ""), flush=True)
print("identical ids", np.array_equal(out["tc"][0], out["simt"][0]), "bit-identical distances", np.array_equal(out["tc"][1].view(np.uint32), out["simt"][1].view(np.uint32)))
