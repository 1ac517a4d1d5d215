"""Dev experiment: search-kernel throughput vs tuning knobs on one built index (needs a GPU).
usage: python scripts/exp_search.py [n] [batch] ; env SWEEP='[{"LB200_RING_BYTES":"24576"}, ...]'"""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from lantern_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dim, k, ef = 768, 10, int(os.environ.get("EF", "64"))
M = int(os.environ.get("M", "16"))
dev = torch.device("cuda", 0)
X = bench.structured_torch(n, dim, 42, dev)
nb = 8
Q = bench.structured_torch(nb * B, dim, 43, dev)
idx = api.Index(dim, os.environ.get("METRIC", "l2sq"), "f32", M=M, efc=128, ef=ef)
idx.reserve(n)
t = time.time(); idx.add_batch_device(np.arange(1, n + 1, dtype=np.uint64), X.data_ptr(), n, dim * 4, "f32"); idx.build(); torch.cuda.synchronize()
print("build %.1fs" % (time.time() - t))
ok = torch.empty((B, k), dtype=torch.int64, device=dev); od = torch.empty((B, k), dtype=torch.float32, device=dev)
nrec = min(B, 512)
tk = torch.empty((nrec, k), dtype=torch.int64, device=dev); td = torch.empty((nrec, k), dtype=torch.float32, device=dev)
api.exact_search_device(X.data_ptr(), n, dim * 4, Q.data_ptr(), nrec, dim * 4, k, tk.data_ptr(), td.data_ptr(), os.environ.get("METRIC", "l2sq"), "f32", dim)
torch.cuda.synchronize(); truth = (tk + 1).cpu().numpy()
stream = torch.cuda.current_stream()
sweep = json.loads(os.environ.get("SWEEP", '[{}]'))
base = None
for env in sweep:
    for kk in list(os.environ):
        if kk.startswith("LB200_"):
            del os.environ[kk]
    os.environ.update({kk: vv for kk, vv in env.items() if kk.startswith("LB200_")})
    idx.set_option("search_expand", int(env.get("expand", 1)))
    for s in range(3):
        idx.search_batch_device(Q[s * B].data_ptr(), B, dim * 4, "f32", k, ef, ok.data_ptr(), od.data_ptr(), 0, stream.cuda_stream)
    torch.cuda.synchronize()
    res = ok.clone()
    idx.search_batch_device(Q.data_ptr(), B, dim * 4, "f32", k, ef, ok.data_ptr(), od.data_ptr(), 0, stream.cuda_stream)
    torch.cuda.synchronize()
    rec = bench.recall_at_k(ok[:nrec].cpu().numpy(), truth)
    ms, ab, nd = 0.0, 0, 0
    for s in range(nb):
        idx.search_batch_device(Q[s * B].data_ptr(), B, dim * 4, "f32", k, ef, ok.data_ptr(), od.data_ptr(), 0, stream.cuda_stream)
        st = idx.last_stats(); ms += st["kernel_ms"]; ab += st["algorithmic_bytes"]; nd += st["computed_distances"]
    if base is None:
        base = res
    same = float((res == base).float().mean())
    print("%-28s kernel %.3f ms/step  %.0f qps  %.0f GB/s (%.1f%% of 6571)  recall %.4f  dist/q %.0f  ids==first %.4f" % (json.dumps(env), ms / nb, B * nb / (ms / 1e3), ab / (ms / 1e3) / 1e9, ab / (ms / 1e3) / 1e9 / 65.712, rec, nd / (nb * B), same))
