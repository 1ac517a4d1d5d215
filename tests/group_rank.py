"""One rank of tests/test_gpu_group.py::test_group_multi_process_ipc (launched by torchrun): rank 0 builds an index, the
group distributes it over CUDA IPC, every rank searches collectively and checks the result against rank 0's 1-GPU search."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lantern_b200 import api  # noqa: E402
from util import structured  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
torch.cuda.set_device(local)


def allgather(send):
    outs = [None] * world
    dist.all_gather_object(outs, send)
    return b"".join(outs)


n, d, nq, k, ef = 8000, 128, 256, 10, 64
Q = structured(nq, d, seed=12)
grp = api.Group.ranked(rank, world, allgather)
idx = None
if rank == 0:
    X = structured(n, d, seed=11)
    idx = api.Index(d, "cos", "f32", M=16, efc=64, ef=ef)
    idx.reserve(n)
    idx.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    idx.build()
grp.distribute(idx, root=0, max_batch=nq)
# host buffers: only the root passes queries, every rank receives every result
keys, dists, counts = grp.search_batch(Q if rank == 0 else None, k, ef, nq=nq, dim_bytes=d * 4, kind="f32")
# device buffers, asynchronous on the current stream
stream = torch.cuda.current_stream()
dq = torch.from_numpy(Q).cuda() if rank == 0 else None
dk = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
dd = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
dc = torch.zeros((nq,), dtype=torch.int32, device="cuda")
for _ in range(3):
    grp.search_batch_device(dq.data_ptr() if rank == 0 else 0, nq, d * 4, "f32", k, ef, dk.data_ptr(), dd.data_ptr(), dc.data_ptr(),
                            stream.cuda_stream)
torch.cuda.synchronize()
st = grp.last_stats()
ref = [None]
if rank == 0:
    k1, d1, c1 = idx.search_batch(Q, k, ef)
    ref = [(k1, d1, idx.last_stats()["computed_distances"])]
dist.broadcast_object_list(ref, src=0)
k1, d1, total = ref[0]
assert np.array_equal(keys, k1) and np.array_equal(dists.view(np.uint32), d1.view(np.uint32)), float(np.mean(keys == k1))
assert np.array_equal(dk.cpu().numpy().astype(np.uint64), k1) and np.array_equal(dd.cpu().numpy().view(np.uint32), d1.view(np.uint32))
evals = [None] * world
dist.all_gather_object(evals, (st["owner_computed_distances"], st["local_rows_evaluated"]))
assert sum(e[0] for e in evals) == total and sum(e[1] for e in evals) == total - nq, (evals, total)
assert st["local_rows_evaluated"] > 0 and 0 < st["rows_held"] < n
print("group rank ok %d/%d: %d local rows evaluated, kernel %.3f ms" % (rank, world, st["local_rows_evaluated"], st["kernel_ms"]))
grp.close()
dist.destroy_process_group()
