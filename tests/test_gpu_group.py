"""GPU parity of the row-sharded multi-GPU group (csrc/group.cu): ONE graph, vectors sharded by row range, every
distance evaluated on the GPU that holds the row.  The contract is stronger than recall parity: on the same graph the
group returns exactly what the 1-GPU kernel returns (same ids, bit-identical distances, same work counters), which in
turn is checked against the oracle elsewhere (tests/test_gpu_search.py).

A single-process group may place several ranks on ONE physical device (they split its resident warps), so the whole
G-rank protocol -- requests, responses, START/EXIT, the fused result all-gather, the done exchange -- is exercised on a
1-GPU box too; the multi-process (CUDA IPC) flavour needs two devices and is skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import structured

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _index(eng, metric, quant, d, n, M=16, efc=64, ef=48, seed=3):
    if quant == "b1":
        rng = np.random.default_rng(seed)
        protos = rng.integers(0, 256, (16, d // 8), dtype=np.uint8)
        X = protos[rng.integers(0, 16, n)] ^ np.packbits(rng.random((n, d)) < 0.1, axis=1)
        Q = protos[rng.integers(0, 16, 300)] ^ np.packbits(rng.random((300, d)) < 0.1, axis=1)
    else:
        X, Q = structured(n, d, seed=seed), structured(300, d, seed=seed + 1)
        if quant == "i8":
            X, Q = X * 0.3, Q * 0.3
    g = eng.Index(d, metric, quant, M=M, efc=efc, ef=ef)
    g.reserve(n)
    g.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    g.build()
    return g, X, Q


@pytest.mark.parametrize("ranks", [1, 2, 4, 8])
@pytest.mark.parametrize("metric,quant,d", [("cos", "f32", 96), ("l2sq", "f32", 768), ("hamming", "b1", 512), ("l2sq", "f16", 72)])
def test_group_equals_one_gpu(eng, ranks, metric, quant, d):
    ndev = eng.device_count()
    devices = [r % ndev for r in range(ranks)]  # several ranks per device when the box has fewer GPUs
    g, X, Q = _index(eng, metric, quant, d, 6000 if d < 768 else 3000)
    k, ef = 10, 48
    k1, d1, c1 = g.search_batch(Q, k, ef)
    st1 = g.last_stats()
    grp = eng.Group.local(devices)
    grp.distribute(g, root=0, max_batch=512)
    for rep in range(2):  # the second batch reuses the mailboxes with new flags
        kg, dg, cg = grp.search_batch(Q, k, ef)
        assert np.array_equal(kg, k1), float(np.mean(kg == k1))
        assert np.array_equal(dg.view(np.uint32), d1.view(np.uint32))
        assert np.array_equal(cg, c1)
    stats = [grp.last_stats(r) for r in range(ranks)]
    assert sum(s["owner_computed_distances"] for s in stats) == st1["computed_distances"]
    assert sum(s["owner_base_pops"] for s in stats) == st1["base_pops"]
    # every evaluation happened on the rank that holds the row; the re-measured start node (index.hpp:3436) is counted, not read
    assert sum(s["local_rows_evaluated"] for s in stats) == st1["computed_distances"] - len(Q)
    assert sum(s["rows_held"] for s in stats) == len(X)
    if ranks > 1:
        assert all(s["local_rows_evaluated"] > 0 for s in stats)
    grp.close()


@pytest.mark.parametrize("M", [32, 48])
def test_group_wide_lists(eng, M):
    """M = 32 (the metric's configuration: 64-wide base lists take the two-halves-at-once path, where a duplicate id may sit in
    either half) and M = 48 (96-wide lists, the general path); duplicates are planted on purpose in a few lists."""
    g, X, Q = _index(eng, "cos", "f32", 64, 5000, M=M, efc=96, ef=64)
    # plant duplicates the way refine_'s stale padding produces them: rewrite the file, load it back
    buf = g.save_buffer().copy()
    n = 5000
    off, planted = 136, 0
    for i in range(n):
        lvl = int(np.frombuffer(buf, np.int16, 1, off + 8)[0])
        base = off + 10
        cnt = int(np.frombuffer(buf, np.uint32, 1, base)[0])
        if i % 7 == 0 and cnt >= 40:
            ids = buf[base + 4: base + 4 + 6 * cnt].reshape(cnt, 6)
            ids[cnt - 1] = ids[2]      # second half repeats an id of the first half
            ids[5] = ids[4]            # and a duplicate inside the first half
            planted += 1
        off = base + 4 + 12 * M + lvl * (4 + 6 * M) + 64 * 4
    assert planted > 50
    g2 = eng.Index(64, "cos", "f32", M=M, efc=96, ef=64)
    g2.load_buffer(buf)
    k1, d1, c1 = g2.search_batch(Q, 10, 64)
    st1 = g2.last_stats()
    ndev = eng.device_count()
    grp = eng.Group.local([r % ndev for r in range(2)])
    grp.distribute(g2, root=0, max_batch=512)
    kg, dg, cg = grp.search_batch(Q, 10, 64)
    assert np.array_equal(kg, k1) and np.array_equal(dg.view(np.uint32), d1.view(np.uint32)) and np.array_equal(cg, c1)
    assert sum(grp.last_stats(r)["owner_computed_distances"] for r in range(2)) == st1["computed_distances"]
    grp.close()


def test_group_wide_beam_small_batch_and_k(eng):
    """ef = 400, k = 100, fewer queries than ranks, and a batch that is not a multiple of the rank count."""
    g, X, Q = _index(eng, "l2sq", "f32", 48, 5000, M=8, efc=48, ef=64)
    ndev = eng.device_count()
    grp = eng.Group.local([r % ndev for r in range(4)])
    grp.distribute(g, root=1, max_batch=512, max_results=512 * 100)
    for nq, k, ef in ((3, 100, 400), (61, 7, 16), (1, 1, 1)):
        k1, d1, c1 = g.search_batch(Q[:nq], k, ef)
        kg, dg, cg = grp.search_batch(Q[:nq], k, ef)
        assert np.array_equal(kg, k1) and np.array_equal(dg.view(np.uint32), d1.view(np.uint32)) and np.array_equal(cg, c1)
    grp.close()


def test_group_refuses_what_it_cannot_do(eng):
    g, X, Q = _index(eng, "l2sq", "f32", 32, 500, M=8, efc=32, ef=16)
    grp = eng.Group.local([0])
    with pytest.raises(eng.EngineError, match="distribute"):
        grp.search_batch(Q, 5)
    grp.distribute(g, root=0, max_batch=16)
    with pytest.raises(eng.EngineError, match="max_batch"):
        grp.search_batch(Q[:64], 5)
    grp.close()


def test_group_multi_process_ipc(eng):
    """One process per GPU (torchrun), CUDA IPC peer mappings, gloo as the bootstrap all-gather."""
    if eng.device_count() < 2:
        pytest.skip("needs two CUDA devices (processes cannot run kernels concurrently on one GPU)")
    world = min(4, eng.device_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29741", os.path.join(ROOT, "tests", "group_rank.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("group rank ok") == world, out.stdout
