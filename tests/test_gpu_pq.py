"""GPU parity of the PQ index path (asymmetric-distance search and build over PQ codes) against the oracle."""
import numpy as np
import pytest

from test_golden import LATTICE, PQ_CODEBOOK, G, same_up_to_ties

pytestmark = pytest.mark.gpu


def port_pq(port, X, cb, nsub, metric="l2sq", **kw):
    idx = port.PortIndex(X.shape[1], metric, "f32", pq=True, num_centroids=len(cb), num_subvectors=nsub, codebook=cb, **kw)
    idx.reserve(len(X))
    for i, v in enumerate(X):
        idx.add(i + 1, v)
    return idx


def test_toy_codebook_matches_reference_outputs(eng):
    """external_index_server_test.rs:684-690 codebook; expected keys/distances come from the unmodified reference."""
    X = (LATTICE * 0.1).astype(np.float32)
    g = eng.Index(3, "l2sq", "f32", M=12, efc=64, ef=32, pq=True, num_centroids=4, num_subvectors=3, codebook=PQ_CODEBOOK)
    g.set_option("build_batch", 1)
    g.reserve(len(X))
    g.add_batch(np.arange(len(X), dtype=np.uint64), X)
    g.build()
    k, d, _ = g.search_batch(X, 5)
    same_up_to_ties(k, d, G["pq_keys"], G["pq_dists"])


@pytest.mark.parametrize("metric", ["l2sq", "cos"])
def test_pq_same_graph_same_ids(eng, port, metric):
    rng = np.random.default_rng(8)
    d, nsub, ncent, n = 32, 8, 64, 2500
    cb = rng.standard_normal((ncent, d)).astype(np.float32)
    X = rng.standard_normal((n, d)).astype(np.float32)  # unstructured: codes are diverse, exact distance ties are rare
    Q = rng.standard_normal((150, d)).astype(np.float32)
    p = port_pq(port, X, cb, nsub, metric, M=16, efc=64, ef=48)
    g = eng.Index(d, metric, "f32", M=16, efc=64, ef=48, pq=True, num_centroids=ncent, num_subvectors=nsub, codebook=cb)
    g.load_buffer(p.save_buffer())  # the oracle's graph + codes
    gk, gd, _ = g.search_batch(Q, 10)
    pk, pd, _, tot = p.search_batch(Q, 10)
    assert np.allclose(gd, pd, rtol=2e-5, atol=2e-6)
    assert np.mean(np.all(gk == pk, axis=1)) > 0.95
    st = g.last_stats()
    assert abs(st["computed_distances"] - tot["computed_distances"]) <= 0.01 * tot["computed_distances"]


def test_pq_baseline_geometry_same_graph(eng, port):
    """BASELINE configs[3] geometry (d = 1536, 96 subvectors x 256 centroids, k = 100 => expansion 100) on a few hundred rows:
    the oracle builds the PQ graph (compat-128 encoder, decode-then-distance semantics of lantern_storage.hpp:249-270), the
    engine loads that file and must return the same ids with distances within 1e-5 -- look-up tables of 96 x 128 entries
    precomputed for the batch, six 16-byte chunks per code row."""
    rng = np.random.default_rng(21)
    d, nsub, ncent, n, nq, k = 1536, 96, 256, 700, 64, 100
    lat = rng.standard_normal((n + nq, 24)).astype(np.float32)
    P = rng.standard_normal((24, d)).astype(np.float32) / 5
    V = (lat @ P + 0.1 * rng.standard_normal((n + nq, d))).astype(np.float32)
    X, Q = V[:n], V[n:]
    cb = X[rng.choice(n, ncent, replace=False)].copy()  # a valid codebook: 256 corpus rows (per-subvector slices of them)
    p = port_pq(port, X, cb, nsub, "l2sq", M=8, efc=40, ef=64)
    g = eng.Index(d, "l2sq", "f32", M=8, efc=40, ef=64, pq=True, num_centroids=ncent, num_subvectors=nsub, codebook=cb)
    g.load_buffer(p.save_buffer())
    gk, gd, gc = g.search_batch(Q, k)
    pk, pd, pc, tot = p.search_batch(Q, k)
    assert np.array_equal(gc.astype(np.int64), pc)
    assert np.allclose(gd, pd, rtol=1e-5, atol=1e-5), np.abs(gd - pd).max()
    assert np.mean(gk == pk) > 0.98, np.mean(gk == pk)
    st = g.last_stats()
    assert abs(st["computed_distances"] - tot["computed_distances"]) <= 0.01 * tot["computed_distances"]
    # and the engine's own batched build at this geometry answers like the oracle-built graph
    g2 = eng.Index(d, "l2sq", "f32", M=8, efc=40, ef=64, pq=True, num_centroids=ncent, num_subvectors=nsub, codebook=cb)
    g2.reserve(n)
    g2.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    g2.build()
    k2, d2, _ = g2.search_batch(Q, 10)
    overlap = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(k2, pk[:, :10])])
    assert overlap > 0.9, overlap


def test_pq_exact_order_build_byte_identical_on_integer_data(eng, port):
    """Integer codebook + integer vectors: every distance is an exact small integer on both sides."""
    rng = np.random.default_rng(12)
    d, nsub, ncent, n = 16, 4, 16, 900
    cb = rng.integers(-4, 5, (ncent, d)).astype(np.float32)
    X = rng.integers(-5, 6, (n, d)).astype(np.float32)
    p = port_pq(port, X, cb, nsub, "l2sq", M=8, efc=48, ef=32)
    g = eng.Index(d, "l2sq", "f32", M=8, efc=48, ef=32, pq=True, num_centroids=ncent, num_subvectors=nsub, codebook=cb)
    g.set_option("build_batch", 1)
    g.reserve(n)
    g.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    g.build()
    gb, pb = g.save_buffer(), p.save_buffer()
    assert len(gb) == len(pb)
    assert np.mean(gb == pb) > 0.99  # ties everywhere (integer distances): only queue tie order may differ


def test_pq_batched_build_recall(eng, port):
    rng = np.random.default_rng(3)
    d, nsub, ncent, n = 64, 16, 256, 20000
    centers = rng.standard_normal((64, d)).astype(np.float32)
    X = (centers[rng.integers(0, 64, n)] + 0.4 * rng.standard_normal((n, d))).astype(np.float32)
    Q = (centers[rng.integers(0, 64, 300)] + 0.4 * rng.standard_normal((300, d))).astype(np.float32)
    # a simple codebook: per-subspace random sample of the data (k-means is not needed for a parity test)
    cb = X[rng.choice(n, ncent, replace=False)].copy()
    g = eng.Index(d, "l2sq", "f32", M=16, efc=128, ef=64, pq=True, num_centroids=ncent, num_subvectors=nsub, codebook=cb)
    g.reserve(n)
    g.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    g.build()
    gk, gd, _ = g.search_batch(Q, 10)
    # PQ-exact ground truth: brute force over the DECODED corpus (what the index can at best return)
    codes = eng.quantize_pq(cb, X, nsub, compat128=True)
    dec = eng.dequantize_pq(cb, codes)
    tk, td = eng.exact_search(dec, Q, 10, "l2sq")
    rec = np.mean([len(set(a.tolist()) & set((b + 1).tolist())) / 10 for a, b in zip(gk, tk)])
    assert rec > 0.85, rec  # hnsw_pq_index.sql:129-131: index recall within 0.1 of PQ-exact
    assert np.allclose(np.sort(gd[:, 0]), np.sort(np.minimum(gd[:, 0], td[:, 0])), rtol=1e-4) or rec > 0.9


@pytest.mark.parametrize("metric", ["l2sq", "cos"])
def test_codebook_training_matches_oracle(eng, port, metric):
    """GPU k-means (product_quantization.c semantics) against the CPU restatement from the same initial rows."""
    rng = np.random.default_rng(6)
    n, d, nsub, ncent = 4000, 24, 4, 16
    centers = rng.standard_normal((ncent, d)).astype(np.float32) * 3
    X = (centers[rng.integers(0, ncent, n)] + rng.standard_normal((n, d))).astype(np.float32)
    init = np.stack([rng.choice(n, ncent, replace=False) for _ in range(nsub)]).astype(np.uint32)
    gcb, grounds = eng.train_pq(X, nsub, ncent, metric, max_iter=25, init_rows=init)
    pcb, prounds = port.kmeans(X, nsub, ncent, init, metric, max_iter=25)
    assert grounds == prounds
    assert np.allclose(gcb, pcb, rtol=1e-4, atol=1e-4)
    # and the trained codebook is usable end to end
    g = eng.Index(d, metric, "f32", M=8, efc=48, ef=32, pq=True, num_centroids=ncent, num_subvectors=nsub, codebook=gcb)
    g.reserve(n)
    g.add_batch(np.arange(1, n + 1, dtype=np.uint64), X)
    k, dd, _ = g.search_batch(X[:50], 5)
    assert (k[:, 0] > 0).all()
    # random initialisation path (no init rows): distinct rows, converges
    cb2, r2 = eng.train_pq(X, nsub, ncent, metric, max_iter=25, seed=7)
    assert r2 >= 1 and np.isfinite(cb2).all()
