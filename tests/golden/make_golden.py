"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the build container (needs oracle/_ref):  python tests/golden/make_golden.py
Inputs are the reference's own test fixtures (file:line under /root/reference):
  * small_world cube                      lantern_hnsw/test/sql/utils/small_world_array.sql, expected/hnsw_dist_func.out
  * 4 collinear points, M=4               lantern_hnsw/test/sql/hnsw_correct.sql:8-46
  * 14-point lattice M=12 efc=64 ef=32    lantern_cli/tests/external_index_server_test.rs:217-232 (f32), :587-602 (hamming)
  * toy PQ codebook 4 centroids x d3      lantern_cli/tests/external_index_server_test.rs:684-690
plus seeded random cases.  Everything stored is an OUTPUT OF THE REFERENCE (keys, distances, file bytes).
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reflib  # noqa: E402

CUBE = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], np.float32)
LATTICE = np.array([[0, 0, 0], [0, 0, 1], [0, 0, 2], [0, 0, 3], [0, 1, 0], [0, 1, 1], [0, 1, 2], [0, 1, 3], [1, 0, 0], [1, 0, 1],
                    [1, 0, 2], [1, 0, 3], [1, 1, 0], [1, 1, 1]], np.float32)
PQ_CODEBOOK = np.array([[0.0, 0.1, 0.0], [0.1, 0.1, 0.1], [0.1, 0.1, 0.2], [0.1, 0.2, 0.1]], np.float32)


def trim(buf):
    n, M, M0 = struct.unpack_from("<QQQ", buf, 80)
    vsz, = struct.unpack_from("<Q", buf, 120)
    off, b = 136, bytes(buf)
    for _ in range(n):
        lvl, = struct.unpack_from("<h", b, off + 8)
        off += 10 + 4 + 6 * M0 + lvl * (4 + 6 * M) + vsz
    return np.array(buf[:off])


def int_bits(v):  # Lantern's hamming payload: integer[] reinterpreted as bits (hnsw.c:316-318)
    return np.ascontiguousarray(v.astype(np.int32)).view(np.uint8).reshape(len(v), -1)


def build(X, metric, quant="f32", keys=None, **kw):
    dim = X.shape[1] * 8 if X.dtype == np.uint8 else X.shape[1]
    idx = reflib.RefIndex(dim, metric, quant, **kw)
    idx.reserve(len(X))
    for i, v in enumerate(X):
        idx.add(i if keys is None else keys[i], v)
    return idx


def searches(idx, Q, k):
    ks = np.full((len(Q), k), np.iinfo(np.uint64).max, np.uint64)
    ds = np.full((len(Q), k), np.inf, np.float32)
    for i, q in enumerate(Q):
        kk, dd = idx.search(q, k)
        ks[i, :len(kk)], ds[i, :len(kk)] = kk, dd
    return ks, ds


out = {}
# --- small world ---
for metric in ("l2sq", "cos"):
    idx = build(CUBE, metric, keys=100 + np.arange(8), M=2, efc=128, ef=4)
    out["cube_%s_keys" % metric], out["cube_%s_dists" % metric] = searches(idx, CUBE, 8)
    out["cube_%s_file" % metric] = trim(idx.save_buffer())
idx = build(int_bits(CUBE), "hamming", "b1", keys=100 + np.arange(8), M=2, efc=128, ef=4)
out["cube_hamming_keys"], out["cube_hamming_dists"] = searches(idx, int_bits(CUBE), 8)
# --- 4-NN per vertex, default M (hnsw_dist_func.sql:62-91) ---
idx = build(CUBE, "l2sq", keys=100 + np.arange(8), M=16, efc=128, ef=64)
out["cube_4nn_keys"], out["cube_4nn_dists"] = searches(idx, CUBE, 4)
# --- hnsw_correct: collinear points, M=4 ---
LINE = np.array([[0, 0], [1, 1], [2, 2], [3, 3]], np.float32)
idx = build(LINE, "l2sq", keys=1 + np.arange(4), M=4, efc=128, ef=64)
out["line_queries"] = np.array([[0, 0], [2, 2], [3.2, 3.2], [0.9, 1.2]], np.float32)
out["line_keys"], out["line_dists"] = searches(idx, out["line_queries"], 4)
# --- 14-point lattice ---
for quant in ("f32", "f16", "i8"):
    scale = 0.25 if quant == "i8" else 1.0
    idx = build(LATTICE * scale, "l2sq", quant, M=12, efc=64, ef=32)
    out["lattice_%s_keys" % quant], out["lattice_%s_dists" % quant] = searches(idx, LATTICE * scale, 5)
    out["lattice_%s_file" % quant] = trim(idx.save_buffer())
idx = build(int_bits(LATTICE), "hamming", "b1", M=12, efc=64, ef=32)
out["lattice_hamming_keys"], out["lattice_hamming_dists"] = searches(idx, int_bits(LATTICE), 5)
# quant_bits=1 storage with an l2sq opclass silently becomes hamming on sign bits (index_plugins.hpp:1465,1477)
# (dims a multiple of 8: for other dims the reference's f32->b1 cast clears dim/8 = too few bytes of its reused
#  buffer, index_plugins.hpp:913, and stale bits leak between vectors -- not a behaviour to pin)
SIGN_X = np.cos(np.arange(14 * 16).reshape(14, 16) * 0.7).astype(np.float32)
idx = build(SIGN_X, "l2sq", "b1", M=12, efc=64, ef=32)
out["lattice_signbits_keys"], out["lattice_signbits_dists"] = searches(idx, SIGN_X, 5)
# --- PQ toy codebook (3 subvectors x 4 centroids): distances are raw-query vs decoded-candidate ---
PQX = (LATTICE * 0.1).astype(np.float32)
idx = build(PQX, "l2sq", M=12, efc=64, ef=32, pq=True, num_centroids=4, num_subvectors=3, codebook=PQ_CODEBOOK)
out["pq_keys"], out["pq_dists"] = searches(idx, PQX, 5)
# --- seeded random graphs (built sequentially by the reference; integer-valued so that fp32 sums are order independent) ---
rng = np.random.default_rng(2024)
for name, metric, d, M, efc, ef in (("rand_l2", "l2sq", 16, 8, 48, 24), ("rand_cos", "cos", 12, 6, 40, 20)):
    X = rng.integers(-6, 7, (400, d)).astype(np.float32)
    Q = rng.integers(-6, 7, (40, d)).astype(np.float32)
    idx = build(X, metric, keys=1000 + np.arange(400), M=M, efc=efc, ef=ef)
    out[name + "_X"], out[name + "_Q"] = X, Q
    out[name + "_file"] = trim(idx.save_buffer())
    out[name + "_keys"], out[name + "_dists"] = searches(idx, Q, 10)
    ek, ed = reflib.exact_search(X, Q, 10, metric)
    out[name + "_exact_dists"] = ed  # offsets among ties are unspecified in the reference (partial_sort)
# --- continuous random data: distances, casts (usearch_distance; f16/i8/b1 via an index round trip) ---
A = rng.standard_normal((64, 40)).astype(np.float32) * 0.4
B = rng.standard_normal((64, 40)).astype(np.float32) * 0.4
A[0] = 0  # zero-norm rows exercise the cosine special cases
B[0] = 0
A[1] = 0
out["dist_A"], out["dist_B"] = A, B
for metric in ("l2sq", "cos"):
    out["dist_f32_" + metric] = np.array([reflib.distance(a, b, metric) for a, b in zip(A, B)], np.float32)
bits_a = np.packbits(A > 0, axis=1)
bits_b = np.packbits(B > 0, axis=1)
out["dist_b1_hamming"] = np.array([reflib.distance(a, b, "hamming", "b1", 40) for a, b in zip(bits_a, bits_b)], np.float32)
# storage-domain distances: 1-vector index per row, query -> distance returned by the reference in the storage kind
for quant in ("f16", "i8", "b1"):
    for metric in ("l2sq", "cos"):
        ds = []
        for a, b in zip(A[:24], B[:24]):
            idx = build(b[None, :], metric, quant, M=4, efc=8, ef=8)
            ds.append(idx.search(a, 1)[1][0])
        out["dist_%s_%s" % (quant, metric)] = np.array(ds, np.float32)
# level generator (index.hpp:3208-3212): levels of 5000 sequential adds with M=16 / M=4, read back from the saved file
for M in (16, 4):
    X = rng.standard_normal((5000, 2)).astype(np.float32)
    idx = build(X, "l2sq", M=M, efc=4, ef=4)
    buf = bytes(trim(idx.save_buffer()))
    off, lv = 136, []
    for _ in range(5000):
        l, = struct.unpack_from("<h", buf, off + 8)
        lv.append(l)
        off += 10 + 4 + 12 * M + l * (4 + 6 * M) + 8
    out["levels_M%d" % M] = np.array(lv, np.int16)

# --- usearch_exact_search (lib.cpp:450-481) on tie-free continuous data: OFFSETS and distances are pinned (k = 1 takes the
#     min_element branch of exact_search_t, index_plugins.hpp:1636-1650; k = 10 the partial_sort branch) ---
rng2 = np.random.default_rng(77)
for metric in ("l2sq", "cos"):
    X = rng2.standard_normal((500, 24)).astype(np.float32)
    Q = rng2.standard_normal((32, 24)).astype(np.float32)
    out["exact_%s_X" % metric], out["exact_%s_Q" % metric] = X, Q
    for kk in (1, 10):
        ek, ed = reflib.exact_search(X, Q, kk, metric)
        out["exact_%s_k%d_offsets" % (metric, kk)], out["exact_%s_k%d_dists" % (metric, kk)] = ek, ed
        gaps = np.diff(np.sort(np.array([reflib.distance(q, x, metric) for q in Q[:4] for x in X]).reshape(4, -1), axis=1)[:, :12], axis=1)
        assert gaps.min() > 1e-6  # tie-free: the offsets are well defined

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
