"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle_port.so (hnsw_oracle.c).

The plain-C restatement of the reference algorithm.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle_port.so")

METRIC = {"cos": 1, "ip": 2, "l2sq": 3, "hamming": 8}
SCALAR = {"f32": 1, "f64": 2, "f16": 3, "i8": 4, "b1": 5}


class Stats(C.Structure):
    _fields_ = [("computed_distances", C.c_uint64), ("visited_members", C.c_uint64), ("base_pops", C.c_uint64),
                ("upper_hops", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        L.ora_init.restype = C.c_void_p
        L.ora_init.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t,
                               C.c_size_t, C.c_void_p, C.c_int]
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_reserve.argtypes = [C.c_void_p, C.c_size_t]
        for f in ("ora_size", "ora_dimensions", "ora_connectivity", "ora_serialized_length"):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_void_p]
        L.ora_max_level.argtypes = [C.c_void_p]
        L.ora_set_engine_order.argtypes = [C.c_void_p, C.c_int]
        L.ora_entry_slot.restype = C.c_uint64
        L.ora_entry_slot.argtypes = [C.c_void_p]
        L.ora_add.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.POINTER(Stats)]
        L.ora_add_batch_engine.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t,
                                           C.c_size_t]
        L.ora_search.restype = C.c_size_t
        L.ora_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                 C.POINTER(Stats)]
        L.ora_node_level.argtypes = [C.c_void_p, C.c_size_t]
        L.ora_node_key.restype = C.c_uint64
        L.ora_node_key.argtypes = [C.c_void_p, C.c_size_t]
        L.ora_node_neighbors.restype = C.c_size_t
        L.ora_node_neighbors.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.ora_save_buffer.restype = C.c_size_t
        L.ora_save_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ora_load_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ora_distance.restype = C.c_float
        L.ora_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
        L.ora_cast_f32.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.ora_f16_to_f32.restype = C.c_float
        L.ora_f16_to_f32.argtypes = [C.c_uint16]
        L.ora_f32_to_f16.restype = C.c_uint16
        L.ora_f32_to_f16.argtypes = [C.c_float]
        L.ora_pq_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.ora_pq_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ora_exact_search.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                       C.c_size_t, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ora_level_sequence.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p]
        L.ora_kmeans.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p,
                                 C.c_void_p]
        _lib = L
    return _lib


class PortIndex:
    def __init__(self, dim, metric="l2sq", quant="f32", M=16, efc=128, ef=64, pq=False, num_centroids=0,
                 num_subvectors=0, codebook=None, pq_compat128=True):
        L = lib()
        self.dim, self.metric, self.quant, self.M = dim, metric, quant, M
        self._codebook = None
        cb = None
        if pq:
            self._codebook = np.ascontiguousarray(codebook, dtype=np.float32)
            cb = self._codebook.ctypes.data
        self.h = L.ora_init(METRIC[metric], SCALAR[quant], dim, M, efc, ef, int(pq), num_centroids, num_subvectors, cb,
                            int(pq_compat128))
        if not self.h:
            raise RuntimeError("ora_init failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_free(self.h)
            self.h = None

    @staticmethod
    def _kind(arr):
        return SCALAR["b1"] if arr.dtype == np.uint8 else SCALAR["f32"]

    def reserve(self, n):
        if not lib().ora_reserve(self.h, n):
            raise MemoryError

    def size(self):
        return lib().ora_size(self.h)

    def set_engine_order(self, on=True):
        lib().ora_set_engine_order(self.h, int(on))

    def add(self, key, vec, level=-1):
        vec = np.ascontiguousarray(vec)
        st = Stats()
        rc = lib().ora_add(self.h, int(key), vec.ctypes.data, self._kind(vec), int(level), C.byref(st))
        if rc:
            raise RuntimeError("ora_add failed: %d" % rc)
        return st

    def add_batch_engine(self, keys, vecs, batch_cap, build_ratio=64):
        """Model of the CUDA engine's batched build schedule (hnsw_oracle.c: ora_add_batch_engine); batch_cap=1 == add()."""
        vecs = np.ascontiguousarray(vecs)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        rc = lib().ora_add_batch_engine(self.h, keys.ctypes.data, vecs.ctypes.data, len(vecs), vecs.strides[0], self._kind(vecs),
                                        int(batch_cap), int(build_ratio))
        if rc:
            raise RuntimeError("ora_add_batch_engine failed: %d" % rc)

    def search(self, q, k, ef=0):
        q = np.ascontiguousarray(q)
        keys = np.zeros(k, np.uint64)
        dists = np.zeros(k, np.float32)
        st = Stats()
        n = lib().ora_search(self.h, q.ctypes.data, self._kind(q), k, ef, keys.ctypes.data, dists.ctypes.data,
                             C.byref(st))
        return keys[:n], dists[:n], st

    def search_batch(self, queries, k, ef=0):
        nq = len(queries)
        keys = np.full((nq, k), np.iinfo(np.uint64).max, np.uint64)
        dists = np.full((nq, k), np.inf, np.float32)
        counts = np.zeros(nq, np.int64)
        tot = dict(computed_distances=0, visited_members=0, base_pops=0, upper_hops=0)
        for i in range(nq):
            kk, dd, st = self.search(queries[i], k, ef)
            keys[i, :len(kk)], dists[i, :len(kk)], counts[i] = kk, dd, len(kk)
            for f in tot:
                tot[f] += getattr(st, f)
        return keys, dists, counts, tot

    def level(self, slot):
        return lib().ora_node_level(self.h, slot)

    def neighbors(self, slot, level):
        out = np.zeros(4 * self.M + 8, np.uint32)
        n = lib().ora_node_neighbors(self.h, slot, level, out.ctypes.data)
        return out[:n].copy()

    def max_level(self):
        return lib().ora_max_level(self.h)

    def entry_slot(self):
        return lib().ora_entry_slot(self.h)

    def save_buffer(self):
        n = lib().ora_serialized_length(self.h)
        buf = np.zeros(n, np.uint8)
        w = lib().ora_save_buffer(self.h, buf.ctypes.data, n)
        assert w == n, (w, n)
        return buf

    def load_buffer(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        rc = lib().ora_load_buffer(self.h, buf.ctypes.data, len(buf))
        if rc:
            raise RuntimeError("ora_load_buffer failed: %d" % rc)


def distance(a, b, metric, quant="f32", dims=None):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if dims is None:
        dims = a.size * 8 if quant == "b1" else a.size
    return lib().ora_distance(a.ctypes.data, b.ctypes.data, SCALAR[quant], dims, METRIC[metric])


def cast_f32(v, quant):
    v = np.ascontiguousarray(v, dtype=np.float32)
    d = v.shape[-1]
    shape = {"f32": (d, np.float32), "f16": (d, np.uint16), "i8": (d, np.int8), "b1": ((d + 7) // 8, np.uint8)}[quant]
    flat = v.reshape(-1, d)
    out = np.zeros((len(flat), shape[0]), shape[1])
    for i in range(len(flat)):
        lib().ora_cast_f32(flat[i].ctypes.data, d, SCALAR[quant], out[i].ctypes.data)
    return out.reshape(v.shape[:-1] + (shape[0],))


def pq_compress(codebook, vecs, num_subvectors, compat128=True):
    codebook = np.ascontiguousarray(codebook, dtype=np.float32)
    vecs = np.ascontiguousarray(vecs, dtype=np.float32).reshape(-1, codebook.shape[1])
    out = np.zeros((len(vecs), num_subvectors), np.uint8)
    for i in range(len(vecs)):
        lib().ora_pq_compress(codebook.ctypes.data, codebook.shape[1], codebook.shape[0], num_subvectors,
                              vecs[i].ctypes.data, out[i].ctypes.data, int(compat128))
    return out


def pq_decompress(codebook, codes):
    codebook = np.ascontiguousarray(codebook, dtype=np.float32)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    codes2 = codes.reshape(-1, codes.shape[-1])
    out = np.zeros((len(codes2), codebook.shape[1]), np.float32)
    for i in range(len(codes2)):
        lib().ora_pq_decompress(codebook.ctypes.data, codebook.shape[1], codebook.shape[0], codes2.shape[1],
                                codes2[i].ctypes.data, out[i].ctypes.data)
    return out


def exact_search(dataset, queries, k, metric="l2sq", quant="f32", dims=None):
    dataset, queries = np.ascontiguousarray(dataset), np.ascontiguousarray(queries)
    if dims is None:
        dims = dataset.shape[1] * 8 if quant == "b1" else dataset.shape[1]
    nq = len(queries)
    keys = np.zeros((nq, k), np.uint64)
    dists = np.zeros((nq, k), np.float32)
    lib().ora_exact_search(dataset.ctypes.data, len(dataset), dataset.strides[0], queries.ctypes.data, nq,
                           queries.strides[0], SCALAR[quant], dims, METRIC[metric], k, keys.ctypes.data,
                           dists.ctypes.data)
    return keys, dists


def level_sequence(M, count):
    out = np.zeros(count, np.int16)
    lib().ora_level_sequence(M, count, out.ctypes.data)
    return out


def kmeans(data, num_subvectors, num_centroids, init_rows, metric="l2sq", max_iter=20):
    data = np.ascontiguousarray(data, dtype=np.float32)
    init_rows = np.ascontiguousarray(init_rows, dtype=np.uint32)
    cb = np.zeros((num_centroids, data.shape[1]), np.float32)
    rounds = lib().ora_kmeans(data.ctypes.data, len(data), data.shape[1], num_subvectors, num_centroids, METRIC[metric], max_iter,
                              init_rows.ctypes.data, cb.ctypes.data)
    return cb, rounds
