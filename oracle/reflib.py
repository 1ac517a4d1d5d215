"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/liboracle_usearch.so.

The .so is the UNMODIFIED reference (usearch fork, c/lib.cpp + headers under
/root/reference/lantern_hnsw/third_party/usearch) built by oracle/Makefile.  This module
mirrors `usearch_init_options_t` (c/usearch.h:74-117) and the handful of C entry points
Lantern calls.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "liboracle_usearch.so")

# usearch_metric_kind_t (usearch.h:51-63) / usearch_scalar_kind_t (:65-72)
METRIC = {"cos": 1, "ip": 2, "l2sq": 3, "hamming": 8}
SCALAR = {"f32": 1, "f64": 2, "f16": 3, "i8": 4, "b1": 5}


class InitOptions(C.Structure):
    _fields_ = [
        ("metric_kind", C.c_int),
        ("metric", C.c_void_p),
        ("quantization", C.c_int),
        ("dimensions", C.c_size_t),
        ("connectivity", C.c_size_t),
        ("expansion_add", C.c_size_t),
        ("expansion_search", C.c_size_t),
        ("multi", C.c_bool),
        ("retriever_ctx", C.c_void_p),
        ("retriever", C.c_void_p),
        ("retriever_mut", C.c_void_p),
        ("num_threads", C.c_size_t),
        ("pq", C.c_bool),
        ("num_centroids", C.c_size_t),
        ("num_subvectors", C.c_size_t),
    ]


def available():
    return os.path.exists(SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        L.usearch_init.restype = C.c_void_p
        L.usearch_init.argtypes = [C.POINTER(InitOptions), C.c_void_p, C.POINTER(C.c_char_p)]
        L.usearch_free.argtypes = [C.c_void_p, C.POINTER(C.c_char_p)]
        L.usearch_reserve.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p)]
        L.usearch_add.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
        L.usearch_search_ef.restype = C.c_size_t
        L.usearch_search_ef.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_bool,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_char_p)]
        L.usearch_size.restype = C.c_size_t
        L.usearch_size.argtypes = [C.c_void_p, C.POINTER(C.c_char_p)]
        L.usearch_serialized_length.restype = C.c_size_t
        L.usearch_serialized_length.argtypes = [C.c_void_p, C.POINTER(C.c_char_p)]
        L.usearch_save_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p)]
        L.usearch_load_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p)]
        L.usearch_save.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]
        L.usearch_load.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]
        L.usearch_distance.restype = C.c_float
        L.usearch_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int,
                                       C.POINTER(C.c_char_p)]
        L.usearch_exact_search.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                           C.c_int, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p,
                                           C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p)]
        L.refx_search_stats.restype = C.c_size_t
        L.refx_search_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.refx_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t,
                                        C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64)]
        L.refx_add_batch.restype = C.c_uint64
        L.refx_add_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                     C.c_size_t]
        L.refx_add_level.restype = C.c_uint64
        L.refx_add_level.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        L.refx_hardware_threads.restype = C.c_size_t
        _lib = L
    return _lib


def _check(err):
    if err.value:
        raise RuntimeError(err.value.decode())


class RefIndex:
    """The reference index behind its own C API (usearch_init ... usearch_search_ef)."""

    def __init__(self, dim, metric="l2sq", quant="f32", M=16, efc=128, ef=64, threads=1, pq=False,
                 num_centroids=0, num_subvectors=0, codebook=None):
        L = lib()
        self.dim, self.metric, self.quant = dim, metric, quant
        self.threads = max(1, min(threads, L.refx_hardware_threads()))
        o = InitOptions()
        o.metric_kind = METRIC[metric]
        o.quantization = SCALAR[quant]
        o.dimensions = dim
        o.connectivity, o.expansion_add, o.expansion_search = M, efc, ef
        o.num_threads = self.threads
        o.pq, o.num_centroids, o.num_subvectors = pq, num_centroids, num_subvectors
        self._codebook = None
        cb = None
        if pq:
            self._codebook = np.ascontiguousarray(codebook, dtype=np.float32)  # borrowed by the index
            cb = self._codebook.ctypes.data
        err = C.c_char_p()
        self.h = L.usearch_init(C.byref(o), cb, C.byref(err))
        _check(err)
        if not self.h:
            raise RuntimeError("usearch_init returned NULL")

    def __del__(self):
        if getattr(self, "h", None):
            err = C.c_char_p()
            lib().usearch_free(self.h, C.byref(err))
            self.h = None

    def _kind(self, arr):
        return SCALAR["b1"] if arr.dtype == np.uint8 else SCALAR["f32"]

    def reserve(self, n):
        err = C.c_char_p()
        lib().usearch_reserve(self.h, n, C.byref(err))
        _check(err)

    def add(self, key, vec):
        vec = np.ascontiguousarray(vec)
        err = C.c_char_p()
        lib().usearch_add(self.h, int(key), vec.ctypes.data, self._kind(vec), C.byref(err))
        _check(err)

    def add_level(self, key, vec, level):
        vec = np.ascontiguousarray(vec)
        return lib().refx_add_level(self.h, int(key), vec.ctypes.data, self._kind(vec), int(level))

    def add_batch(self, keys, vecs, threads=None):
        vecs = np.ascontiguousarray(vecs)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        return lib().refx_add_batch(self.h, keys.ctypes.data, vecs.ctypes.data, len(keys), vecs.strides[0],
                                    self._kind(vecs), threads or self.threads)

    def size(self):
        err = C.c_char_p()
        return lib().usearch_size(self.h, C.byref(err))

    def search(self, q, k):
        q = np.ascontiguousarray(q)
        keys = np.zeros(k, np.uint64)
        dists = np.zeros(k, np.float32)
        err = C.c_char_p()
        n = lib().usearch_search_ef(self.h, q.ctypes.data, self._kind(q), k, 0, False, keys.ctypes.data,
                                    dists.ctypes.data, C.byref(err))
        _check(err)
        return keys[:n], dists[:n]

    def search_batch(self, queries, k, threads=None):
        queries = np.ascontiguousarray(queries)
        nq = len(queries)
        keys = np.zeros((nq, k), np.uint64)
        dists = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint64)
        comp, vis = C.c_uint64(), C.c_uint64()
        lib().refx_search_batch(self.h, queries.ctypes.data, nq, queries.strides[0], self._kind(queries), k,
                                threads or self.threads, keys.ctypes.data, dists.ctypes.data, counts.ctypes.data,
                                C.byref(comp), C.byref(vis))
        return keys, dists, counts, comp.value, vis.value

    def save_buffer(self):
        err = C.c_char_p()
        n = lib().usearch_serialized_length(self.h, C.byref(err))
        buf = np.zeros(n + 64, np.uint8)
        lib().usearch_save_buffer(self.h, buf.ctypes.data, len(buf), C.byref(err))
        _check(err)
        return buf  # over-allocated: parse to find the true end

    def load_buffer(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        err = C.c_char_p()
        lib().usearch_load_buffer(self.h, buf.ctypes.data, len(buf), C.byref(err))
        _check(err)
        self._loaded = buf


def distance(a, b, metric, quant="f32", dims=None):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    err = C.c_char_p()
    if dims is None:
        dims = a.size * 8 if quant == "b1" else a.size
    return lib().usearch_distance(a.ctypes.data, b.ctypes.data, SCALAR[quant], dims, METRIC[metric], C.byref(err))


def exact_search(dataset, queries, k, metric="l2sq", quant="f32", dims=None, threads=1):
    dataset, queries = np.ascontiguousarray(dataset), np.ascontiguousarray(queries)
    if dims is None:
        dims = dataset.shape[1] * 8 if quant == "b1" else dataset.shape[1]
    nq = len(queries)
    keys = np.zeros((nq, k), np.uint64)
    dists = np.zeros((nq, k), np.float32)
    err = C.c_char_p()
    lib().usearch_exact_search(dataset.ctypes.data, len(dataset), dataset.strides[0], queries.ctypes.data, nq,
                               queries.strides[0], SCALAR[quant], dims, METRIC[metric], k, threads,
                               keys.ctypes.data, keys.strides[0], dists.ctypes.data, dists.strides[0], C.byref(err))
    _check(err)
    return keys, dists


# ---- the reference's k-means (product_quantization.c compiled unmodified: oracle/Makefile target refpq) ----------------
SO_PQ = os.path.join(HERE, "_ref", "liboracle_refpq.so")
_pqlib = None


def pq_available():
    return os.path.exists(SO_PQ) and available()


def ref_kmeans(data, num_subvectors, num_centroids, init_rows, metric="l2sq", max_iter=20):
    """product_quantization() of the reference from the given initial rows (init_rows[nsub][ncent], distinct per
    subvector); returns the codebook tape float[num_centroids][dim]."""
    global _pqlib
    if _pqlib is None:
        L = C.CDLL(SO_PQ)
        L.refpq_train.restype = C.c_int
        L.refpq_train.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
        _pqlib = L
    data = np.ascontiguousarray(data, dtype=np.float32)
    init_rows = np.ascontiguousarray(init_rows, dtype=np.uint32)
    assert init_rows.shape == (num_subvectors, num_centroids)
    assert all(len(set(r.tolist())) == num_centroids for r in init_rows), "initial rows must be distinct per subvector"
    cb = np.zeros((num_centroids, data.shape[1]), np.float32)
    _pqlib.refpq_train(data.ctypes.data, len(data), data.shape[1], num_subvectors, num_centroids, METRIC[metric], max_iter,
                       init_rows.ctypes.data, cb.ctypes.data)
    return cb
