"""lantern_b200: B200-native HNSW search/build engine behind Lantern's usearch C API.

The product is `liblantern_b200.so` (CUDA, sm_100a) with the C ABI in include/lantern_b200.h;
`lantern_b200.api` is a thin ctypes harness over it used by the tests and bench.py.
"""
from . import api  # noqa: F401
