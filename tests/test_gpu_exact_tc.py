"""Exhaustive search on the tensor cores (csrc/exact_tc.cu: tcgen05 3xTF32 filter + exact fp32 re-rank) against the SIMT
kernels (csrc/exact.cu), which are themselves pinned to usearch_exact_search (tests/test_gpu_golden.py): same offsets,
bit-identical distances, ties included -- the tensor cores only decide WHICH rows get the exact arithmetic."""
import os

import numpy as np
import pytest

from util import structured

pytestmark = pytest.mark.gpu


def both(eng, X, Q, k, metric):
    os.environ["LB200_EXACT"] = "simt"
    try:
        ks, ds = eng.exact_search(X, Q, k, metric)
        os.environ["LB200_EXACT"] = "tc"
        os.environ["LB200_EXACT_REPORT"] = "1"
        kt, dt = eng.exact_search(X, Q, k, metric)
    finally:
        os.environ.pop("LB200_EXACT", None)
        os.environ.pop("LB200_EXACT_REPORT", None)
    return ks, ds, kt, dt


@pytest.mark.parametrize("metric", ["l2sq", "cos"])
@pytest.mark.parametrize("n,d,nq,k", [(70_000, 96, 300, 10), (40_000, 100, 130, 100), (5_000, 768, 64, 10), (300, 32, 5, 20)])
def test_tensor_core_exact_search_equals_simt(eng, metric, n, d, nq, k):
    X = structured(n, d, seed=5)
    Q = structured(nq, d, seed=6)
    ks, ds, kt, dt = both(eng, X, Q, k, metric)
    assert np.array_equal(kt, ks), float(np.mean(kt == ks))
    assert np.array_equal(dt.view(np.uint32), ds.view(np.uint32))


def test_tensor_core_exact_search_with_exact_ties_and_zero_rows(eng):
    """Integer vectors: whole groups of rows at exactly the same distance (the order is by offset); zero rows and a zero query
    exercise the cosine special cases (index_plugins.hpp:1022-1026)."""
    rng = np.random.default_rng(3)
    X = rng.integers(-2, 3, (20_000, 24)).astype(np.float32)
    X[::997] = 0
    Q = rng.integers(-2, 3, (40, 24)).astype(np.float32)
    Q[3] = 0
    for metric in ("l2sq", "cos"):
        ks, ds, kt, dt = both(eng, X, Q, 25, metric)
        assert np.array_equal(kt, ks) and np.array_equal(dt.view(np.uint32), ds.view(np.uint32))


def test_default_path_uses_the_tensor_cores_for_large_problems(eng):
    """No environment override: n >= 32768 rows of f32 take the tensor-core path; results equal the forced SIMT run."""
    X = structured(40_000, 64, seed=9)
    Q = structured(50, 64, seed=10)
    k0, d0 = eng.exact_search(X, Q, 10, "l2sq")
    os.environ["LB200_EXACT"] = "simt"
    try:
        k1, d1 = eng.exact_search(X, Q, 10, "l2sq")
    finally:
        os.environ.pop("LB200_EXACT", None)
    assert np.array_equal(k0, k1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
