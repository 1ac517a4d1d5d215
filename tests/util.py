"""Shared helpers for the parity tests."""
import numpy as np


def structured(n, d, seed, latent=32, noise=0.05):
    """Low-intrinsic-dimension synthetic vectors: x = z P + noise * eps."""
    rng = np.random.default_rng(seed)
    P = (np.random.default_rng(1234).standard_normal((latent, d)) / np.sqrt(latent)).astype(np.float32)
    z = rng.standard_normal((n, latent)).astype(np.float32)
    return (z @ P + noise * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def build_port_index(portlib, X, metric="l2sq", quant="f32", M=16, efc=128, ef=64, keys=None, **kw):
    n = len(X)
    dim = X.shape[1] * 8 if X.dtype == np.uint8 else X.shape[1]
    idx = portlib.PortIndex(dim, metric, quant, M=M, efc=efc, ef=ef, **kw)
    idx.reserve(n)
    keys = np.arange(1, n + 1, dtype=np.uint64) if keys is None else keys
    for i in range(n):
        idx.add(keys[i], X[i])
    return idx


def compare_results(keys_a, dists_a, keys_b, dists_b, rtol=1e-5, atol=1e-6):
    """Position-wise id equality, tolerating permutations inside groups of (near-)equal distances.
    Returns (n_exact_rows, n_rows_ok)."""
    nq = len(keys_a)
    exact = ok = 0
    for q in range(nq):
        ka, kb, da, db = keys_a[q], keys_b[q], dists_a[q], dists_b[q]
        if np.array_equal(ka, kb):
            exact += 1
            ok += 1
            continue
        # distances must agree position-wise within tolerance and differing ids must sit in near-tie groups
        if not np.allclose(da, db, rtol=rtol * 10, atol=atol * 10):
            continue
        good = True
        for i in np.nonzero(ka != kb)[0]:
            # id kb[i] must appear in ka at a position whose distance is within tolerance of da[i]
            pos = np.nonzero(ka == kb[i])[0]
            if len(pos) == 0:
                # allowed only at the tail boundary: distance equal to the last one
                if not np.isclose(db[i], da[-1], rtol=rtol * 10, atol=atol * 10):
                    good = False
            elif not np.isclose(da[pos[0]], db[i], rtol=rtol * 10, atol=atol * 10):
                good = False
        ok += good
    return exact, ok


def recall(found, truth):
    hits = 0
    for f, t in zip(found, truth):
        hits += len(set(f.tolist()) & set(t.tolist()))
    return hits / truth.size
