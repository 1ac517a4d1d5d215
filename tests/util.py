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


def adjacency_lists(buf, M, stored_bytes):
    """Parse a usearch/lantern index file into {(node, level): tuple(neighbour ids)} plus the level array.
    Vectors and keys are skipped: comparisons through this helper look at the graph only."""
    b = np.asarray(buf, dtype=np.uint8).tobytes()
    n = int(np.frombuffer(b, np.uint64, 1, 80)[0])
    M0 = 2 * M
    lists, levels = {}, np.zeros(n, np.int16)
    p = 136
    for i in range(n):
        lvl = int(np.frombuffer(b, np.int16, 1, p + 8)[0])
        levels[i] = lvl
        p += 10
        for l in range(lvl + 1):
            width = M0 if l == 0 else M
            cnt = int(np.frombuffer(b, np.uint32, 1, p)[0])
            raw = np.frombuffer(b, np.uint8, 6 * width, p + 4).reshape(width, 6)[:cnt]
            ids = raw[:, :4].copy().view(np.uint32)[:, 0]
            lists[(i, l)] = tuple(int(x) for x in ids)
            p += 4 + 6 * width
        p += stored_bytes
    assert p == len(b), (p, len(b))
    return lists, levels


def graph_agreement(buf_a, buf_b, M, stored_bytes):
    """Fraction of (node, level) adjacency lists that are identical (same ids, same order) in two index files."""
    la, lva = adjacency_lists(buf_a, M, stored_bytes)
    lb, lvb = adjacency_lists(buf_b, M, stored_bytes)
    assert np.array_equal(lva, lvb), "level draws differ"
    assert la.keys() == lb.keys()
    same = sum(1 for k in la if la[k] == lb[k])
    return same / max(1, len(la))
