"""Row-range sharding helpers: which rows a rank owns, and the host-side statement of the per-query
G-way merge that `lb200_merge_shards_device` performs on the GPU after the all-gather of per-shard top-k lists."""
import numpy as np

EMPTY_KEY = np.iinfo(np.uint64).max


def row_range(n, rank, world):
    """Contiguous row range [lo, hi) of `rank` out of `world` shards; sizes differ by at most one."""
    return (n * rank) // world, (n * (rank + 1)) // world


def global_keys(lo, hi):
    """Keys stay global across shards (usearch_key_t = the caller's label): row r has key r + 1 in the benches."""
    return np.arange(lo + 1, hi + 1, dtype=np.uint64)


def merge_topk_host(keys, dists):
    """keys/dists: [G][nq][k] per-shard ascending lists (EMPTY_KEY / +inf padded).  Returns [nq][k]: the k smallest by
    (distance, key) -- the order the CUDA merge kernel uses (ties -> lower key)."""
    keys = np.asarray(keys, dtype=np.uint64)
    dists = np.asarray(dists, dtype=np.float32)
    G, nq, k = keys.shape
    out_k = np.full((nq, k), EMPTY_KEY, np.uint64)
    out_d = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        kk = keys[:, q, :].reshape(-1)
        dd = dists[:, q, :].reshape(-1)
        valid = kk != EMPTY_KEY
        kk, dd = kk[valid], dd[valid]
        order = np.lexsort((kk, dd))[:k]
        out_k[q, :len(order)], out_d[q, :len(order)] = kk[order], dd[order]
    return out_k, out_d
