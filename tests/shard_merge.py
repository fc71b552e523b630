"""Test-side helpers for the sharded matcher: contiguous database partition and a numpy reference of the
top-2 merge (the product's merge is the CUDA kernel k_merge_top2; this is only the checker)."""
import numpy as np

SENTINEL = 10000


def shard_range(ndb, world, rank):
    """Contiguous partition by index: shard r owns [r*ceil(ndb/world), ...)."""
    per = (ndb + world - 1) // world
    lo = min(ndb, rank * per)
    return lo, min(ndb, lo + per)


def merge_top2_numpy(parts):
    """parts: list (ascending database ranges) of (best_idx, best, second) uint32 arrays with GLOBAL indices.
    Sequential-scan semantics of feature_matching.rs:37-50: two smallest of the union multiset, lowest
    index attaining the minimum."""
    nq = len(parts[0][0])
    best = np.full(nq, SENTINEL, np.int64)
    second = np.full(nq, SENTINEL, np.int64)
    idx = np.zeros(nq, np.int64)
    for bi, b, s in parts:
        b = b.astype(np.int64)
        s = s.astype(np.int64)
        better = b < best
        second = np.where(better, best, np.where(b < second, b, second))
        idx = np.where(better, bi, idx)
        best = np.where(better, b, best)
        second = np.where(s < second, s, second)
    none = best == SENTINEL
    idx = np.where(none, parts[0][0], idx)
    return idx.astype(np.uint32), best.astype(np.uint32), second.astype(np.uint32)
