"""CPU-only, world_size 2 over gloo: the multi-GPU matching plan (database sharded contiguously by index,
queries replicated, all-gather of the per-shard top-2 records, merge with the lowest-index tie rule).
Per-shard results come from the oracle and the merge from tests/shard_merge.py; the product's pieces
(akz_match_top2_device with db_index_base, k_merge_top2) are checked against the same functions in the
gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shard_merge import merge_top2_numpy, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(nq, ndb, seed):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (nq, 61), dtype=np.uint8)
    db = rng.integers(0, 256, (ndb, 61), dtype=np.uint8)
    q[:, 60] &= 0x3F
    db[:, 60] &= 0x3F
    db[5] = q[0]
    db[ndb - 2] = q[0]      # duplicate of the best in the other shard: lowest index must win
    db[ndb // 2 + 3] = q[1]
    return q, db


def _worker(rank, world, port, nq, ndb, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import akaze_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, db = _make(nq, ndb, 123)
    lo, hi = shard_range(ndb, world, rank)
    bi, b, s = O.match_top2(q, db[lo:hi])
    rec = torch.from_numpy(np.stack([bi.astype(np.int64) + lo, b.astype(np.int64), s.astype(np.int64)], 1))
    gathered = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec)
    parts = [(g[:, 0].numpy().astype(np.uint32), g[:, 1].numpy().astype(np.uint32), g[:, 2].numpy().astype(np.uint32)) for g in gathered]
    mi, mb, ms = merge_top2_numpy(parts)
    fi, fb, fs = O.match_top2(q, db)
    ok = bool(np.array_equal(mi, fi) and np.array_equal(mb, fb) and np.array_equal(ms, fs))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_match_world2(oracle):
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 64, 301, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_merge_edge_cases(oracle):
    q, db = _make(16, 40, 7)
    full = oracle.match_top2(q, db)
    for world in (1, 2, 3, 5, 8, 64):
        parts = []
        for r in range(world):
            lo, hi = shard_range(len(db), world, r)
            bi, b, s = oracle.match_top2(q, db[lo:hi])  # empty shards give (0, 10000, 10000)
            parts.append((bi + np.uint32(lo), b, s))
        m = merge_top2_numpy(parts)
        assert all(np.array_equal(a, b) for a, b in zip(m, full)), world
