"""GPU matcher parity (bit-exact) through the C ABI, including the sharded path and size-independent
properties at sizes the oracle cannot finish."""
import numpy as np
import pytest
import torch

from shard_merge import merge_top2_numpy, shard_range

pytestmark = pytest.mark.gpu


def make(nq, ndb, seed, width=64):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (nq, width), dtype=np.uint8)
    db = rng.integers(0, 256, (ndb, width), dtype=np.uint8)
    for a in (q, db):
        if a.shape[0]:
            a[:, 60] &= 0x3F
            a[:, 61:] = 0
    return q, db


@pytest.fixture(params=["popc", "tensor"])
def path_engine(engine, request):
    """The session engine forced onto one matcher kernel: integer popc (matcher.cu) or tcgen05 int8 (matcher_tc.cu)."""
    engine.set_match_path(request.param)
    yield engine
    engine.set_match_path("auto")


def assert_same(t, o):
    bi, b, s = o
    assert np.array_equal(t["best"], b) and np.array_equal(t["second"], s) and np.array_equal(t["best_idx"], bi)


@pytest.mark.parametrize("nq,ndb", [(1, 1), (1, 0), (3, 2), (31, 33), (512, 128), (513, 129), (300, 1000),
                                     (2000, 5000), (64, 70000), (7395, 5629), (257, 127), (255, 1025), (128, 128)])
def test_top2_exact(path_engine, oracle, nq, ndb):
    engine = path_engine
    q, db = make(nq, ndb, nq * 7919 + ndb)
    k = min(nq, ndb) // 3
    if k:
        db[:k] = q[:k]          # distance-0 hits
        db[ndb - 1] = db[0]     # tie on the best distance at a higher index: the lower index must win
    assert_same(engine.match_top2(q, db, desc_len=61), oracle.match_top2(q, db, desc_len=61))


def test_unpadded_rows_and_desc_len(path_engine, oracle):
    engine = path_engine
    q, db = make(200, 333, 5, width=61)  # rows exactly as Descriptor.vector (61 bytes, descriptors.rs:45)
    assert_same(engine.match_top2(q, db), oracle.match_top2(q, db))
    q, db = make(100, 100, 6)
    q[:, 40:] = 0xFF                      # bytes beyond desc_len must be ignored
    assert_same(engine.match_top2(q, db, desc_len=32), oracle.match_top2(q, db, desc_len=32))


def test_full_width_rows(path_engine, oracle):
    """All 512 bits of a 64-byte row in use (desc_len = 64): no spare bits for either kernel to rely on."""
    engine = path_engine
    rng = np.random.default_rng(77)
    q = rng.integers(0, 256, (300, 64), dtype=np.uint8)
    db = rng.integers(0, 256, (900, 64), dtype=np.uint8)
    db[5] = q[7]
    db[600] = q[7]
    db[100] = ~q[9]
    assert_same(engine.match_top2(q, db, desc_len=64), oracle.match_top2(q, db, desc_len=64))


def test_all_equal_and_far(path_engine, oracle):
    engine = path_engine
    q = np.zeros((50, 64), np.uint8)
    db = np.zeros((70, 64), np.uint8)
    t = engine.match_top2(q, db, desc_len=61)
    assert np.all(t["best"] == 0) and np.all(t["second"] == 0) and np.all(t["best_idx"] == 0)
    db[:] = 0xFF
    db[:, 60] = 0x3F
    db[:, 61:] = 0
    t = engine.match_top2(q, db, desc_len=61)
    assert np.all(t["best"] == 486) and np.all(t["second"] == 486) and np.all(t["best_idx"] == 0)


def test_descriptor_match_lowe(path_engine, oracle):
    engine = path_engine
    q, db = make(500, 800, 9)
    for i in range(0, 500, 5):                  # plant near matches so that the Lowe test passes sometimes
        db[(i * 7) % 800] = q[i]
        db[(i * 7) % 800, i % 60] ^= 0x11
    for ratio in (0.86, 0.5, 1.0, 1.2):
        assert np.array_equal(engine.descriptor_match(q, db, 10000, ratio, desc_len=61),
                              oracle.descriptor_match(q, db, 10000, ratio, desc_len=61))


def _dev(a):
    return torch.from_numpy(a).cuda()


def test_sharded_device_path_matches_unsharded(path_engine, oracle, akz):
    """Database sharded contiguously, per-shard top-2 with db_index_base, merged by the CUDA merge kernel
    (what each rank does after the NCCL all-gather)."""
    engine = path_engine
    q, db = make(700, 4099, 31)
    db[4000] = q[0]
    db[17] = q[0]
    dq = _dev(q)
    full = oracle.match_top2(q, db, desc_len=61)
    for world in (1, 2, 3, 8):
        parts = torch.zeros((world, len(q)), dtype=torch.int64, device="cuda")  # 8-byte akz_top2 records
        keep_alive = []  # the engine runs on its own stream: torch must not recycle a shard while it is in use
        for r in range(world):
            lo, hi = shard_range(len(db), world, r)
            ddb = _dev(db[lo:hi]) if hi > lo else torch.zeros((1, 64), dtype=torch.uint8, device="cuda")
            keep_alive.append(ddb)
            torch.cuda.synchronize()
            engine.match_top2_device(dq.data_ptr(), len(q), ddb.data_ptr(), hi - lo, parts[r].data_ptr(), db_index_base=lo)
        out = torch.zeros(len(q), dtype=torch.int64, device="cuda")
        engine.merge_top2_device(parts.data_ptr(), world, len(q), out.data_ptr())
        torch.cuda.synchronize()
        t = out.cpu().numpy().view(akz.TOP2_DTYPE)
        assert_same(t, full)
        # and the numpy merge used by the gloo test agrees with the CUDA merge
        pn = parts.cpu().numpy().view(akz.TOP2_DTYPE).reshape(world, len(q))
        m = merge_top2_numpy([(p["best_idx"], p["best"].astype(np.uint32), p["second"].astype(np.uint32)) for p in pn])
        assert np.array_equal(m[0], t["best_idx"]) and np.array_equal(m[1], t["best"]) and np.array_equal(m[2], t["second"])


def test_large_properties(engine):
    """256k x 256k (6.9e10 pairs): properties that need no oracle. db = permuted queries with <= 8 flipped
    bits each, so every query's best is its own copy (random 486-bit strings are ~243 bits apart)."""
    n = 1 << 18
    g = torch.Generator(device="cuda").manual_seed(7)
    q = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=g)
    q[:, 60] &= 0x3F
    q[:, 61:] = 0
    perm = torch.randperm(n, device="cuda", generator=g)
    db = q[perm].clone()
    flips = torch.randint(0, 9, (n,), device="cuda", generator=g)
    for b in range(8):
        col = torch.randint(0, 60, (n,), device="cuda", generator=g)
        bit = (1 << torch.randint(0, 8, (n,), device="cuda", generator=g)).to(torch.uint8)
        sel = flips > b
        rows = torch.nonzero(sel).squeeze(1)
        db[rows, col[rows]] ^= bit[rows]
    out = torch.zeros(n, dtype=torch.int64, device="cuda")
    engine.match_top2_device(q.data_ptr(), n, db.data_ptr(), n, out.data_ptr())
    torch.cuda.synchronize()
    rec = out.cpu().numpy().view(np.dtype([("best_idx", "<u4"), ("best", "<u2"), ("second", "<u2")]))
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device="cuda")
    assert np.array_equal(rec["best_idx"], inv.cpu().numpy().astype(np.uint32))
    assert rec["best"].max() <= 8 and np.all(rec["second"] >= rec["best"]) and rec["second"].min() > 100
    # distances recomputed on the device with torch for a sample
    idx = torch.arange(0, n, 997, device="cuda")
    x = q[idx] ^ db[inv[idx]]
    pop = torch.zeros(len(idx), dtype=torch.int64, device="cuda")
    for b in range(8):
        pop += ((x >> b) & 1).sum(dim=1)
    assert np.array_equal(pop.cpu().numpy(), rec["best"][idx.cpu().numpy()].astype(np.int64))
