#!/usr/bin/env python
"""Latency of the host-buffer matching calls at the sizes the reference's own use produces (a few thousand descriptors per image)
and around the popc/tensor switch-over:  python tools/match_latency.py [nq x ndb ...]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ak = importlib.import_module("akaze-rust_b200")


def med(f, n=20):
    for _ in range(3):
        f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(32, 32), (128, 128), (256, 256), (512, 512), (1000, 1000), (7395, 5629), (30000, 30000)]
    rng = np.random.default_rng(1)
    eng = ak.Engine(0, 640, 480, 1)
    for nq, ndb in sizes:
        q = rng.integers(0, 256, (nq, 61), dtype=np.uint8)
        db = rng.integers(0, 256, (ndb, 61), dtype=np.uint8)
        row = {}
        for path in ("popc", "tensor", "auto"):
            eng.set_match_path(path)
            row[path] = med(lambda: eng.match_top2(q, db))
        eng.set_match_path("auto")
        row["descriptor_match"] = med(lambda: eng.descriptor_match(q, db))
        print("%6d x %6d  " % (nq, ndb) + "  ".join("%s %.3f ms" % kv for kv in row.items()))
    eng.close()


if __name__ == "__main__":
    main()
