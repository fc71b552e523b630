#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py
Covers the default-mode pipeline on two small shapes (streaming and fallback kernels), a dense-noise image (busy cache pass),
both matcher kernels, the one-rank NCCL path and GPU RANSAC."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import akaze_rust_b200 as A  # noqa: E402
import np_restatement as R  # noqa: E402

rng = np.random.default_rng(3)
for shape in ((272, 360), (135, 333)):
    eng = A.Engine(0, shape[1], shape[0], 3)
    fs = eng.extract_batch_u8([R.synthetic_image(shape[0], shape[1], seed=s) for s in (1, 2, 3)])
    print(shape, [f.count for f in fs])
    eng.close()
eng = A.Engine(0, 320, 240, 1)
f = eng.extract_u8(rng.integers(0, 256, (240, 320), dtype=np.uint8))
print("noise", f.count, f.num_candidates)
q = rng.integers(0, 256, (700, 64), dtype=np.uint8)
db = rng.integers(0, 256, (1900, 64), dtype=np.uint8)
q[:, 61:] = 0
db[:, 61:] = 0
for path in ("popc", "tensor"):
    eng.set_match_path(path)
    print(path, int(eng.match_top2(q, db, desc_len=61)["best"].sum()))
if "--no-nccl" not in sys.argv:
    A.comm_init_all([eng])
    print("sharded", int(A.match_top2_sharded([eng], q, db, desc_len=61)["best"].sum()))
    eng.comm_destroy()
k = np.zeros(60, A.KEYPOINT_DTYPE)
k["x"], k["y"] = rng.uniform(0, 600, 60), rng.uniform(0, 400, 60)
k2 = k.copy()
k2["x"] += 12.0
m = np.zeros(60, A.MATCH_DTYPE)
m["index_0"] = m["index_1"] = np.arange(60)
print("ransac", len(eng.remove_outliers(k, k2, m, 50, 0.05, 3.0, sampling="advancing")))
eng.close()
