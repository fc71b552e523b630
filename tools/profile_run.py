#!/usr/bin/env python
"""One small extraction (and optionally one match) with no torch kernels in the process, meant to be run
under ncu:

    ncu --set full --clock-control none --import-source on -k regex:k_ -o gpurun_out/extract \
        python tools/profile_run.py --images 8
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/l.csv \
        python tools/profile_run.py --images 64

Images come from the numpy generator of tests/np_restatement.py (same recipe as bench.py's workload).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--unique", type=int, default=4)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--runs", type=int, default=1)
    ap.add_argument("--match", type=int, default=0, help="also run an n x n match")
    ap.add_argument("--no-extract", action="store_true")
    args = ap.parse_args()
    import akaze_rust_b200 as A
    import np_restatement as R
    eng = A.Engine(0, args.width, args.height, max(1, args.images))
    if not args.no_extract:
        cache = os.path.join(ROOT, "gpurun_out", "_ab_imgs_%dx%d_%d.npy" % (args.width, args.height, args.unique))
        if os.path.exists(cache):  # written by tools/stage_ab.py: skips ~10 s of image synthesis per image
            uniq = list(np.load(cache))[:max(1, min(args.unique, args.images))]
        else:
            uniq = [R.natural_image(args.height, args.width, 1000 + i) for i in range(min(args.unique, args.images))]
        imgs = [uniq[i % len(uniq)] for i in range(args.images)]
        for _ in range(args.runs):
            fs = eng.extract_batch_u8(imgs)
            print("keypoints/image: %.1f" % np.mean([f.count for f in fs]))
            for f in fs:
                f.release()
    if args.match:
        rng = np.random.default_rng(1)
        q = rng.integers(0, 256, (args.match, 64), dtype=np.uint8)
        db = rng.integers(0, 256, (args.match, 64), dtype=np.uint8)
        q[:, 61:] = 0
        db[:, 61:] = 0
        t = eng.match_top2(q, db, desc_len=61)
        print("match: mean best %.1f" % t["best"].mean())
    print("launches", eng.launch_count)
    eng.close()


if __name__ == "__main__":
    main()
