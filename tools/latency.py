#!/usr/bin/env python
"""Single-call latency of the extraction path (one image, host buffer in, host results out), with the per-stage device
times of the same call. Run on a GPU box:  python tools/latency.py [--sizes 1080x1920,2160x3840,600x800] [--reps 30]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1080x1920,2160x3840,600x800")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--pinned", action="store_true", help="the input images in page-locked host memory (torch)")
    args = ap.parse_args()
    ak = importlib.import_module("akaze-rust_b200")
    from np_restatement import natural_image
    out = {}
    for s in args.sizes.split(","):
        h, w = (int(v) for v in s.split("x"))
        imgs = np.stack([natural_image(h, w, 4242 + i) for i in range(args.batch)])
        if args.pinned:
            import torch
            keep = torch.from_numpy(imgs).pin_memory()
            imgs = keep.numpy()
        eng = ak.Engine(device=0, max_width=w, max_height=h, max_batch=args.batch)
        for _ in range(5):
            f = eng.extract_batch_u8(imgs)
            for x in f:
                x.keypoints
                x.descriptors_padded
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            f = eng.extract_batch_u8(imgs)
            kp = [x.keypoints for x in f]
            de = [x.descriptors_padded for x in f]
            ts.append((time.perf_counter() - t0) * 1e3)
        eng.enable_timing(True)
        eng.stage_times(reset=True)
        for _ in range(5):
            eng.extract_batch_u8(imgs)
        st = eng.stage_times(reset=True)
        eng.enable_timing(False)
        ts.sort()
        out[s] = {"batch": args.batch, "keypoints": int(sum(len(k) for k in kp)), "ms_median": round(ts[len(ts) // 2], 3), "ms_min": round(ts[0], 3),
                  "stage_ms_per_call": {k: round(v[0] / 5, 4) for k, v in st.items()}, "launches_per_call": int(sum(v[1] for v in st.values()) // 5)}
        eng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
