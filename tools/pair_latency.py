#!/usr/bin/env python
"""configs[1] analogue: extract_and_match on one image pair, everything through the public host API (images in host memory in,
inlier matches out), per step and in total:  python tools/pair_latency.py [HxW]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
ak = importlib.import_module("akaze-rust_b200")
from np_restatement import natural_image  # noqa: E402
ransac_host = importlib.import_module("akaze-rust_b200.ransac")


def main():
    h, w = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "750x1000").split("x"))
    a = natural_image(h, w + 40, 77)
    img0, img1 = np.ascontiguousarray(a[:, :w]), np.ascontiguousarray(a[:, 25:25 + w])  # the second view: 25 px to the side
    eng = ak.Engine(0, w, h, 1)
    out = {}
    for ransac in ("host", "gpu"):
        rows = []
        for rep in range(12):
            t = [time.perf_counter()]
            f0 = eng.extract_u8(img0)
            k0, d0 = f0.keypoints, f0.descriptors
            t.append(time.perf_counter())
            f1 = eng.extract_u8(img1)
            k1, d1 = f1.keypoints, f1.descriptors
            t.append(time.perf_counter())
            put = eng.descriptor_match(d0, d1, 10000, 0.8)
            t.append(time.perf_counter())
            if ransac == "gpu":
                inl = eng.remove_outliers(k0, k1, put, 1000, 0.05, 3.0)
            else:
                inl = ransac_host.remove_outliers(k0, k1, put, 1000, 0.05, 3.0)
            t.append(time.perf_counter())
            f0.release()
            f1.release()
            if rep >= 2:
                rows.append([(t[i + 1] - t[i]) * 1e3 for i in range(4)] + [(t[4] - t[0]) * 1e3])
        m = np.median(np.array(rows), axis=0)
        out[ransac] = {"extract_0_ms": round(m[0], 3), "extract_1_ms": round(m[1], 3), "descriptor_match_ms": round(m[2], 3),
                       "remove_outliers_1000_trials_ms": round(m[3], 3), "total_ms": round(m[4], 3), "keypoints": [len(k0), len(k1)],
                       "putative": len(put), "inliers": len(inl)}
    # the same pair with both images in ONE extraction call (akz_extract_batch_u8): the kernels of a single image leave most of
    # the GPU idle, a second image rides along
    eng2 = ak.Engine(0, w, h, 2)
    pair = np.stack([img0, img1])
    rows = []
    for rep in range(12):
        t0 = time.perf_counter()
        f0, f1 = eng2.extract_batch_u8(pair)
        k0, d0, k1, d1 = f0.keypoints, f0.descriptors, f1.keypoints, f1.descriptors
        t1 = time.perf_counter()
        put = eng2.descriptor_match(d0, d1, 10000, 0.8)
        inl = eng2.remove_outliers(k0, k1, put, 1000, 0.05, 3.0)
        t2 = time.perf_counter()
        f0.release()
        f1.release()
        if rep >= 2:
            rows.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3])
    m = np.median(np.array(rows), axis=0)
    out["gpu, both images in one extraction call"] = {"extract_pair_ms": round(m[0], 3), "match_and_ransac_ms": round(m[1], 3), "total_ms": round(m[2], 3),
                                                      "inliers": len(inl)}
    eng2.close()
    eng.close()
    print(json.dumps({"shape": "%dx%d" % (h, w), "ransac": out}, indent=1))


if __name__ == "__main__":
    main()
