#!/usr/bin/env python
"""Quick A/B of kernel variants on the GPU box: runs one 1080p batch per variant (each in a fresh process, the
A/B switches are environment variables read once), prints the serialised per-stage times (ms per image, CUDA
events inside the library), whole-batch throughput and a checksum of keypoints + descriptors, so that variants
can be compared for speed AND bit-equality in one visit.

    python tools/stage_ab.py --images 256 -- "" "AKZ_FED_OLD=1" "AKZ_FED_CAP=1"
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(args):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch
    import akaze_rust_b200 as A
    import np_restatement as R
    W, H = args.width, args.height
    cache = os.path.join(ROOT, "gpurun_out", "_ab_imgs_%dx%d_%d.npy" % (W, H, args.unique))
    if os.path.exists(cache):
        uniq = np.load(cache)
    else:
        uniq = np.stack([R.natural_image(H, W, 1000 + i) for i in range(args.unique)])
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.save(cache, uniq)
    n = args.images
    d = torch.from_numpy(uniq).cuda()
    d_imgs = d[torch.arange(n, device="cuda") % args.unique].contiguous()
    eng = A.Engine(0, W, H, n)
    cfg = A.Config.default()
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda:0"))
    for _ in range(2):
        counts = eng.extract_batch_u8_device(d_imgs.data_ptr(), n, W, H, W, cfg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.reps):
        counts = eng.extract_batch_u8_device(d_imgs.data_ptr(), n, W, H, W, cfg)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    eng.enable_timing(True)
    eng.stage_times(reset=True)
    eng.extract_batch_u8_device(d_imgs.data_ptr(), n, W, H, W, cfg)
    st = eng.stage_times(reset=True)
    eng.enable_timing(False)
    # checksum over the first `unique` images through the host path
    fs = eng.extract_batch_u8([uniq[i] for i in range(args.unique)], cfg)
    h = hashlib.sha256()
    for f in fs:
        h.update(np.ascontiguousarray(f.keypoints).tobytes())
        h.update(np.ascontiguousarray(f.descriptors).tobytes())
        f.release()
    print(json.dumps({"ips": n / (ms / 1e3), "stages": {k: round(v[0] / n, 5) for k, v in st.items()},
                      "kp": float(np.mean(counts)), "sha": h.hexdigest()[:16]}))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--unique", type=int, default=4)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("variants", nargs="*")
    args = ap.parse_args()
    if args.child:
        return child(args)
    variants = args.variants or [""]
    for v in variants:
        env = dict(os.environ)
        for kv in v.split():
            k, _, val = kv.partition("=")
            env[k] = val
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--images", str(args.images), "--unique", str(args.unique),
               "--width", str(args.width), "--height", str(args.height), "--reps", str(args.reps)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("FAILED: " + r.stderr[-800:])
        print("[%s] %s" % (v or "default", out), flush=True)


if __name__ == "__main__":
    main()
