#!/usr/bin/env python
"""What the opt-in fast-math build (AKZ_FAST_MATH=1: fused multiply-adds in the stencil kernels) costs in parity and buys
in speed, next to the default bit-exact build. For each build, in a fresh process: N of bench.py's 1080p images through a
default-mode engine; against the CPU oracle: max-abs / max-rel error of Lt and Ldet per level, keypoint agreement by the
north-star rule (an oracle keypoint counts if the engine has one within 0.5 px in the same octave), descriptor bits on
the agreeing keypoints; and device-resident images/s. VERDICT r1 item 5b.

    python tools/fast_math_report.py [--images 8] [--time-images 256] > gpurun_out/fast_math.json
"""
import argparse
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
    from scipy.spatial import cKDTree
    import __graft_entry__ as G
    G.build()
    import akaze_rust_b200 as A
    from oracle import akaze_oracle as O
    import np_restatement as R
    imgs = [R.natural_image(1080, 1920, 1000 + i) for i in range(args.images)]
    eng = A.Engine(0, 1920, 1080, max(args.images, args.time_images))
    fs = eng.extract_batch_u8(np.stack(imgs))
    n_levels = 16
    err = {k: {"max_abs": [0.0] * n_levels, "max_rel": [0.0] * n_levels} for k in ("Lt", "Ldet")}
    tot = {"oracle_kp": 0, "engine_kp": 0, "agree": 0, "identical": 0, "bits": 0, "bits_diff": 0, "max_angle_diff": 0.0}
    # the planes of the last two sub-batches are resident: compare those images' planes, all images' features
    for i, (img, f) in enumerate(zip(imgs, fs)):
        ref = O.extract(O.unit_float_from_u8(img), threads=8)
        kg, kr = f.keypoints, ref.keypoints
        tot["oracle_kp"] += len(kr)
        tot["engine_kp"] += len(kg)
        if len(kg) and len(kr):
            # nearest engine keypoint of the SAME class (several classes can hold a keypoint at one position): the class is
            # folded into the tree as a third coordinate far larger than any image
            t = cKDTree(np.stack([kg["x"], kg["y"], kg["class_id"] * 1.0e5], axis=1))
            d, j = t.query(np.stack([kr["x"], kr["y"], kr["class_id"] * 1.0e5], axis=1))
            ok = (d <= 0.5) & (kg["octave"][j] == kr["octave"])
            tot["agree"] += int(ok.sum())
            same = ok & (kg["x"][j] == kr["x"]) & (kg["y"][j] == kr["y"]) & (kg["response"][j] == kr["response"])
            tot["identical"] += int(same.sum())
            x = np.unpackbits(f.descriptors[j[ok]] ^ ref.descriptors[ok], axis=1)
            tot["bits"] += int(x.size)
            tot["bits_diff"] += int(x.sum())
            tot["max_angle_diff"] = max(tot["max_angle_diff"], float(np.abs(kg["angle"][j[ok]] - kr["angle"][ok]).max()))
        try:
            for lv in range(ref.num_levels):
                for kind in ("Lt", "Ldet"):
                    a, b = f.evolution(lv, kind).astype(np.float64), ref.image(lv, kind).astype(np.float64)
                    e = np.abs(a - b)
                    err[kind]["max_abs"][lv] = max(err[kind]["max_abs"][lv], float(e.max()))
                    err[kind]["max_rel"][lv] = max(err[kind]["max_rel"][lv], float((e / np.maximum(np.abs(b), 1e-3)).max()))
        except A.AkazeError:
            pass  # an earlier sub-batch: planes already overwritten
        ref.close()
    # throughput, device-resident
    n = args.time_images
    d = torch.from_numpy(np.stack(imgs)).cuda()
    d_imgs = d[torch.arange(n, device="cuda") % len(imgs)].contiguous()
    cfg = A.Config.default()
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda:0"))
    for _ in range(2):
        eng.extract_batch_u8_device(d_imgs.data_ptr(), n, 1920, 1080, 1920, cfg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(3):
        eng.extract_batch_u8_device(d_imgs.data_ptr(), n, 1920, 1080, 1920, cfg)
    e1.record(stream)
    torch.cuda.synchronize()
    ips = 3 * n / (e0.elapsed_time(e1) / 1e3)
    print(json.dumps({"fast_math": A.FAST_MATH, "images": args.images, "images_per_s": ips, "time_images": n,
                      "keypoint_agreement": tot["agree"] / max(1, tot["oracle_kp"]), "identical_keypoints": tot["identical"] / max(1, tot["oracle_kp"]),
                      "descriptor_bit_agreement": 1.0 - tot["bits_diff"] / max(1, tot["bits"]), "counts": tot, "evolution_error": err,
                      "tolerance": "north star / SURVEY 8(d): evolutions max-abs <= 1e-5, rel <= 1e-4 per level; >= 99 % keypoints within 0.5 px "
                                   "and the same octave; >= 99 % descriptor bits on those"}))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--time-images", type=int, default=256)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    out = {}
    for name, flag in (("exact", "0"), ("fast_math", "1")):
        env = dict(os.environ, AKZ_FAST_MATH=flag)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--images", str(args.images), "--time-images", str(args.time_images)],
                           env=env, capture_output=True, text=True)
        try:
            out[name] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:  # noqa: BLE001
            out[name] = {"failed": r.stderr[-1500:]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
