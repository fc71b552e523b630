"""Stage-by-stage GPU-vs-oracle parity report (debug aid; the gating tests live in tests/).
Usage: python tools/parity_report.py [--fixture] [--size HxW] -> prints a table, writes gpurun_out/parity.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import akaze_rust_b200 as A  # noqa: E402
from oracle import akaze_oracle as O  # noqa: E402
import np_restatement as R  # noqa: E402


def cmp_img(a, b):
    if a is None or b is None:
        return {"missing": True}
    eq = bool(np.array_equal(a, b))
    d = {"equal": eq}
    if not eq:
        diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
        bad = np.argwhere(a != b)
        d.update(max_abs=float(diff.max()), n_diff=int(len(bad)), first=[int(v) for v in bad[0]],
                 last=[int(v) for v in bad[-1]], a=float(a[tuple(bad[0])]), b=float(b[tuple(bad[0])]),
                 nan_a=int(np.isnan(a).sum()))
    return d


def report(name, img8, eng, out):
    t0 = time.time()
    f = eng.extract_u8(img8)
    t_gpu = time.time() - t0
    t0 = time.time()
    ref = O.extract(O.unit_float_from_u8(img8), threads=8)
    t_cpu = time.time() - t0
    rep = {"shape": list(img8.shape), "t_gpu_s": t_gpu, "t_cpu_s": t_cpu, "levels": []}
    rep["contrast"] = [f.contrast_factor, ref.contrast_factor, f.contrast_factor == ref.contrast_factor]
    rep["n_levels"] = [len(f.evolutions), ref.num_levels]
    for lv in range(min(len(f.evolutions), ref.num_levels)):
        row = {"level": lv, "wh": [f.evolutions[lv].width, f.evolutions[lv].height],
               "tau_eq": bool(np.array_equal(f.evolutions[lv].fed_tau_steps, ref.levels[lv]["fed_tau_steps"]))}
        for kind in ("Lt", "Lsmooth", "Lflow", "Lx", "Ly", "Lxx", "Lyy", "Lxy", "Ldet", "Lstep"):
            if lv == 0 and kind in ("Lflow", "Lstep"):
                continue
            try:
                g = f.evolution(lv, kind)
            except Exception as e:  # noqa: BLE001
                row[kind] = {"error": str(e)}
                continue
            row[kind] = cmp_img(g, ref.image(lv, kind))
        rep["levels"].append(row)
    rep["candidates"] = [int(f.num_candidates), int(ref.num_candidates)]
    rep["cache"] = [int(f.num_cache), int(ref.num_cache)]
    rep["keypoints"] = [len(f.keypoints), len(ref.keypoints)]
    n = min(len(f.keypoints), len(ref.keypoints))
    if n:
        kg, kr = f.keypoints[:n], ref.keypoints[:n]
        rep["kp_fields_equal"] = {k: int((kg[k] == kr[k]).sum()) for k in kg.dtype.names}
        rep["kp_angle_maxdiff"] = float(np.abs(kg["angle"] - kr["angle"]).max())
        same = np.all([kg[k] == kr[k] for k in ("x", "y", "class_id")], axis=0)
        if same.any():
            x = np.unpackbits(f.descriptors[:n][same] ^ ref.descriptors[:n][same], axis=1)
            rep["desc_bits_diff"] = int(x.sum())
            rep["desc_bits_total"] = int(x.size)
            rep["desc_rows_equal"] = int((x.sum(axis=1) == 0).sum())
            rep["desc_rows"] = int(same.sum())
        if not same.all():
            i = int(np.argmin(same))
            rep["first_kp_mismatch"] = {"i": i, "gpu": [float(v) for v in kg[i].tolist()], "ref": [float(v) for v in kr[i].tolist()]}
    out[name] = rep
    # compact print
    print("== %s %s  gpu %.3fs cpu %.3fs" % (name, img8.shape, t_gpu, t_cpu))
    print("contrast", rep["contrast"], "cand", rep["candidates"], "cache", rep["cache"], "kp", rep["keypoints"])
    for row in rep["levels"]:
        bad = {k: v for k, v in row.items() if isinstance(v, dict) and not v.get("equal", False)}
        print(" L%-2d %s tau_eq=%s %s" % (row["level"], row["wh"], row["tau_eq"], "ALL EQUAL" if not bad else json.dumps(bad)))
    for k in ("kp_fields_equal", "kp_angle_maxdiff", "desc_bits_diff", "desc_bits_total", "desc_rows_equal", "desc_rows", "first_kp_mismatch"):
        if k in rep:
            print(" ", k, rep[k])
    return f, ref


def matcher_report(eng, out):
    rng = np.random.default_rng(1)
    res = {}
    for nq, ndb in ((1, 1), (5, 3), (300, 1000), (1000, 257), (7395, 5629), (513, 70000)):
        q = rng.integers(0, 256, (nq, 64), dtype=np.uint8)
        db = rng.integers(0, 256, (ndb, 64), dtype=np.uint8)
        q[:, 61:] = 0
        db[:, 61:] = 0
        q[:, 60] &= 0x3F
        db[:, 60] &= 0x3F
        k = min(nq, ndb) // 2
        db[:k] = q[:k]  # exact duplicates -> distance-0 ties on index
        if ndb > 4:
            db[3] = db[1]
        t = eng.match_top2(q, db, desc_len=61)
        bi, b, s = O.match_top2(q, db, desc_len=61)
        ok = bool(np.array_equal(t["best_idx"], bi) and np.array_equal(t["best"], b) and np.array_equal(t["second"], s))
        res["%dx%d" % (nq, ndb)] = ok
        print("match %dx%d exact=%s" % (nq, ndb, ok))
        if not ok:
            bad = np.flatnonzero((t["best_idx"] != bi) | (t["best"] != b) | (t["second"] != s))
            i = bad[0]
            print("   first bad", i, t[i], bi[i], b[i], s[i], "n_bad", len(bad))
    out["matcher"] = res


def main():
    out = {}
    eng = A.Engine(0, 4096, 4096, 2, keep_evolutions=True)
    matcher_report(eng, out)
    sizes = [(240, 320), (135, 333)]
    for a in sys.argv[1:]:
        if "x" in a and a[0].isdigit():
            sizes = [tuple(int(v) for v in a.split("x"))]
    for i, (h, w) in enumerate(sizes):
        report("synthetic_%dx%d" % (h, w), R.synthetic_image(h, w, seed=11 + i), eng, out)
    if "--fixture" in sys.argv:
        g1 = A.load_gray(os.path.join(ROOT, "tests", "golden", "1.jpg"))
        f1, r1 = report("fixture_1", g1, eng, out)
        g2 = A.load_gray(os.path.join(ROOT, "tests", "golden", "2.jpg"))
        f2, r2 = report("fixture_2", g2, eng, out)
        t0 = time.time()
        m = eng.descriptor_match(f1.descriptors_padded, f2.descriptors_padded, lowes_ratio=0.86, desc_len=61)
        tg = time.time() - t0
        t0 = time.time()
        mo = O.descriptor_match(r1.descriptors, r2.descriptors, lowes_ratio=0.86)
        tc = time.time() - t0
        print("fixture match: gpu %d (%.3fs) oracle %d (%.3fs)" % (len(m), tg, len(mo), tc))
        mg = eng.descriptor_match(r1.descriptors, r2.descriptors, lowes_ratio=0.86)
        print("  same-descriptor match identical:", bool(np.array_equal(mg, mo)))
        out["fixture_match"] = [len(m), len(mo), bool(np.array_equal(mg, mo))]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
