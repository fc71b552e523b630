"""CPU-only: the C oracle's keypoint stages (cache pass, upper-scale filter, sub-pixel step, orientation, M-LDB) against a
second restatement written independently from the Rust sources in Python (tests/np_keypoints.py), on the oracle's own
evolution images. Together with tests/test_oracle_vs_numpy.py (stencil stages) every stage of the oracle is pinned by two
independently written implementations that agree bit for bit."""
import numpy as np
import pytest

import np_keypoints as K
import np_restatement as R


def _images(akz):
    import os
    yield "synthetic 264x200", R.synthetic_image(200, 264, seed=5)
    g = akz.load_gray(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "1.jpg"))
    yield "crop of the reference's test-data/1.jpg", np.ascontiguousarray(g[300:780, 300:940])


@pytest.mark.parametrize("which", [0, 1])
def test_keypoint_stages_match_independent_restatement(oracle, akz, which):
    _name, img = list(_images(akz))[which]
    ref = oracle.extract(oracle.unit_float_from_u8(img))
    assert ref.status == 0 and len(ref.keypoints) > 150 and ref.num_candidates > 2 * ref.num_cache
    ldets = [ref.image(l, "Ldet") for l in range(ref.num_levels)]
    cache_out, n_cache = K.find_scale_space_extrema(ldets, ref.levels)
    assert n_cache == ref.num_cache
    kps = K.do_subpixel_refinement(cache_out, ldets)
    assert len(kps) == len(ref.keypoints)
    rk = ref.keypoints
    for name in ("x", "y", "response", "size"):
        assert np.array_equal(np.array([k[name] for k in kps], np.float32), rk[name]), name
    assert np.array_equal(np.array([k["octave"] for k in kps]), rk["octave"])
    assert np.array_equal(np.array([k["class_id"] for k in kps]), rk["class_id"])
    # orientation and descriptors on a subset (pure-Python loops): every 7th keypoint
    for i in range(0, len(kps), 7):
        kp = kps[i]
        lv = kp["class_id"]
        lx, ly, lt = ref.image(lv, "Lx"), ref.image(lv, "Ly"), ref.image(lv, "Lt")
        ang = K.compute_main_orientation(kp, lx, ly, kp["octave"])
        assert ang == rk["angle"][i], (i, float(ang), float(rk["angle"][i]))
        kp["angle"] = ang
        d = K.mldb_descriptor(kp, lt, lx, ly)
        assert np.array_equal(d, ref.descriptors[i]), i


def _descriptor_match_literal(d0, d1, distance_threshold, lowes_ratio):
    """feature_matching.rs:23-123 taken literally, bail-out included (it may return a PARTIAL distance that is larger
    than the running second minimum; the update rule then ignores it, which is why the GPU may use full distances)."""
    pop = np.array([bin(i).count("1") for i in range(256)])
    out = []
    for i, a in enumerate(d0):
        mn, mj, sec = distance_threshold, 0, distance_threshold
        for j, b in enumerate(d1):
            dist = 0
            for x0, x1 in zip(a, b):
                dist += pop[x0 ^ x1]
                if dist > sec:
                    break
            if dist < mn:
                sec, mn, mj = mn, dist, j
            elif dist < sec:
                sec = dist
        if float(mn) < float(sec) * lowes_ratio ** 2 and mn < distance_threshold:
            out.append((i, mj, float(mn)))
    return out


def test_descriptor_match_matches_independent_restatement(oracle):
    import os
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    a = np.load(os.path.join(g, "features_1.npz"))["descriptors"][:60]
    b = np.load(os.path.join(g, "features_2.npz"))["descriptors"][:400]
    rng = np.random.default_rng(3)
    b = b.copy()
    b[rng.integers(0, len(b), 25)] = a[rng.integers(0, len(a), 25)]  # exact and near duplicates: ties on the minimum
    b[7] = b[3]
    for lowes in (0.86, 1.5):
        m = oracle.descriptor_match(a, b, 10000, lowes)
        lit = _descriptor_match_literal(a.tolist(), b.tolist(), 10000, lowes)
        assert [(int(r["index_0"]), int(r["index_1"]), float(r["distance"])) for r in m] == lit
    # raw top-2 against a vectorised numpy brute force (lowest index wins ties)
    d = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(axis=2)
    bi, best, second = oracle.match_top2(a, b)
    assert np.array_equal(bi, d.argmin(axis=1)) and np.array_equal(best, d.min(axis=1))
    assert np.array_equal(second, np.sort(d, axis=1)[:, 1])


@pytest.mark.parametrize("name", ["1", "2"])
def test_reference_test_images_full_size(oracle, akz, name):
    """The reference's own test-data/1.jpg and 2.jpg at full size: the independent restatement reproduces the committed
    golden keypoints (all 7 395 / 5 629, every field) from the oracle's Ldet images -- including the cases where the
    upper-scale filter and the sub-pixel test drop cache entries -- and angles + descriptors of a sample."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gray = akz.load_gray(os.path.join(gold, name + ".jpg"))
    ref = oracle.extract(oracle.unit_float_from_u8(gray), threads=4)
    g = np.load(os.path.join(gold, "features_%s.npz" % name))
    ldets = [ref.image(l, "Ldet") for l in range(ref.num_levels)]
    cache_out, n_cache = K.find_scale_space_extrema_vec(ldets, ref.levels)
    assert n_cache == ref.num_cache
    kps = K.do_subpixel_refinement(cache_out, ldets)
    gk = g["keypoints"]
    assert len(kps) == len(gk) and len(kps) <= n_cache  # 2.jpg: 5 630 cache entries, 5 629 survive the upper-scale filter
    for f in ("x", "y", "response", "size"):
        assert np.array_equal(np.array([k[f] for k in kps], np.float32), gk[f]), f
    assert np.array_equal(np.array([k["class_id"] for k in kps]), gk["class_id"])
    planes = {}
    for i in range(0, len(kps), 97):
        kp = kps[i]
        lv = kp["class_id"]
        if lv not in planes:
            planes[lv] = (ref.image(lv, "Lx"), ref.image(lv, "Ly"), ref.image(lv, "Lt"))
        lx, ly, lt = planes[lv]
        kp["angle"] = K.compute_main_orientation(kp, lx, ly, kp["octave"])
        assert kp["angle"] == gk["angle"][i], i
        assert np.array_equal(K.mldb_descriptor(kp, lt, lx, ly), g["descriptors"][i]), i
