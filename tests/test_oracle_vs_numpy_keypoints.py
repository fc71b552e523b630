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
