"""The C++ host mirror (include/akaze_b200.hpp: the crate's names on top of the C ABI, RANSAC, akaze-util's
bincode formats) and the Python formats module (SURVEY.md section 8 rows b and f-1)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    """host_mirror_check built with g++ against the in-tree shared library."""
    import __graft_entry__ as G
    G.build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_mirror_check")
    libdir = os.path.join(ROOT, "akaze-rust_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", out, "-L" + libdir,
           "-lakaze_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def _fixed_features():
    import akaze_rust_b200  # noqa: F401
    from akaze_rust_b200 import formats as F
    k = np.zeros(5, F.KP_BIN)
    d = np.zeros((5, 61), np.uint8)
    for i in range(5):
        k[i] = (np.float32(10.5) + np.float32(i), np.float32(20.25) * np.float32(i), np.float32(0.001) * np.float32(i + 1),
                np.float32(2.4) * np.float32(1 << (i % 3)), i % 3, 4 * (i % 3) + 1, np.float32(0.1) * np.float32(i) - np.float32(0.2))
        d[i] = [(i * 37 + j * 11) & 255 for j in range(61)]
    return F, k, d


def test_bincode_layout_matches_cpp_writer(checker, tmp_path):
    F, k, d = _fixed_features()
    r = subprocess.run([checker, "formats", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp = open(tmp_path / "features.bin", "rb").read()
    assert len(cpp) == 8 + 5 * 36 + 8 + 5 * (8 + 61)
    assert cpp == F.features_to_bytes(k, d, 61)
    m = np.array([(0, 3, 12.0), (1, 1, 0.0), (4, 2, 77.0)], F.MATCH_BIN)
    assert open(tmp_path / "matches.bin", "rb").read() == F.matches_to_bytes(m)
    k2, d2 = F.deserialize_features_from_file(tmp_path / "features.bin")
    assert np.array_equal(k2, k) and all(np.array_equal(a, b) for a, b in zip(d2, d))
    assert np.array_equal(F.deserialize_matches_from_file(tmp_path / "matches.bin"), m)
    # RANSAC of the C++ mirror on a synthetic two-view scene: 40 inliers, 10 gross outliers
    kept, total = [int(x) for x in r.stdout.split()[1:3]]
    assert total == 50 and 35 <= kept <= 40


def test_json_and_bincode_round_trip(tmp_path):
    F, k, d = _fixed_features()
    for name in ("f.json", "f.bin", "f"):
        F.serialize_features_to_file(k, d, tmp_path / name)
        k2, d2 = F.deserialize_features_from_file(tmp_path / name)
        assert np.array_equal(k2, k) and all(np.array_equal(a, b) for a, b in zip(d2, d)), name
    m = np.array([(0, 3, 12.0), (7, 1, 0.5)], F.MATCH_BIN)
    for name in ("m.json", "m.bin"):
        F.serialize_matches_to_file(m, tmp_path / name)
        assert np.array_equal(F.deserialize_matches_from_file(tmp_path / name), m)
    import json
    doc = json.load(open(tmp_path / "f.json"))
    assert set(doc) == {"keypoints", "descriptors"} and set(doc["keypoints"][0]) == {"point", "response", "size", "octave", "class_id", "angle"}
    assert doc["keypoints"][1]["point"] == [11.5, 20.25] and doc["descriptors"][0] == {"vector": [int(v) for v in d[0]]}


@pytest.mark.gpu
def test_cpp_mirror_matches_python_mirror(checker, akz, tmp_path):
    """extract_features / descriptor_match / match_features through the C++ mirror give what the Python mirror gives."""
    import np_restatement as R
    from akaze_rust_b200 import formats as F
    w, h = 640, 480
    img0 = R.natural_image(h, w, 501)
    img1 = np.roll(img0, (3, 7), axis=(0, 1))
    img0.tofile(tmp_path / "i0.u8")
    img1.tofile(tmp_path / "i1.u8")
    r = subprocess.run([checker, "gpu", str(w), str(h), str(tmp_path / "i0.u8"), str(tmp_path / "i1.u8"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    eng = akz.Engine(0, w, h, 1)
    f0, f1 = eng.extract_u8(img0), eng.extract_u8(img1)
    for name, f in (("features0.bin", f0), ("features1.bin", f1)):
        k, d = F.deserialize_features_from_file(tmp_path / name)
        assert len(k) == len(f.keypoints) > 100
        for fld in ("x", "y", "response", "size", "octave", "class_id", "angle"):
            assert np.array_equal(k[fld], f.keypoints[fld].astype(k[fld].dtype)), fld
        assert np.array_equal(np.stack(d), f.descriptors[:, :61])
    dm = F.deserialize_matches_from_file(tmp_path / "descriptor_matches.bin")
    ref = eng.descriptor_match(f0.descriptors, f1.descriptors, 10000, 0.86, desc_len=61)
    assert len(dm) > 20 and np.array_equal(dm["index_0"], ref["index_0"]) and np.array_equal(dm["index_1"], ref["index_1"])
    assert np.array_equal(dm["distance"], ref["distance"])
    mf = F.deserialize_matches_from_file(tmp_path / "matches.bin")
    assert 0 < len(mf) <= len(dm)
    assert set(zip(mf["index_0"], mf["index_1"])).issubset(set(zip(dm["index_0"], dm["index_1"])))
    # GPU RANSAC (akz_remove_outliers): identical to the host mirror under the reference's sampling, at least as
    # many inliers with 1000 distinct hypotheses, and every kept match is an inlier of the returned model
    line = [l for l in r.stdout.splitlines() if l.startswith("ransac host")][0].split()
    host_n, gpu_n, same, adv_n, consistent = int(line[2]), int(line[4]), int(line[6]), int(line[8]), int(line[10])
    assert same == 1 and host_n == gpu_n == len(mf)
    assert adv_n >= gpu_n and consistent == adv_n
    adv = F.deserialize_matches_from_file(tmp_path / "matches_ransac_advancing.bin")
    assert set(zip(adv["index_0"], adv["index_1"])).issubset(set(zip(dm["index_0"], dm["index_1"])))
    # the same through the Python layer
    g = eng.remove_outliers(f0.keypoints, f1.keypoints, ref, 1000, 0.05, 3.0, sampling="reference")
    assert np.array_equal(g["index_0"], mf["index_0"]) and np.array_equal(g["index_1"], mf["index_1"])
    g2, model = eng.remove_outliers(f0.keypoints, f1.keypoints, ref, 1000, 0.05, 3.0, sampling="advancing", return_model=True)
    assert np.array_equal(g2["index_0"], adv["index_0"]) and model.shape == (3, 3) and np.isfinite(model).all()
    few = eng.remove_outliers(f0.keypoints, f1.keypoints, ref[:5], 1000, 0.05, 3.0)
    assert np.array_equal(few, ref[:5])  # fewer than 8 matches come back untouched (estimate_fundamental_matrix.rs:107-110)
    eng.close()
