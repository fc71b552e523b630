"""Mirrors akaze/tests/integration-test.rs (the reference's own tests for this path) on the drop-in API."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEST_DATA = os.path.join(ROOT, "tests", "golden")


def test_locate_data():
    # integration-test.rs:22-38
    for n in ("1.jpg", "2.jpg"):
        assert os.path.exists(os.path.join(TEST_DATA, n))


def test_extract_features(akz):
    # integration-test.rs:40-70: extract_features(test-data/1.jpg, Config::default())
    options = akz.Config.default()
    evolutions, keypoints, descriptors = akz.extract_features(os.path.join(TEST_DATA, "1.jpg"), options)
    assert len(evolutions) == 16 and len(keypoints) == len(descriptors) == 7395
    assert descriptors.shape[1] == 61
    # the optional scale-space dump of the reference test reads every image of every evolution
    lt = evolutions[15].Lt
    assert lt.shape == (189, 252) and np.isfinite(lt).all()


def test_match_features(akz):
    # integration-test.rs:72-123: match_features(kp0, d0, kp1, d1, 0.86, 1000, 3.0)
    options = akz.Config.default()
    _e0, keypoints_0, descriptors_0 = akz.extract_features(os.path.join(TEST_DATA, "1.jpg"), options)
    _e1, keypoints_1, descriptors_1 = akz.extract_features(os.path.join(TEST_DATA, "2.jpg"), options)
    matches = akz.match_features(keypoints_0, descriptors_0, keypoints_1, descriptors_1, 0.86, 1000, 3.0)
    putative = akz.descriptor_match(descriptors_0, descriptors_1, 10000, 0.86)
    assert 550 <= len(putative) <= 565          # oracle: 557
    assert 0 < len(matches) <= len(putative)
    assert set(zip(matches["index_0"], matches["index_1"])).issubset(set(zip(putative["index_0"], putative["index_1"])))


def test_command_line_tools(akz, tmp_path):
    # akaze-util/src/bin/extract_and_match.rs + match_features.rs on the engine (akaze-rust_b200/cli.py)
    from akaze_rust_b200 import cli, formats
    prefix = str(tmp_path / "run")
    assert cli.main(["extract_and_match", os.path.join(TEST_DATA, "1.jpg"), os.path.join(TEST_DATA, "2.jpg"), prefix]) == 0
    k0, d0 = formats.deserialize_features_from_file(prefix + "-extractions_0.cbor")
    k1, d1 = formats.deserialize_features_from_file(prefix + "-extractions_1.cbor")
    assert (len(k0), len(k1)) == (7395, 5629) and len(d0[0]) == 61
    m = formats.deserialize_matches_from_file(prefix + "-matches.cbor")
    assert 0 < len(m) <= 565
    out = str(tmp_path / "m.json")
    assert cli.main(["match_features", prefix + "-extractions_0.cbor", prefix + "-extractions_1.cbor", out]) == 0
    assert 0 < len(formats.deserialize_matches_from_file(out)) <= 565
