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
