"""Pins the CPU oracle against every known-answer vector the reference holds for the hot path, plus
structural invariants of the restated scalar code. (CPU only.)"""
import numpy as np


def test_gaussian_kernel_known_answer(oracle):
    # akaze/src/types/image.rs:486-502 `gaussian_kernel_correct`: the only asserted numeric test of the reference
    known = [0.10628852, 0.14032133, 0.16577007, 0.17524014, 0.16577007, 0.14032133, 0.10628852]
    k = oracle.gaussian_kernel(3.0, 7)
    assert np.all(np.abs(k - np.array(known, np.float32)) < 1e-4)
    assert abs(float(k.sum()) - 1.0) < 1e-6


def test_scharr_taps_known_answer(oracle):
    # akaze/src/ops/derivatives.rs:12,22 (stated there, not asserted)
    assert np.array_equal(oracle.scharr_main_axis_kernel(1), np.array([0.09375, 0.3125, 0.09375], np.float32))
    assert np.array_equal(oracle.scharr_off_axis_kernel(1), np.array([-1, 0, 1], np.float32))
    for s in (2, 3, 4):
        m = oracle.scharr_main_axis_kernel(s)
        o = oracle.scharr_off_axis_kernel(s)
        assert len(m) == len(o) == 2 * s + 1  # derivatives.rs:9,19 check only the length
        assert np.count_nonzero(m) == 3 and np.count_nonzero(o) == 2
        assert abs(float(m.sum()) * 2 * s - 1.0) < 1e-6


def test_default_config(oracle):
    c = oracle.default_config()  # evolution.rs:41-54
    assert (c.num_sublevels, c.max_octave_evolution) == (4, 4)
    assert (c.base_scale_offset, c.initial_contrast, c.contrast_percentile) == (1.6, 0.001, 0.7)
    assert (c.contrast_factor_num_bins, c.derivative_factor, c.detector_threshold) == (300, 1.5, 0.001)
    assert (c.descriptor_channels, c.descriptor_pattern_size) == (3, 10)


def test_fed_schedule_invariants(oracle):
    # sum of the FED steps equals the requested process time; reordering is a permutation
    esig = [1.6 * 2.0 ** (s / 4.0 + o) for o in range(4) for s in range(4)]
    et = [0.5 * e * e for e in esig]
    ns = []
    for i in range(1, 16):
        T = et[i] - et[i - 1]
        tau = oracle.fed_tau_by_process_time(T, 1, 0.25, True)
        plain = oracle.fed_tau_by_process_time(T, 1, 0.25, False)
        ns.append(len(tau))
        assert abs(tau.sum() - T) < 1e-9 * max(1.0, T)
        assert np.array_equal(np.sort(tau), np.sort(plain))
    # SURVEY.md appendix A (derived independently during the survey)
    assert ns == [3, 3, 4, 4, 5, 6, 7, 8, 10, 12, 14, 17, 20, 24, 29]
    t3 = oracle.fed_tau_by_process_time(et[3] - et[2], 1, 0.25, True)
    assert np.allclose(t3, [0.106, 0.680, 0.082, 0.192], atol=1e-3)


def test_level_table(oracle):
    import np_restatement as R
    img = R.synthetic_image(96, 200, 1)
    r = oracle.extract(oracle.unit_float_from_u8(img), stop_after=1)
    # 200x96: octave 1 is 100x48 (>=80x40), octave 2 would be 50x24 -> 8 levels (evolution.rs:138-149)
    assert r.num_levels == 8
    assert [(l["width"], l["height"]) for l in r.levels] == [(200, 96)] * 4 + [(100, 48)] * 4
    assert [l["sigma_size"] for l in r.levels] == [2, 2, 2, 3, 3, 4, 5, 5]
    for l in r.levels:
        assert l["esigma"] == 1.6 * 2.0 ** (l["sublevel"] / 4.0 + l["octave"])


def test_descriptor_shape(oracle):
    import np_restatement as R
    img = R.synthetic_image(200, 260, 2)
    r = oracle.extract(oracle.unit_float_from_u8(img))
    assert r.status == 0 and len(r.keypoints) > 20
    assert r.descriptor_len == 61  # 486 bits (descriptors.rs:42-46)
    assert np.all(r.descriptors[:, 60] < 64)  # bits 486,487 never set
    kp = r.keypoints
    assert np.all((kp["angle"] >= 0) & (kp["angle"] < np.float32(np.pi)))
    assert np.all(kp["size"] == (np.float32(1.5) * 0 + kp["size"]))
    assert np.all(kp["octave"] == kp["class_id"] // 4)


def test_matcher_semantics(oracle):
    # feature_matching.rs:37-50: two smallest of {d_j} U {10000,10000}, lowest j wins
    q = np.zeros((2, 61), np.uint8)
    q[1, 0] = 0xFF
    db = np.zeros((4, 61), np.uint8)
    db[0, 0] = 0x0F   # d(q0)=4  d(q1)=4
    db[1, 0] = 0x01   # d(q0)=1  d(q1)=7
    db[2, 0] = 0x80   # d(q0)=1  d(q1)=7
    db[3, 0] = 0xFF   # d(q0)=8  d(q1)=0
    bi, b, s = oracle.match_top2(q, db)
    assert list(bi) == [1, 3] and list(b) == [1, 0] and list(s) == [1, 4]
    bi, b, s = oracle.match_top2(q, db[:0])
    assert list(bi) == [0, 0] and list(b) == [10000, 10000] and list(s) == [10000, 10000]
    bi, b, s = oracle.match_top2(q, db[:1])
    assert list(b) == [4, 4] and list(s) == [10000, 10000]
    m = oracle.descriptor_match(q, db, 10000, 0.86)
    # q0: 1 < 1*0.7396 false -> rejected; q1: 0 < 4*0.7396 -> accepted
    assert len(m) == 1 and m[0]["index_0"] == 1 and m[0]["index_1"] == 3 and m[0]["distance"] == 0.0
