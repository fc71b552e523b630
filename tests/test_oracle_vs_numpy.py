"""The literal C oracle against the independent closed-form numpy restatement (tests/np_restatement.py).
Proves that the formulation the CUDA kernels use (valid-interior correlation + clamp borders, one flux
per edge, zero flux across the image border) is bit-identical to the reference's loop structure. (CPU only.)"""
import numpy as np
import pytest

import np_restatement as R


@pytest.mark.parametrize("shape", [(97, 131), (40, 80), (64, 64), (33, 201)])
def test_filters_closed_form(oracle, shape):
    img8 = R.synthetic_image(shape[0], shape[1], 7)
    u = oracle.unit_float_from_u8(img8)
    assert np.array_equal(u, (img8.astype(np.float32) * np.float32(1)) / np.float32(255))
    assert np.array_equal(oracle.gaussian_kernel(1.6, 5), R.gaussian_kernel(1.6, 5))
    assert np.array_equal(oracle.gaussian_kernel(1.0, 3), R.gaussian_kernel(1.0, 3))
    for r in (1.0, 1.6):
        assert np.array_equal(oracle.gaussian_blur(u, r), R.gaussian_blur(u, r))
    b = oracle.gaussian_blur(u, 1.6)
    for s in (1, 2, 3, 4):
        assert np.array_equal(oracle.scharr(b, 1, 0, s), R.scharr_x(b, s))
        assert np.array_equal(oracle.scharr(b, 0, 1, s), R.scharr_y(b, s))
    assert np.array_equal(oracle.half_size(b), R.half_size(b))


def test_contrast_g2_fed_closed_form(oracle):
    img8 = R.synthetic_image(90, 120, 5)
    b = oracle.gaussian_blur(oracle.unit_float_from_u8(img8), 1.6)
    k = oracle.compute_contrast_factor(b)
    assert k == R.contrast_factor(b)
    lx, ly = oracle.scharr(b, 1, 0, 1), oracle.scharr(b, 0, 1, 1)
    c = oracle.pm_g2(lx, ly, k)
    assert np.array_equal(c, R.pm_g2(lx, ly, k))
    L1, L2 = b.copy(), b.copy()
    for tau in [0.106, 0.68, 0.082, 0.19, 5.2, 41.3]:  # up to 165x the stability limit, like level 15
        L1, s1 = oracle.calculate_step(L1, c, tau)
        L2, s2 = R.fed_step(L2, c, tau)
        assert np.array_equal(L1, L2) and np.array_equal(s1, s2)


def test_flat_image_contrast(oracle):
    flat = np.full((50, 90), 0.5, np.float32)
    assert oracle.compute_contrast_factor(flat) == 0.0  # hmax = 0, threshold = 0 -> hmax*0/nbins


def test_pipeline_levels_closed_form(oracle):
    img8 = R.synthetic_image(120, 168, 9)
    r = oracle.extract(oracle.unit_float_from_u8(img8), stop_after=2)
    assert r.num_levels == 8
    k = r.contrast_factor
    prev = None
    for lv in range(r.num_levels):
        info = r.levels[lv]
        if lv == 0:
            lt = R.gaussian_blur(oracle.unit_float_from_u8(img8), 1.6)
            assert np.array_equal(lt, r.image(0, "Lt")) and np.array_equal(lt, r.image(0, "Lsmooth"))
            assert k == R.contrast_factor(lt)
        else:
            if info["octave"] > r.levels[lv - 1]["octave"]:
                p = R.half_size(prev)
                k *= 0.75
            else:
                p = prev
            ls = R.gaussian_blur(p, 1.0)
            assert np.array_equal(ls, r.image(lv, "Lsmooth"))
            c = R.pm_g2(R.scharr_x(ls, 1), R.scharr_y(ls, 1), k)
            assert np.array_equal(c, r.image(lv, "Lflow"))
            lt = p
            for tau in info["fed_tau_steps"]:
                lt, _ = R.fed_step(lt, c, tau)
            assert np.array_equal(lt, r.image(lv, "Lt"))
        s = int(round(info["esigma"] * 1.5 / 2 ** info["octave"]))
        lx, ly, lxx, lyy, lxy, ldet = R.detector(r.image(lv, "Lsmooth"), s)
        for name, a in (("Lx", lx), ("Ly", ly), ("Lxx", lxx), ("Lyy", lyy), ("Lxy", lxy), ("Ldet", ldet)):
            assert np.array_equal(a, r.image(lv, name)), (lv, name)
        prev = r.image(lv, "Lt")
