"""CPU-only: oracle vs the committed golden fixtures; the C-ABI library loads, exports every declared
symbol and refuses to run without a GPU; host-side RANSAC restatement sanity."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def summary():
    with open(os.path.join(GOLD, "summary.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ["1", "2"])
def test_oracle_matches_golden(oracle, akz, summary, name):
    gray = akz.load_gray(os.path.join(GOLD, name + ".jpg"))
    s = summary[name]
    assert hashlib.sha256(gray.tobytes()).hexdigest() == s["gray_sha256"], "JPEG decode / to_luma changed"
    r = oracle.extract(oracle.unit_float_from_u8(gray), threads=4)
    assert r.status == 0
    assert float(r.contrast_factor).hex() == s["contrast_factor_hex"]
    assert (r.num_candidates, r.num_cache, len(r.keypoints)) == (s["num_candidates"], s["num_cache"], s["num_keypoints"])
    assert [l["n_steps"] for l in r.levels] == s["n_steps"]
    for lv in range(r.num_levels):
        assert hashlib.sha256(r.image(lv, "Lt").tobytes()).hexdigest()[:16] == s["Lt_sha256_16"][lv]
        assert hashlib.sha256(r.image(lv, "Ldet").tobytes()).hexdigest()[:16] == s["Ldet_sha256_16"][lv]
    g = np.load(os.path.join(GOLD, "features_%s.npz" % name))
    assert np.array_equal(g["keypoints"], r.keypoints)
    assert np.array_equal(g["descriptors"], r.descriptors)


def test_oracle_match_golden(oracle, summary):
    a = np.load(os.path.join(GOLD, "features_1.npz"))
    b = np.load(os.path.join(GOLD, "features_2.npz"))
    m = oracle.descriptor_match(a["descriptors"], b["descriptors"], 10000, 0.86)
    g = np.load(os.path.join(GOLD, "matches_1_2.npz"))["matches"]
    assert np.array_equal(m, g) and len(m) == summary["matches_1_2"]["count"]


def test_library_exports_every_declared_symbol(akz):
    hdr = open(os.path.join(ROOT, "include", "akaze_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(akz_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    L = ctypes.CDLL(akz.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(akz.EXPORTS) == declared, "python binding list out of sync with the header"
    L.akz_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.akz_version()


def test_default_config_matches_reference(akz):
    c = akz.Config.default().to_dict()
    assert c == {"num_sublevels": 4, "max_octave_evolution": 4, "base_scale_offset": 1.6, "initial_contrast": 0.001,
                 "contrast_percentile": 0.7, "contrast_factor_num_bins": 300, "derivative_factor": 1.5,
                 "detector_threshold": 0.001, "descriptor_channels": 3, "descriptor_pattern_size": 10}


def test_no_cpu_fallback(akz):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(akz.AkazeError) as e:
        akz.Engine(0, 1024, 1024, 1)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "akaze-rust_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".rs", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "akaze_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, fn


def test_ransac_host(akz):
    from akaze_rust_b200 import ransac
    rng = np.random.default_rng(0)
    n = 60
    kp0 = np.zeros(n, akz.KEYPOINT_DTYPE)
    kp1 = np.zeros(n, akz.KEYPOINT_DTYPE)
    kp0["x"], kp0["y"] = rng.uniform(0, 1000, n), rng.uniform(0, 800, n)
    kp1["x"], kp1["y"] = kp0["x"] + 25.0, kp0["y"]  # pure horizontal shift: consistent epipolar geometry
    m = np.zeros(n, akz.MATCH_DTYPE)
    m["index_0"] = np.arange(n)
    m["index_1"] = np.arange(n)
    few = ransac.remove_outliers(kp0, kp1, m[:5], 10, 0.05, 3.0)
    assert np.array_equal(few, m[:5])  # < 8 matches are returned untouched (estimate_fundamental_matrix.rs:107-110)
    out = ransac.remove_outliers(kp0, kp1, m, 20, 0.05, 3.0)
    assert out.dtype == akz.MATCH_DTYPE and len(out) <= n
    assert set(out["index_0"]).issubset(set(m["index_0"]))


def test_no_contracted_packed_fma_in_the_library(akz):
    """The stencils use packed f32x2 multiplies (FMUL2) but must never contain a packed fused multiply-add: ptxas
    contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, which would silently break the bit-exact
    match with the reference's unfused arithmetic (profiles/r1z_pipes_microbench.txt). Scalar FFMA may only appear in
    the correctly rounded division / sqrt / double-precision helper sequences the compiler emits, never in the filters:
    the streaming kernels' steady loops are checked to be free of it."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not installed")
    sass = subprocess.run(["cuobjdump", "-sass", akz.LIB_PATH], capture_output=True, text=True).stdout
    assert "FMUL2" in sass, "the packed-product path is gone (or the library was built for another arch)"
    assert "FFMA2" not in sass, "ptxas contracted a packed multiply-add: results would no longer be bit-exact"
    # per kernel: the detector, FED and smooth+gradient stream kernels contain no scalar FFMA at all
    for fn in re.split(r"\n\s*Function : ", sass)[1:]:
        name = fn.split("\n", 1)[0]
        if any(k in name for k in ("k_detector_tmem", "k_detector_stream", "k_fed_pp")):
            assert " FFMA " not in fn, name


def test_contrast_bin_is_a_monotone_step_function():
    """k_contrast_hist_ew replaces the per-pixel f64 sqrt / divide / floor of contrast_factor.rs:45-50 by a lookup in a
    table of thresholds found by bisection (k_contrast_thresholds). That is exact iff bin(g2) is non-decreasing in g2
    and the bisection evaluates the same formula; restated in numpy float64 (IEEE, like the device code) and checked
    on random gradients, including values next to every threshold."""
    def bin_of(g2, hmax, n_bins):
        modg = np.sqrt(g2)
        bf = np.floor(n_bins * (modg / hmax))
        return np.minimum(np.where(bf > 0.0, np.minimum(bf, n_bins), 0.0), n_bins - 1).astype(np.int64)

    rng = np.random.default_rng(7)
    n_bins = 300
    lx = rng.normal(0, 0.02, 200000).astype(np.float32).astype(np.float64)
    ly = rng.normal(0, 0.02, 200000).astype(np.float32).astype(np.float64)
    g2 = lx * lx + ly * ly
    g2 = g2[g2 > 0]
    hmax = float(np.sqrt(g2.max()))
    # thresholds by bisection over the f64 bit patterns (positive doubles order like u64)
    thr = np.empty(n_bins + 1)
    thr[0], thr[n_bins] = 0.0, np.inf
    for b in range(1, n_bins):
        lo, hi = np.uint64(1), np.uint64(0x7FEFFFFFFFFFFFFF)
        while hi - lo > 1:
            mid = lo + ((hi - lo) >> np.uint64(1))
            if bin_of(np.array([mid], np.uint64).view(np.float64), hmax, n_bins)[0] >= b:
                hi = mid
            else:
                lo = mid
        thr[b] = np.array([hi], np.uint64).view(np.float64)[0]
    assert np.all(np.diff(thr[:-1]) >= 0)
    by_table = np.searchsorted(thr, g2, side="right") - 1
    assert np.array_equal(by_table, bin_of(g2, hmax, n_bins))
    # the doubles on either side of every threshold
    edge = np.concatenate([np.nextafter(thr[1:-1], 0.0), thr[1:-1], np.nextafter(thr[1:-1], np.inf)])
    edge = edge[edge > 0]
    assert np.array_equal(np.searchsorted(thr, edge, side="right") - 1, bin_of(edge, hmax, n_bins))
