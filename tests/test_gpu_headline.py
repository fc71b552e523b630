"""Parity gates on the HEADLINE path: bench.py's own 1920x1080 images, batch call, default mode.

The session `engine` fixture keeps all ten evolution images (AKZ_KEEP_EVOLUTIONS), which routes the detector through the
generic tile kernel; the kernels the bench times (k_detector_tmem<2,3,4>, k_fed_pp, k_prep_stream, the fused contrast
pass) run only in default mode. Here they are compared image for image, level for level, bit for bit with the CPU
oracle, through akz_features_evolution_download, which serves the four persistent planes (Lt, Lx, Ly, Ldet) without the
keep flag. Tolerances as in test_gpu_parity.py (north star: evolutions max-abs <= 1e-5 -- asserted as bit-exact).
"""
import numpy as np
import pytest

import np_restatement as R

pytestmark = pytest.mark.gpu

ANGLE_TOL = 5e-7
DESC_BIT_FRACTION = 1e-4
PLANES = ("Lt", "Lx", "Ly", "Ldet")


def _check_features(f, ref, planes=True):
    assert ref.status == 0
    assert f.contrast_factor == ref.contrast_factor
    assert f.num_cache == ref.num_cache and len(f.keypoints) == len(ref.keypoints)
    for k in ("x", "y", "response", "size", "octave", "class_id"):
        assert np.array_equal(f.keypoints[k], ref.keypoints[k]), k
    assert np.abs(f.keypoints["angle"] - ref.keypoints["angle"]).max() <= ANGLE_TOL
    bits = int(np.unpackbits(f.descriptors ^ ref.descriptors).sum())
    assert bits <= max(2, DESC_BIT_FRACTION * ref.descriptors.size * 8), bits
    if planes:
        for lv in range(ref.num_levels):
            for kind in PLANES:
                a, b = f.evolution(lv, kind), ref.image(lv, kind)
                assert np.array_equal(a, b), "level %d %s: max abs diff %g" % (lv, kind, np.abs(a - b).max())
        assert np.array_equal(f.evolution(0, "Lsmooth"), ref.image(0, "Lsmooth"))  # Lsmooth_0 is Lt_0 (lib.rs:58)


@pytest.fixture(scope="module")
def bench_images():
    return [R.natural_image(1080, 1920, 1000 + i) for i in range(4)]  # bench.py's images 0..3 (seed0 = 1000)


def test_default_mode_1080p_batch_is_bit_exact(akz, oracle, bench_images):
    """configs[2] shape, default mode, one batch of 4: every persistent plane of all 16 levels == oracle, plus keypoints
    and descriptors. This is the direct check of k_detector_tmem<S>, k_fed_pp<T>, k_prep_stream and the fused contrast pass."""
    eng = akz.Engine(0, 1920, 1080, 4)
    fs = eng.extract_batch_u8(bench_images)
    for img, f in zip(bench_images, fs):
        ref = oracle.extract(oracle.unit_float_from_u8(img), threads=8)
        assert ref.num_levels == 16 and len(ref.keypoints) > 3000
        _check_features(f, ref)
        with pytest.raises(akz.AkazeError):
            f.evolution(3, "Lflow")  # scratch plane: needs keep_evolutions
        ref.close()
    eng.close()
    with pytest.raises(akz.AkazeError):  # the engine is gone: no use-after-free, a clean error
        fs[0].evolution(1, "Lt")


def test_default_mode_sub_batches_and_residency(akz, oracle, bench_images):
    """Three sub-batches of one image: the two work-buffer lanes alternate, so the last two images' planes are still
    resident (and bit-exact), the first one's were overwritten and its download is refused."""
    eng = akz.Engine(0, 1920, 1080, 3)
    eng.set_sub_batch(1)
    fs = eng.extract_batch_u8(bench_images[:3])
    with pytest.raises(akz.AkazeError):
        fs[0].evolution(2, "Ldet")
    for i in (1, 2):
        ref = oracle.extract(oracle.unit_float_from_u8(bench_images[i]), threads=8)
        _check_features(fs[i], ref)
        ref.close()
    ref0 = oracle.extract(oracle.unit_float_from_u8(bench_images[0]), threads=8)
    _check_features(fs[0], ref0, planes=False)
    # a later extraction invalidates the earlier handles' evolutions
    g = eng.extract_u8(bench_images[3])
    with pytest.raises(akz.AkazeError):
        fs[2].evolution(0, "Lt")
    ref3 = oracle.extract(oracle.unit_float_from_u8(bench_images[3]), threads=8)
    _check_features(g, ref3)
    eng.close()


def test_raw_4k_frame_fits_default_capacities(akz, oracle):
    """configs[3]: an unsoftened 3840x2160 frame of the bench generator works with the default limits (they scale with
    the image: the reference's Vec<Keypoint> has no capacity) and matches the oracle keypoint for keypoint."""
    img = R.natural_image(2160, 3840, 5000)
    eng = akz.Engine(0, 3840, 2160, 1)
    f = eng.extract_u8(img)
    ref = oracle.extract(oracle.unit_float_from_u8(img), threads=8)
    assert len(ref.keypoints) > 15000
    _check_features(f, ref)
    eng.close()


def test_thirty_two_levels(akz, oracle):
    """The most levels the engine admits (8 sublevels x 4 octaves = 32): the level-pipelined cache pass then runs one warp per
    level in a 1024-thread block, in default mode and with every evolution kept."""
    img = R.synthetic_image(480, 640, seed=23)
    cg, co = akz.Config.default(), oracle.default_config()
    cg.num_sublevels = co.num_sublevels = 8
    ref = oracle.extract(oracle.unit_float_from_u8(img), co, threads=8)
    assert ref.status == 0 and ref.num_levels == 32 and len(ref.keypoints) > 200
    for keep in (False, True):
        eng = akz.Engine(0, 640, 480, 1, keep_evolutions=keep)
        f = eng.extract_u8(img, cg)
        assert len(f.evolutions) == 32
        _check_features(f, ref)
        eng.close()


def _literal_descriptor_match(d0, d1, T, lowes):
    """feature_matching.rs:23-94 with full distances (the bail-out never changes the outcome, see
    test_oracle_vs_numpy_keypoints.py)."""
    d = np.unpackbits(d0[:, None, :] ^ d1[None, :, :], axis=2).sum(axis=2) if len(d1) else np.zeros((len(d0), 0), np.int64)
    out = []
    for i in range(len(d0)):
        mn, mj, sec = T, 0, T
        for j in range(d.shape[1]):
            v = int(d[i, j])
            if v < mn:
                sec, mn, mj = mn, v, j
            elif v < sec:
                sec = v
        if float(mn) < float(sec) * lowes ** 2 and mn < T:
            out.append((i, mj, float(mn)))
    return out


def test_descriptor_match_any_threshold(akz, oracle):
    """descriptor_match is public and takes any distance_threshold (feature_matching.rs:23-28); the seeds of min and
    second-to-min ARE the threshold (:38-40), so it changes the Lowe test too."""
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (80, 61), dtype=np.uint8)
    b = rng.integers(0, 256, (300, 61), dtype=np.uint8)
    for i in range(0, 80, 3):  # near duplicates at graded distances
        b[(7 * i) % 300] = a[i]
        flips = rng.integers(0, 61 * 8, size=i)
        for bit in flips:
            b[(7 * i) % 300, bit // 8] ^= 1 << (bit % 8)
    eng = akz.Engine(0, 64, 64, 1)
    for T in (10000, 300, 120, 60, 20, 1, 0, 1 << 40):
        for lowes in (0.86, 1.0, 2.0):
            m = eng.descriptor_match(a, b, T, lowes)
            lit = _literal_descriptor_match(a, b, T, lowes)
            assert [(int(r["index_0"]), int(r["index_1"]), float(r["distance"])) for r in m] == lit, (T, lowes)
            if T <= 10000:
                mo = oracle.descriptor_match(a, b, T, lowes)
                assert np.array_equal(m, mo), (T, lowes)
    for nb in (0, 1):  # fewer than two database descriptors: the seeds show through
        for T in (10000, 50000, 200):
            m = eng.descriptor_match(a, b[:nb], T, 0.86)
            assert [(int(r["index_0"]), int(r["index_1"]), float(r["distance"])) for r in m] == _literal_descriptor_match(a, b[:nb], T, 0.86)
    with pytest.raises(ValueError):
        eng.descriptor_match(np.zeros((4, 64), np.uint8), np.zeros((4, 61), np.uint8))
    with pytest.raises(ValueError):
        eng.extract_u8(np.zeros((64, 64, 3), np.uint8))
    eng.close()


def test_sharded_match_inside_the_library(akz, oracle):
    """akz_match_top2_sharded* (NCCL all-gather + merge inside the library) == the unsharded scan, bit for bit, ties
    included. With one visible GPU the communicator has one rank; with several, one engine per device in one process."""
    import torch
    n_dev = torch.cuda.device_count()
    rng = np.random.default_rng(17)
    q = rng.integers(0, 256, (3000, 64), dtype=np.uint8)
    db = rng.integers(0, 256, (5003, 64), dtype=np.uint8)
    q[:, 61:] = 0
    db[:, 61:] = 0
    db[rng.integers(0, len(db), 400)] = q[rng.integers(0, len(q), 400)]  # duplicates across shards: ties on the minimum
    db[4000:4100] = db[100:200]
    bi, b, s = oracle.match_top2(q, db, desc_len=61)
    for n in sorted({1, min(2, n_dev), n_dev}):
        engines = [akz.Engine(d, 64, 64, 1) for d in range(n)]
        akz.comm_init_all(engines)
        for path in ("popc", "tensor"):
            for e in engines:
                e.set_match_path(path)
            t = akz.match_top2_sharded(engines, q, db, desc_len=61)
            assert np.array_equal(t["best_idx"], bi) and np.array_equal(t["best"], b) and np.array_equal(t["second"], s), (n, path)
        t = akz.match_top2_sharded(engines, q[:5], db[:1], desc_len=61)  # more ranks than descriptors: empty shards
        bi1, b1, s1 = oracle.match_top2(q[:5], db[:1], desc_len=61)
        assert np.array_equal(t["best_idx"], bi1) and np.array_equal(t["best"], b1) and np.array_equal(t["second"], s1)
        for e in engines:
            e.comm_destroy()
            e.close()
    # per-rank form with a one-rank communicator built from a unique id
    eng = akz.Engine(0, 64, 64, 1)
    eng.comm_init(akz.comm_unique_id(), 0, 1)
    dq, ddb = torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda()
    out = torch.zeros(len(q), dtype=torch.int64, device="cuda")
    eng.match_top2_sharded_device(dq.data_ptr(), len(q), ddb.data_ptr(), len(db), 0, out.data_ptr())
    torch.cuda.synchronize()
    r = out.cpu().numpy().view(akz.TOP2_DTYPE)
    assert np.array_equal(r["best_idx"], bi) and np.array_equal(r["best"], b) and np.array_equal(r["second"], s)
    eng.close()


def test_cli_debug_dir_on_the_engine(akz, tmp_path):
    """extract_features -d DIR writes every evolution image it has (level 0 has no Lflow, like the reference)."""
    import os
    from akaze_rust_b200 import cli
    img = R.synthetic_image(240, 320, seed=5)
    from PIL import Image
    p = tmp_path / "in.png"
    Image.fromarray(img).save(p)  # mode L: read back unchanged
    assert np.array_equal(akz.load_gray(str(p)), img)
    rc = cli.main(["extract_features", str(p), str(tmp_path / "out.bin"), "-d", str(tmp_path / "dbg")])
    assert rc == 0
    names = sorted(os.listdir(tmp_path / "dbg"))
    # write_evolutions' file names, reference quirk included (set_extension(".png") on "Lt_00000.png")
    assert "Lt_00000..png" in names and "Ldet_00011..png" in names and "Ldet_00012..png" not in names
    assert "Lflow_00001..png" in names and "Lflow_00000..png" not in names and "Lstep_00003..png" in names
    assert "Lt_00.npy" in names and "keypoints.png" in names
    png = np.asarray(Image.open(tmp_path / "dbg" / "Lt_00000..png"))
    assert png.shape == img.shape and png.min() == 0 and png.max() == 255  # normalised to the full range


def test_device_inputs_in_rotating_buffers(akz):
    """akz_extract_batch_u8_device on a caller that rotates three device buffers (one-sub-batch calls replay cached CUDA graphs,
    keyed by the input pointer among other things): every call returns what the host-buffer call returns for the same images."""
    import ctypes
    import torch
    h, w, n = 360, 480, 2
    imgs = [np.stack([R.synthetic_image(h, w, seed=40 + 2 * b + i) for i in range(n)]) for b in range(3)]
    eng = akz.Engine(0, w, h, n)
    want = []
    for b in range(3):
        fs = eng.extract_batch_u8(imgs[b])
        want.append([(f.keypoints.copy(), f.descriptors_padded.copy()) for f in fs])
    bufs = [torch.from_numpy(imgs[b]).cuda() for b in range(3)]
    rt = ctypes.CDLL("libcudart.so.12")
    for rnd in range(3):  # every buffer is met again after the other two
        for b in (0, 1, 2, 1):
            counts = eng.extract_batch_u8_device(bufs[b].data_ptr(), n, w, h)
            d_kp, d_desc, cap = eng.device_results()
            for i in range(n):
                c = int(counts[i])
                assert c == len(want[b][i][0]) > 50
                kp = np.zeros(c, akz.KEYPOINT_DTYPE)
                de = np.zeros((c, 64), np.uint8)
                assert rt.cudaMemcpy(ctypes.c_void_p(kp.ctypes.data), ctypes.c_void_p(d_kp + i * cap * kp.itemsize), ctypes.c_size_t(kp.nbytes), 2) == 0
                assert rt.cudaMemcpy(ctypes.c_void_p(de.ctypes.data), ctypes.c_void_p(d_desc + i * cap * 64), ctypes.c_size_t(de.nbytes), 2) == 0
                assert kp.tobytes() == want[b][i][0].tobytes() and np.array_equal(de, want[b][i][1]), (rnd, b, i)
    eng.close()
