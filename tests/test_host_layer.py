"""CPU-only checks of the host layer (no compute calls): on-disk formats for every descriptor width, gray-image loading,
argument validation that must not reach the library, the C ABI's symbol list."""
import os
import re
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _features(akz, n, width, seed=0):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, akz.KEYPOINT_DTYPE)
    k["x"], k["y"] = rng.uniform(0, 640, n), rng.uniform(0, 480, n)
    k["response"], k["size"], k["angle"] = rng.uniform(0, 0.1, n), 4.8, rng.uniform(0, 3.1, n)
    k["octave"], k["class_id"] = rng.integers(0, 4, n), rng.integers(0, 16, n)
    return k, rng.integers(0, 256, (n, width), dtype=np.uint8)


@pytest.mark.parametrize("channels", [1, 2, 3])
def test_feature_files_for_every_descriptor_width(akz, tmp_path, channels):
    """Descriptor.vector has (162*channels+7)/8 = 21 / 41 / 61 bytes (descriptors.rs:42-46) and akaze-util writes each
    vector with its own length (akaze-util/src/lib.rs:11-30): bincode and JSON round trips for all three widths."""
    from akaze_rust_b200 import formats
    width = (162 * channels + 7) // 8
    k, d = _features(akz, 9, width, seed=channels)
    b = formats.features_to_bytes(k, d)
    assert len(b) == 8 + 9 * 36 + 8 + 9 * (8 + width)
    k2, d2 = formats.features_from_bytes(b)
    assert np.array_equal(np.stack(d2), d) and np.array_equal(k2["x"], k["x"]) and np.array_equal(k2["class_id"], k["class_id"])
    for ext in ("bin", "json"):
        p = tmp_path / ("f." + ext)
        formats.serialize_features_to_file(k, d, str(p))
        k3, d3 = formats.deserialize_features_from_file(str(p))
        assert all(len(v) == width for v in d3) and np.array_equal(np.stack(d3), d)
        assert np.array_equal(k3["angle"], k["angle"]) and np.array_equal(k3["octave"], k["octave"])
    # the engine's padded 64-byte rows cut to the descriptor length
    pad = np.zeros((9, 64), np.uint8)
    pad[:, :width] = d
    assert formats.features_to_bytes(k, pad, width) == b
    # ragged vectors (a list) keep their own lengths
    rag = [d[i][: 1 + i] for i in range(9)]
    k4, d4 = formats.features_from_bytes(formats.features_to_bytes(k, rag))
    assert [len(v) for v in d4] == [1 + i for i in range(9)]
    with pytest.raises(ValueError):
        formats.features_to_bytes(k, d[:5])


def test_load_gray_keeps_gray_sources_unchanged(akz, tmp_path):
    """image::open + to_luma (lib.rs:171, image.rs:128): Luma8 sources pass through unchanged (to_luma is the identity
    there); only RGB sources go through the Rec.709 weights."""
    from PIL import Image
    ramp = np.tile(np.arange(256, dtype=np.uint8), (4, 1))
    p = tmp_path / "ramp.png"
    Image.fromarray(ramp).save(p)
    assert np.array_equal(akz.load_gray(str(p)), ramp)  # all 256 gray levels survive
    rgb = np.stack([ramp, ramp, ramp], axis=-1)
    q = tmp_path / "ramp_rgb.png"
    Image.fromarray(rgb).save(q)
    g = akz.load_gray(str(q))
    assert g.shape == ramp.shape and np.abs(g.astype(int) - ramp.astype(int)).max() <= 1  # truncating f32 weights
    assert np.array_equal(akz.to_luma_u8(ramp), ramp)


def test_header_and_bindings_agree(akz):
    """Every function include/akaze_b200.h declares is bound by the host layer and exported by the library."""
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "akaze_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(akz_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(akz.EXPORTS), declared ^ set(akz.EXPORTS)
    L = ctypes.CDLL(akz.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    for name in ("akz_comm_unique_id", "akz_context_comm_init", "akz_context_comm_init_all", "akz_match_top2_sharded_device", "akz_match_top2_sharded"):
        assert name in declared


def test_argument_validation_needs_no_gpu(akz):
    e = akz.Engine.__new__(akz.Engine)  # no context: the checks below must fire before the library is called
    e._h = None
    with pytest.raises(ValueError):
        e.extract_u8(np.zeros((32, 64, 3), np.uint8))
    with pytest.raises(ValueError):
        e.extract_batch_u8([np.zeros((32, 64), np.uint8), np.zeros((32, 65), np.uint8)])
    with pytest.raises(ValueError):
        e.descriptor_match(np.zeros((3, 64), np.uint8), np.zeros((3, 61), np.uint8))
    with pytest.raises(ValueError):
        e.match_top2(np.zeros((3, 65), np.uint8), np.zeros((3, 65), np.uint8))
    assert e.extract_batch_u8([]) == []


def test_rust_overlay_lists_every_source_and_symbol(akz):
    """The Rust overlay cannot be compiled here; at least keep it honest: build.rs names every CUDA source build.py
    compiles, every extern "C" function ffi.rs declares exists in the header, lib.rs only calls functions ffi.rs declares."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("akz_build", os.path.join(ROOT, "akaze-rust_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    rust = os.path.join(ROOT, "akaze-rust_b200", "rust")
    build_rs = open(os.path.join(rust, "build.rs")).read()
    assert sorted(re.findall(r'"(\w+\.cu)"', build_rs)) == sorted(b.SOURCES)
    ffi = open(os.path.join(rust, "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (akz_\w+)", ffi))
    assert declared and declared <= set(akz.EXPORTS), declared - set(akz.EXPORTS)
    lib_rs = open(os.path.join(rust, "src", "lib.rs")).read()
    assert set(re.findall(r"ffi::(akz_[a-z0-9_]+)\(", lib_rs)) <= declared
    for const in set(re.findall(r"ffi::(AKZ_\w+)", lib_rs)):
        assert re.search(r"pub const %s\b" % const, ffi), const
    assert not os.path.exists(os.path.join(rust, "src", "types")) and "download_all(f" in lib_rs and "unsafe fn download_all" in lib_rs


def test_visualisation_helpers_follow_the_reference(akz, tmp_path):
    """types/image.rs:148-210,385-480, evolution.rs:163-218, keypoint.rs:52-72, feature_match.rs:32-82."""
    from akaze_rust_b200 import visualize as V
    a = np.array([[0.25, 0.5], [0.75, 1.25]], np.float32)
    n = V.normalize(a)
    assert n.dtype == np.float32 and n.min() == 0.0 and n.max() == 1.0 and n[0, 1] == np.float32(0.25) / np.float32(1.0)
    assert np.array_equal(V.create_dynamic_image(np.array([[0.0, 0.5, 0.999, 1.0]], np.float32)), [[0, 127, 254, 255]])  # truncating cast
    assert V.build_path(tmp_path, "Lt_", 3).endswith("Lt_00003..png")  # the reference's set_extension(".png") quirk
    assert V.save(np.zeros((0, 0), np.float32), str(tmp_path / "empty.png")) is False  # 0x0 images are skipped
    assert V.save(a, str(tmp_path / "a.png")) and os.path.getsize(tmp_path / "a.png") > 0
    assert V.random_color() == V.random_color()  # a fresh default source per call: one colour for everything
    img = np.zeros((40, 60, 3), np.uint8)
    V.draw_circle(img, (20.0, 20.0), (200, 100, 50), 5.0)
    assert tuple(img[20, 20]) == (100, 50, 25)          # blend = truncated mean with the old pixel
    assert tuple(img[20, 24]) == (100, 50, 25) and tuple(img[20, 25]) == (0, 0, 0)  # half-open pixel range: x in [15, 25)
    assert tuple(img[15, 20]) == (100, 50, 25) and tuple(img[14, 20]) == (0, 0, 0)
    with pytest.raises(IndexError):                      # the reference panics when a circle leaves the image
        V.draw_circle(img, (58.0, 20.0), (1, 2, 3), 5.0)
    img2 = np.zeros((40, 60, 3), np.uint8)
    V.draw_line(img2, (10.0, 10.0), (40.0, 25.0), (255, 255, 255), 1.0)
    assert img2[10, 10].any() and img2[17, 25].any() and not img2[30, 10].any()
    k = np.zeros(2, akz.KEYPOINT_DTYPE)
    k["x"], k["y"], k["size"] = (12.0, 30.0), (12.0, 20.0), (3.0, 4.0)
    g0 = np.full((32, 48), 10, np.uint8)
    over = V.draw_keypoints(g0, k)
    assert over.shape == (32, 48, 3) and over[12, 12].tolist() != [10, 10, 10] and over[0, 0].tolist() == [10, 10, 10]
    m = np.zeros(1, akz.MATCH_DTYPE)
    m["index_0"], m["index_1"] = 0, 1
    both = V.draw_matches(g0, np.full((40, 48), 20, np.uint8), k, k, m)
    assert both.shape == (40, 96, 3) and both[0, 0, 0] == 10 and both[0, 48, 0] == 20 and both[35, 0, 0] == 0
    ev = [types.SimpleNamespace(Lt=a, Lsmooth=a, Lx=a, Ly=a, Lxx=a, Lyy=a, Lxy=a, Lflow=np.zeros((0, 0), np.float32),
                                Lstep=np.zeros((0, 0), np.float32), Ldet=a)]
    written = V.write_evolutions(ev, tmp_path / "evo")
    assert len(written) == 8 and os.path.exists(tmp_path / "evo" / "Ldet_00000..png")
