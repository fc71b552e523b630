"""ctypes loader for the CPU oracle (oracle/akaze_oracle.c). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The product package (akaze-rust_b200) never does. PARITY UNPINNED -- see akaze_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libakaze_oracle.so")

IMAGE_KINDS = {"Lt": 0, "Lsmooth": 1, "Lx": 2, "Ly": 3, "Lxx": 4, "Lyy": 5, "Lxy": 6, "Lflow": 7,
               "Lstep": 8, "Ldet": 9}


class Config(C.Structure):
    """types::evolution::Config, field for field (akaze/src/types/evolution.rs:8-38)."""
    _fields_ = [
        ("num_sublevels", C.c_uint32),
        ("max_octave_evolution", C.c_uint32),
        ("base_scale_offset", C.c_double),
        ("initial_contrast", C.c_double),
        ("contrast_percentile", C.c_double),
        ("contrast_factor_num_bins", C.c_uint64),
        ("derivative_factor", C.c_double),
        ("detector_threshold", C.c_double),
        ("descriptor_channels", C.c_uint64),
        ("descriptor_pattern_size", C.c_uint64),
    ]


class LevelInfo(C.Structure):
    _fields_ = [("octave", C.c_uint32), ("sublevel", C.c_uint32), ("sigma_size", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("n_steps", C.c_uint32),
                ("esigma", C.c_double), ("etime", C.c_double)]


KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("size", "<f4"),
                           ("octave", "<u4"), ("class_id", "<u4"), ("angle", "<f4")])
MATCH_DTYPE = np.dtype([("index_0", "<u8"), ("index_1", "<u8"), ("distance", "<f8")])


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("akaze_oracle.c", "akaze_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    L.akzo_default_config.argtypes = [C.POINTER(Config)]
    L.akzo_gaussian_kernel.argtypes = [C.c_float, C.c_int, f32p]
    L.akzo_scharr_main_axis_kernel.argtypes = [C.c_uint32, f32p]
    L.akzo_scharr_off_axis_kernel.argtypes = [C.c_uint32, f32p]
    L.akzo_fed_tau_by_process_time.argtypes = [C.c_double, C.c_int, C.c_double, C.c_int, f64p, C.c_int]
    L.akzo_fed_tau_by_process_time.restype = C.c_int
    L.akzo_unit_float_from_u8.argtypes = [u8p, C.c_size_t, f32p]
    L.akzo_horizontal_filter.argtypes = [f32p, C.c_int, C.c_int, f32p, C.c_int, f32p]
    L.akzo_vertical_filter.argtypes = [f32p, C.c_int, C.c_int, f32p, C.c_int, f32p]
    L.akzo_gaussian_blur.argtypes = [f32p, C.c_int, C.c_int, C.c_float, f32p]
    L.akzo_scharr.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, f32p]
    L.akzo_half_size.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.akzo_pm_g2.argtypes = [f32p, f32p, C.c_size_t, C.c_double, f32p]
    L.akzo_compute_contrast_factor.argtypes = [f32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint64]
    L.akzo_compute_contrast_factor.restype = C.c_double
    L.akzo_calculate_step.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_double]
    L.akzo_extract.argtypes = [f32p, C.c_uint32, C.c_uint32, C.POINTER(Config), C.c_int, C.c_int]
    L.akzo_extract.restype = C.c_void_p
    L.akzo_result_free.argtypes = [C.c_void_p]
    L.akzo_result_status.argtypes = [C.c_void_p]
    L.akzo_result_num_levels.argtypes = [C.c_void_p]
    L.akzo_result_num_levels.restype = C.c_uint32
    L.akzo_result_level_info.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(LevelInfo)]
    L.akzo_result_fed_tau.argtypes = [C.c_void_p, C.c_uint32]
    L.akzo_result_fed_tau.restype = C.POINTER(C.c_double)
    L.akzo_result_image.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
    L.akzo_result_image.restype = C.POINTER(C.c_float)
    L.akzo_result_contrast_factor.argtypes = [C.c_void_p]
    L.akzo_result_contrast_factor.restype = C.c_double
    for name in ("num_candidates", "num_cache", "num_keypoints", "descriptor_len"):
        fn = getattr(L, "akzo_result_" + name)
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_uint64
    L.akzo_result_keypoints.argtypes = [C.c_void_p]
    L.akzo_result_keypoints.restype = C.c_void_p
    L.akzo_result_descriptors.argtypes = [C.c_void_p]
    L.akzo_result_descriptors.restype = C.c_void_p
    L.akzo_match_top2.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_uint64, C.c_uint64, u32p, u32p, u32p]
    L.akzo_descriptor_match.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.c_uint64, C.c_double, C.c_void_p]
    L.akzo_descriptor_match.restype = C.c_uint64
    _lib = L
    return L


def default_config():
    c = Config()
    lib().akzo_default_config(C.byref(c))
    return c


def gaussian_kernel(r, size):
    out = np.zeros(size, np.float32)
    lib().akzo_gaussian_kernel(r, size, out)
    return out


def scharr_main_axis_kernel(scale):
    out = np.zeros(2 * scale + 1, np.float32)
    lib().akzo_scharr_main_axis_kernel(scale, out)
    return out


def scharr_off_axis_kernel(scale):
    out = np.zeros(2 * scale + 1, np.float32)
    lib().akzo_scharr_off_axis_kernel(scale, out)
    return out


def fed_tau_by_process_time(T, M=1, tau_max=0.25, reordering=True):
    out = np.zeros(4096, np.float64)
    n = lib().akzo_fed_tau_by_process_time(T, M, tau_max, int(reordering), out, 4096)
    if n < 0:
        raise RuntimeError("fed_tau: reference would not terminate (n==1) or too many steps")
    return out[:n].copy()


def unit_float_from_u8(gray):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty(gray.shape, np.float32)
    lib().akzo_unit_float_from_u8(gray.reshape(-1), gray.size, out.reshape(-1))
    return out


def _img(a):
    a = np.ascontiguousarray(a, np.float32)
    assert a.ndim == 2
    return a


def horizontal_filter(img, kernel):
    img = _img(img)
    k = np.ascontiguousarray(kernel, np.float32)
    out = np.empty_like(img)
    lib().akzo_horizontal_filter(img, img.shape[1], img.shape[0], k, k.size, out)
    return out


def vertical_filter(img, kernel):
    img = _img(img)
    k = np.ascontiguousarray(kernel, np.float32)
    out = np.empty_like(img)
    lib().akzo_vertical_filter(img, img.shape[1], img.shape[0], k, k.size, out)
    return out


def gaussian_blur(img, r):
    img = _img(img)
    out = np.empty_like(img)
    lib().akzo_gaussian_blur(img, img.shape[1], img.shape[0], r, out)
    return out


def scharr(img, x_order, y_order, sigma_size):
    img = _img(img)
    out = np.empty_like(img)
    lib().akzo_scharr(img, img.shape[1], img.shape[0], int(x_order), int(y_order), sigma_size, out)
    return out


def half_size(img):
    img = _img(img)
    out = np.empty((img.shape[0] // 2, img.shape[1] // 2), np.float32)
    lib().akzo_half_size(img, img.shape[1], img.shape[0], out)
    return out


def pm_g2(lx, ly, k):
    lx, ly = _img(lx), _img(ly)
    out = np.empty_like(lx)
    lib().akzo_pm_g2(lx, ly, lx.size, k, out)
    return out


def compute_contrast_factor(img, percentile=0.7, scale=1.0, nbins=300):
    img = _img(img)
    return lib().akzo_compute_contrast_factor(img, img.shape[1], img.shape[0], percentile, scale, nbins)


def calculate_step(lt, lflow, step_size):
    """Returns (new Lt, Lstep); inputs untouched."""
    lt = _img(lt).copy()
    lflow = _img(lflow)
    lstep = np.zeros_like(lt)
    lib().akzo_calculate_step(lt, lflow, lstep, lt.shape[1], lt.shape[0], step_size)
    return lt, lstep


class Result:
    """Owns an akzo_result; mirrors the tuple returned by akaze::extract_features (lib.rs:167-194)."""

    def __init__(self, handle):
        self._h = handle
        L = lib()
        self.status = L.akzo_result_status(handle)
        self.num_levels = L.akzo_result_num_levels(handle)
        self.levels = []
        for i in range(self.num_levels):
            info = LevelInfo()
            L.akzo_result_level_info(handle, i, C.byref(info))
            taus = np.ctypeslib.as_array(L.akzo_result_fed_tau(handle, i), (info.n_steps,)).copy() \
                if info.n_steps else np.zeros(0)
            self.levels.append(dict(octave=info.octave, sublevel=info.sublevel, sigma_size=info.sigma_size,
                                    width=info.width, height=info.height, n_steps=info.n_steps,
                                    esigma=info.esigma, etime=info.etime, fed_tau_steps=taus))
        self.contrast_factor = L.akzo_result_contrast_factor(handle)
        self.num_candidates = L.akzo_result_num_candidates(handle)
        self.num_cache = L.akzo_result_num_cache(handle)
        n = L.akzo_result_num_keypoints(handle)
        self.descriptor_len = L.akzo_result_descriptor_len(handle)
        if n:
            kp = (C.c_uint8 * (n * KEYPOINT_DTYPE.itemsize)).from_address(L.akzo_result_keypoints(handle))
            self.keypoints = np.frombuffer(kp, KEYPOINT_DTYPE).copy()
            d = (C.c_uint8 * (n * self.descriptor_len)).from_address(L.akzo_result_descriptors(handle))
            self.descriptors = np.frombuffer(d, np.uint8).reshape(n, self.descriptor_len).copy()
        else:
            self.keypoints = np.zeros(0, KEYPOINT_DTYPE)
            self.descriptors = np.zeros((0, max(self.descriptor_len, 1)), np.uint8)

    def image(self, level, kind):
        p = lib().akzo_result_image(self._h, level, IMAGE_KINDS[kind] if isinstance(kind, str) else kind)
        if not p:
            return None
        lv = self.levels[level]
        return np.ctypeslib.as_array(p, (lv["height"], lv["width"])).copy()

    def close(self):
        if self._h:
            lib().akzo_result_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def extract(unit_gray, cfg=None, threads=1, stop_after=0):
    """lib.rs:167-194 from the GrayFloatImage on (decode and to_luma stay with the caller)."""
    img = _img(unit_gray)
    cfg = cfg or default_config()
    h = lib().akzo_extract(img, img.shape[1], img.shape[0], C.byref(cfg), threads, stop_after)
    if not h:
        raise MemoryError("akzo_extract")
    return Result(h)


def match_top2(q, db, desc_len=None):
    """Raw top-2 per query (feature_matching.rs:37-50): (best_idx, best, second) as uint32."""
    q = np.ascontiguousarray(q, np.uint8)
    db = np.ascontiguousarray(db, np.uint8)
    stride = q.shape[1] if q.ndim == 2 and q.shape[0] else (db.shape[1] if db.ndim == 2 else 64)
    desc_len = desc_len or stride
    nq, ndb = q.shape[0], db.shape[0]
    bi = np.zeros(nq, np.uint32)
    b = np.zeros(nq, np.uint32)
    s = np.zeros(nq, np.uint32)
    lib().akzo_match_top2(q.reshape(-1), nq, db.reshape(-1), ndb, desc_len, stride, bi, b, s)
    return bi, b, s


def descriptor_match(d0, d1, distance_threshold=10000, lowes_ratio=0.86, desc_len=None):
    d0 = np.ascontiguousarray(d0, np.uint8)
    d1 = np.ascontiguousarray(d1, np.uint8)
    stride = d0.shape[1]
    desc_len = desc_len or stride
    out = np.zeros(d0.shape[0], MATCH_DTYPE)
    n = lib().akzo_descriptor_match(d0.reshape(-1), d0.shape[0], d1.reshape(-1), d1.shape[0], desc_len, stride,
                                    distance_threshold, lowes_ratio, out.ctypes.data)
    return out[:n].copy()
