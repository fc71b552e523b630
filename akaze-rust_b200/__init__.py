"""akaze_rust_b200 -- host-side mirror of the akaze-rust public API over the B200 C ABI.

The product is libakaze_b200.so (CUDA sm_100a, include/akaze_b200.h). This module is the thin host layer
a Rust `ffi.rs` would be, written in Python because this image has no Rust toolchain.
It mirrors the reference interface for the hot path:

    akaze::extract_features(path, Config) -> (evolutions, keypoints, descriptors)   akaze/src/lib.rs:167
    akaze::match_features(kp0, d0, kp1, d1, lowes_ratio, trials, eps) -> matches     akaze/src/lib.rs:252
    akaze::types::evolution::Config (+ Default)                                      types/evolution.rs:8-55
    akaze::ops::feature_matching::descriptor_match                                    ops/feature_matching.rs:23

There is NO CPU fallback: if the shared library is missing or no B200 is present every call raises.
Nothing in this package imports oracle/ (the CPU restatement is test infrastructure only).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AKZ_FAST_MATH=1 (read once, at import) selects the opt-in build whose stencil kernels use fused multiply-adds: faster,
# no longer bit-identical to the reference (inside the north-star tolerances; tools/fast_math_report.py). Default: exact.
FAST_MATH = os.environ.get("AKZ_FAST_MATH", "0") not in ("", "0")
LIB_PATH = os.path.join(_HERE, "libakaze_b200_fast.so" if FAST_MATH else "libakaze_b200.so")

DESCRIPTOR_STRIDE = 64
AKZ_KEEP_EVOLUTIONS = 1

IMAGE_KINDS = {"Lt": 0, "Lsmooth": 1, "Lx": 2, "Ly": 3, "Lxx": 4, "Lyy": 5, "Lxy": 6, "Lflow": 7,
               "Lstep": 8, "Ldet": 9}

STAGES = ("level0", "contrast", "prep", "fed", "detector", "compact", "dedup", "finalize", "descriptor")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("size", "<f4"),
                           ("octave", "<u4"), ("class_id", "<u4"), ("angle", "<f4")])
TOP2_DTYPE = np.dtype([("best_idx", "<u4"), ("best", "<u2"), ("second", "<u2")])
MATCH_DTYPE = np.dtype([("index_0", "<u8"), ("index_1", "<u8"), ("distance", "<f8")])


class AkazeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("akaze_b200 error %d: %s" % (code, msg))
        self.code = code


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

    @classmethod
    def default(cls):
        """Config::default() (evolution.rs:41-54)."""
        c = cls()
        _check(lib().akz_default_config(C.byref(c)))
        return c

    def to_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class LevelInfo(C.Structure):
    _fields_ = [("octave", C.c_uint32), ("sublevel", C.c_uint32), ("sigma_size", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("n_steps", C.c_uint32),
                ("esigma", C.c_double), ("etime", C.c_double)]


_lib = None

# every symbol include/akaze_b200.h declares
EXPORTS = [
    "akz_last_error", "akz_version", "akz_default_config", "akz_create", "akz_destroy", "akz_context_stream",
    "akz_context_launch_count", "akz_context_set_limits", "akz_context_set_sub_batch", "akz_context_set_match_path", "akz_context_enable_timing", "akz_context_stage_times", "akz_extract_u8", "akz_extract_f32",
    "akz_extract_batch_u8", "akz_extract_batch_u8_device", "akz_context_device_results", "akz_features_count",
    "akz_features_keypoints", "akz_features_descriptors", "akz_features_descriptor_len",
    "akz_features_num_levels", "akz_features_level_info", "akz_features_fed_tau",
    "akz_features_contrast_factor", "akz_features_num_candidates", "akz_features_num_cache",
    "akz_features_evolution_download", "akz_features_free", "akz_match_top2", "akz_match_top2_device",
    "akz_merge_top2_device", "akz_descriptor_match",
    "akz_comm_unique_id", "akz_context_comm_init", "akz_context_comm_init_all", "akz_context_comm_destroy",
    "akz_match_top2_sharded_device", "akz_match_top2_sharded", "akz_remove_outliers",
]


def _nccl_hint():
    """The multi-GPU matcher binds NCCL at run time (dlopen "libnccl.so.2"); point it at the wheel's copy when the
    process has not loaded one already (torch loads its own) and the caller did not choose."""
    if os.environ.get("AKZ_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["AKZ_NCCL_LIB"] = cand
                return
    except Exception:  # noqa: BLE001 -- a hint only; akz_comm_* reports a missing NCCL itself
        pass


def lib():
    """Loads libakaze_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    _nccl_hint()
    if not os.path.exists(LIB_PATH):
        raise ImportError("libakaze_b200.so is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first; "
                          "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.akz_last_error.restype = C.c_char_p
    L.akz_version.restype = C.c_char_p
    L.akz_default_config.argtypes = [C.POINTER(Config)]
    L.akz_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.akz_destroy.argtypes = [vp]
    L.akz_destroy.restype = None
    L.akz_context_stream.argtypes = [vp]
    L.akz_context_stream.restype = vp
    L.akz_context_launch_count.argtypes = [vp]
    L.akz_context_launch_count.restype = C.c_uint64
    L.akz_context_set_limits.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.akz_context_set_sub_batch.argtypes = [vp, C.c_uint32]
    L.akz_context_set_match_path.argtypes = [vp, C.c_int]
    L.akz_context_enable_timing.argtypes = [vp, C.c_int]
    L.akz_context_stage_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.akz_extract_u8.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_size_t, C.POINTER(Config), C.POINTER(vp)]
    L.akz_extract_f32.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.POINTER(Config), C.POINTER(vp)]
    L.akz_extract_batch_u8.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_size_t,
                                       C.POINTER(Config), C.POINTER(vp)]
    L.akz_extract_batch_u8_device.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_size_t,
                                              C.POINTER(Config), C.POINTER(C.c_uint32)]
    L.akz_context_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32)]
    L.akz_features_count.argtypes = [vp]
    L.akz_features_count.restype = C.c_uint64
    L.akz_features_keypoints.argtypes = [vp]
    L.akz_features_keypoints.restype = vp
    L.akz_features_descriptors.argtypes = [vp]
    L.akz_features_descriptors.restype = vp
    L.akz_features_descriptor_len.argtypes = [vp]
    L.akz_features_descriptor_len.restype = C.c_uint32
    L.akz_features_num_levels.argtypes = [vp]
    L.akz_features_num_levels.restype = C.c_uint32
    L.akz_features_level_info.argtypes = [vp, C.c_uint32, C.POINTER(LevelInfo)]
    L.akz_features_fed_tau.argtypes = [vp, C.c_uint32, C.POINTER(C.c_double), C.c_uint32]
    L.akz_features_contrast_factor.argtypes = [vp]
    L.akz_features_contrast_factor.restype = C.c_double
    L.akz_features_num_candidates.argtypes = [vp]
    L.akz_features_num_candidates.restype = C.c_uint64
    L.akz_features_num_cache.argtypes = [vp]
    L.akz_features_num_cache.restype = C.c_uint64
    L.akz_features_evolution_download.argtypes = [vp, C.c_uint32, C.c_int, vp]
    L.akz_features_free.argtypes = [vp]
    L.akz_features_free.restype = None
    L.akz_match_top2.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.c_size_t, vp]
    L.akz_match_top2_device.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, vp]
    L.akz_merge_top2_device.argtypes = [vp, vp, C.c_uint32, C.c_uint64, vp]
    L.akz_descriptor_match.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.c_size_t, C.c_uint64,
                                       C.c_double, vp, C.POINTER(C.c_uint64)]
    L.akz_remove_outliers.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_int, vp,
                                      C.POINTER(C.c_uint64), vp]
    L.akz_comm_unique_id.argtypes = [vp]
    L.akz_context_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.akz_context_comm_init_all.argtypes = [C.POINTER(vp), C.c_int]
    L.akz_context_comm_destroy.argtypes = [vp]
    L.akz_match_top2_sharded_device.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, vp]
    L.akz_match_top2_sharded.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.c_size_t, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise AkazeError(rc, lib().akz_last_error().decode("utf-8", "replace"))


class EvolutionStep:
    """Mirror of types::evolution::EvolutionStep (evolution.rs:59-92); images are downloaded lazily
    from the device (only on an Engine created with keep_evolutions=True)."""

    def __init__(self, features, level, info, taus):
        self._f = features
        self._level = level
        self.etime = info.etime
        self.esigma = info.esigma
        self.octave = info.octave
        self.sublevel = info.sublevel
        self.sigma_size = info.sigma_size
        self.width = info.width
        self.height = info.height
        self.fed_tau_steps = taus

    def image(self, kind):
        return self._f.evolution(self._level, kind)

    def __getattr__(self, name):
        if name in IMAGE_KINDS:
            return self.image(name)
        raise AttributeError(name)


class Features:
    """Owns one akz_features handle: keypoints, descriptors and (optionally) the evolutions."""

    def __init__(self, handle, engine=None):
        L = lib()
        self._h = handle
        self._engine = engine  # keeps the engine alive while its evolutions can still be downloaded
        self.count = int(L.akz_features_count(handle))
        self._meta = None      # the other scalars are fetched on first use (a 1024-image batch creates 1024 of these)
        self._kp = self._desc = self._evo = None

    def _scalars(self):
        if self._meta is None:
            L, h = lib(), self._h
            if not h:
                raise AkazeError(1, "features were released before their metadata was read")
            self._meta = (L.akz_features_descriptor_len(h), L.akz_features_contrast_factor(h), L.akz_features_num_candidates(h),
                          L.akz_features_num_cache(h))
        return self._meta

    descriptor_len = property(lambda self: self._scalars()[0])
    contrast_factor = property(lambda self: self._scalars()[1])
    num_candidates = property(lambda self: self._scalars()[2])
    num_cache = property(lambda self: self._scalars()[3])

    # keypoints / descriptors are copied out of the library-owned (pinned) buffers on first access, so the
    # arrays stay valid after close(); `count` is free
    @property
    def keypoints(self):
        if self._kp is None:
            n = self.count
            if n:
                kp = (C.c_uint8 * (n * KEYPOINT_DTYPE.itemsize)).from_address(lib().akz_features_keypoints(self._h))
                self._kp = np.frombuffer(kp, KEYPOINT_DTYPE).copy()
            else:
                self._kp = np.zeros(0, KEYPOINT_DTYPE)
        return self._kp

    @property
    def descriptors_padded(self):
        if self._desc is None:
            n = self.count
            if n:
                d = (C.c_uint8 * (n * DESCRIPTOR_STRIDE)).from_address(lib().akz_features_descriptors(self._h))
                self._desc = np.frombuffer(d, np.uint8).reshape(n, DESCRIPTOR_STRIDE).copy()
            else:
                self._desc = np.zeros((0, DESCRIPTOR_STRIDE), np.uint8)
        return self._desc

    @property
    def descriptors(self):
        """Descriptor.vector of the reference has (162*channels+7)/8 bytes (descriptors.rs:42-46)."""
        return self.descriptors_padded[:, :self.descriptor_len]

    @property
    def evolutions(self):
        if self._evo is None:
            L = lib()
            self._evo = []
            for lv in range(L.akz_features_num_levels(self._h)):
                info = LevelInfo()
                _check(L.akz_features_level_info(self._h, lv, C.byref(info)))
                taus = np.zeros(info.n_steps, np.float64)
                if info.n_steps:
                    _check(L.akz_features_fed_tau(self._h, lv, taus.ctypes.data_as(C.POINTER(C.c_double)), info.n_steps))
                self._evo.append(EvolutionStep(self, lv, info, taus))
        return self._evo

    def evolution(self, level, kind):
        """One image of EvolutionStep `level`. Level 0 has no Lflow/Lstep: they stay 0x0 as in the reference
        (evolution.rs:118-119 allocates them empty, lib.rs:78 starts the diffusion loop at level 1)."""
        k = IMAGE_KINDS[kind] if isinstance(kind, str) else int(kind)
        if not self._h:
            raise AkazeError(1, "features were closed: evolution images are downloaded from the device on demand")
        if self._engine is not None and not self._engine._h:
            raise AkazeError(1, "the engine of these features was closed")
        ev = self.evolutions[level]
        if level == 0 and k in (IMAGE_KINDS["Lflow"], IMAGE_KINDS["Lstep"]):
            return np.zeros((0, 0), np.float32)
        out = np.empty((ev.height, ev.width), np.float32)
        _check(lib().akz_features_evolution_download(self._h, level, k, out.ctypes.data))
        return out

    def close(self):
        if self._h:
            # materialise what callers may still read after close (cheap; evolutions' images are not kept)
            self._scalars(), self.keypoints, self.descriptors_padded, self.evolutions  # noqa: B018
            lib().akz_features_free(self._h)
            self._h = None

    def release(self):
        """Frees the handle without copying anything out (throughput loops that only need `count`)."""
        if self._h:
            lib().akz_features_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _pad64(d):
    d = np.ascontiguousarray(d, np.uint8)
    if d.ndim != 2:
        raise ValueError("descriptors must be a 2-D uint8 array")
    if d.shape[1] > DESCRIPTOR_STRIDE:
        raise ValueError("descriptor rows are at most %d bytes, got %d" % (DESCRIPTOR_STRIDE, d.shape[1]))
    return d


def _gray2d(gray, dtype):
    g = np.ascontiguousarray(gray, dtype)
    if g.ndim != 2:
        raise ValueError("expected a 2-D gray image, got shape %r (convert with to_luma_u8 first)" % (g.shape,))
    return g


class Engine:
    """One akz_context: a B200 engine bound to a device (akz_create / akz_destroy)."""

    def __init__(self, device=0, max_width=4096, max_height=4096, max_batch=1, keep_evolutions=False,
                 max_candidates=None, max_keypoints=None):
        L = lib()
        h = C.c_void_p()
        _check(L.akz_create(device, max_width, max_height, max_batch, AKZ_KEEP_EVOLUTIONS if keep_evolutions else 0,
                            C.byref(h)))
        self._h = h
        self.device = device
        if max_candidates or max_keypoints:
            _check(L.akz_context_set_limits(h, max_candidates or 262144, max_keypoints or 65536))

    # -- housekeeping
    def close(self):
        if getattr(self, "_h", None):
            lib().akz_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self):
        """cudaStream_t the engine launches on (int), e.g. for torch.cuda.ExternalStream."""
        return lib().akz_context_stream(self._h)

    @property
    def launch_count(self):
        return lib().akz_context_launch_count(self._h)

    def set_sub_batch(self, images):
        """Images per pipeline sub-batch (akz_context_set_sub_batch)."""
        _check(lib().akz_context_set_sub_batch(self._h, int(images)))

    def set_match_path(self, path):
        """"auto" | "popc" | "tensor" (akz_context_set_match_path)."""
        _check(lib().akz_context_set_match_path(self._h, {"auto": 0, "popc": 1, "tensor": 2}[path]))

    def enable_timing(self, on=True):
        _check(lib().akz_context_enable_timing(self._h, int(on)))

    def stage_times(self, reset=False):
        """{stage: (milliseconds, launches)} accumulated by CUDA events since the last reset."""
        ms = (C.c_double * len(STAGES))()
        ln = (C.c_uint64 * len(STAGES))()
        _check(lib().akz_context_stage_times(self._h, ms, ln, int(reset)))
        return {name: (ms[i], ln[i]) for i, name in enumerate(STAGES)}

    # -- extraction
    def extract_u8(self, gray, config=None):
        gray = _gray2d(gray, np.uint8)
        cfg = config or Config.default()
        out = C.c_void_p()
        _check(lib().akz_extract_u8(self._h, gray.ctypes.data, gray.shape[1], gray.shape[0], gray.strides[0],
                                    C.byref(cfg), C.byref(out)))
        return Features(out, self)

    def extract_f32(self, unit_gray, config=None):
        img = _gray2d(unit_gray, np.float32)
        cfg = config or Config.default()
        out = C.c_void_p()
        _check(lib().akz_extract_f32(self._h, img.ctypes.data, img.shape[1], img.shape[0], C.byref(cfg), C.byref(out)))
        return Features(out, self)

    def extract_batch_u8(self, grays, config=None):
        """n images of one size: a list of 2-D uint8 arrays, or one C-contiguous (n, h, w) uint8 array (the cheap form for
        large batches: no per-image Python work, and the library uploads each pipeline sub-batch as a single copy)."""
        cfg = config or Config.default()
        if isinstance(grays, np.ndarray) and grays.ndim == 3:
            if grays.dtype != np.uint8 or not grays.flags.c_contiguous:
                grays = np.ascontiguousarray(grays, np.uint8)
            n, h, w = grays.shape
            if n == 0:
                return []
            addr = grays.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(h * w)
            ptrs = (C.c_void_p * n).from_buffer(addr)
            keep = grays
        else:
            keep = [_gray2d(g, np.uint8) for g in grays]
            n = len(keep)
            if n == 0:
                return []
            h, w = keep[0].shape
            if any(g.shape != (h, w) for g in keep):
                raise ValueError("all images of a batch must have the same size")
            ptrs = (C.c_void_p * n)(*[g.ctypes.data for g in keep])
        outs = (C.c_void_p * n)()
        _check(lib().akz_extract_batch_u8(self._h, n, ptrs, w, h, w, C.byref(cfg), outs))
        del keep
        return [Features(C.c_void_p(o), self) for o in outs]

    def extract_batch_u8_device(self, d_ptr, n, width, height, stride=None, config=None):
        """Images already in device memory (n*height*stride bytes); returns per-image keypoint counts."""
        cfg = config or Config.default()
        counts = (C.c_uint32 * n)()
        _check(lib().akz_extract_batch_u8_device(self._h, n, C.c_void_p(d_ptr), width, height, stride or width,
                                                 C.byref(cfg), counts))
        return np.frombuffer(counts, np.uint32).copy()

    def device_results(self):
        kp, ds, cap = C.c_void_p(), C.c_void_p(), C.c_uint32()
        _check(lib().akz_context_device_results(self._h, C.byref(kp), C.byref(ds), C.byref(cap)))
        return kp.value, ds.value, cap.value

    # -- matching
    def match_top2(self, q, db, desc_len=None):
        """Raw brute-force top-2 (feature_matching.rs:37-50); returns a TOP2_DTYPE array."""
        q, db = _pad64(q), _pad64(db)
        stride = q.shape[1]
        if db.shape[0] and db.shape[1] != stride:
            raise ValueError("query and database descriptors must have the same row length")
        desc_len = desc_len or min(stride, DESCRIPTOR_STRIDE)
        out = np.zeros(q.shape[0], TOP2_DTYPE)
        _check(lib().akz_match_top2(self._h, q.ctypes.data, q.shape[0], db.ctypes.data, db.shape[0], desc_len, stride,
                                    out.ctypes.data))
        return out

    def match_top2_device(self, d_q, nq, d_db, ndb, d_out, db_index_base=0):
        _check(lib().akz_match_top2_device(self._h, C.c_void_p(d_q), nq, C.c_void_p(d_db), ndb, db_index_base,
                                           C.c_void_p(d_out)))

    def merge_top2_device(self, d_parts, n_parts, nq, d_out):
        _check(lib().akz_merge_top2_device(self._h, C.c_void_p(d_parts), n_parts, nq, C.c_void_p(d_out)))

    def descriptor_match(self, d0, d1, distance_threshold=10000, lowes_ratio=0.86, desc_len=None):
        """ops::feature_matching::descriptor_match (feature_matching.rs:23-94)."""
        d0, d1 = _pad64(d0), _pad64(d1)
        stride = d0.shape[1]
        if d1.shape[0] and d1.shape[1] != stride:
            raise ValueError("both descriptor sets must have the same row length (%d vs %d)" % (stride, d1.shape[1]))
        desc_len = desc_len or min(stride, DESCRIPTOR_STRIDE)
        out = np.zeros(max(d0.shape[0], 1), MATCH_DTYPE)
        n = C.c_uint64()
        _check(lib().akz_descriptor_match(self._h, d0.ctypes.data, d0.shape[0], d1.ctypes.data, d1.shape[0], desc_len,
                                          stride, distance_threshold, lowes_ratio, out.ctypes.data, C.byref(n)))
        return out[:n.value].copy()

    def remove_outliers(self, keypoints_0, keypoints_1, matches, num_trials, epsilon_model, epsilon_inlier, sampling="reference",
                        return_model=False):
        """ops::estimate_fundamental_matrix::remove_outliers (estimate_fundamental_matrix.rs:99-165) on the GPU, all trials in
        parallel (akz_remove_outliers). sampling: "reference" = a fresh default random source per trial as the crate does (every
        trial draws the same eight matches), "advancing" = one source across the trials."""
        k0 = np.ascontiguousarray(keypoints_0, KEYPOINT_DTYPE)
        k1 = np.ascontiguousarray(keypoints_1, KEYPOINT_DTYPE)
        m = np.zeros(len(matches), MATCH_DTYPE)
        for f in MATCH_DTYPE.names:
            m[f] = matches[f]
        out = np.zeros(max(len(m), 1), MATCH_DTYPE)
        n = C.c_uint64()
        model = np.zeros(9, np.float32)
        _check(lib().akz_remove_outliers(self._h, k0.ctypes.data, len(k0), k1.ctypes.data, len(k1), m.ctypes.data, len(m), int(num_trials),
                                         float(epsilon_model), float(epsilon_inlier), {"reference": 0, "advancing": 1}[sampling],
                                         out.ctypes.data, C.byref(n), model.ctypes.data))
        res = out[:n.value].copy()
        return (res, model.reshape(3, 3)) if return_model else res

    # -- multi-GPU matching: database sharded by index over the ranks, NCCL all-gather + merge inside the library
    def comm_init(self, unique_id, rank, n_ranks):
        """Joins a communicator of n_ranks engines (one process per GPU); unique_id from comm_unique_id() of rank 0."""
        buf = (C.c_uint8 * COMM_UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        _check(lib().akz_context_comm_init(self._h, buf, int(rank), int(n_ranks)))

    def comm_destroy(self):
        _check(lib().akz_context_comm_destroy(self._h))

    def match_top2_sharded_device(self, d_q, nq, d_db_shard, ndb_shard, db_index_base, d_out):
        _check(lib().akz_match_top2_sharded_device(self._h, C.c_void_p(d_q), nq, C.c_void_p(d_db_shard), ndb_shard,
                                                   db_index_base, C.c_void_p(d_out)))


COMM_UNIQUE_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId through the library (akz_comm_unique_id): 128 bytes to ship to the other ranks."""
    buf = (C.c_uint8 * COMM_UNIQUE_ID_BYTES)()
    _check(lib().akz_comm_unique_id(buf))
    return bytes(buf)


def comm_init_all(engines):
    """One process driving several GPUs: a communicator over `engines` (one per device), engines[i] = rank i."""
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    _check(lib().akz_context_comm_init_all(arr, len(engines)))


def match_top2_sharded(engines, q, db, desc_len=None):
    """akz_match_top2_sharded: host buffers, the database sharded over the engines of one communicator."""
    q, db = _pad64(q), _pad64(db)
    stride = q.shape[1]
    if db.shape[0] and db.shape[1] != stride:
        raise ValueError("query and database descriptors must have the same row length")
    desc_len = desc_len or min(stride, DESCRIPTOR_STRIDE)
    out = np.zeros(q.shape[0], TOP2_DTYPE)
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    _check(lib().akz_match_top2_sharded(arr, len(engines), q.ctypes.data, q.shape[0], db.ctypes.data, db.shape[0], desc_len,
                                        stride, out.ctypes.data))
    return out


# ---- module-level mirror of the crate's two public functions ----------------------------------------
_default_engine = None


def default_engine(keep_evolutions=True):
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0, 8192, 8192, 1, keep_evolutions=keep_evolutions)
    return _default_engine


def to_luma_u8(rgb):
    """DynamicImage::to_luma of the `image` crate (^0.21): Rec.709 weights in f32, truncating cast.
    Third-party arithmetic that is not under /root/reference -- parity unpinned, and outside the
    drop-in boundary (the engine starts at the gray image)."""
    rgb = np.asarray(rgb)
    if rgb.ndim == 2:
        return rgb.astype(np.uint8)
    r = rgb[..., 0].astype(np.float32)
    g = rgb[..., 1].astype(np.float32)
    b = rgb[..., 2].astype(np.float32)
    lum = np.float32(0.2126) * r + np.float32(0.7152) * g + np.float32(0.0722) * b
    return lum.astype(np.uint8)


def load_gray(path):
    """image::open + to_luma (lib.rs:171, image.rs:128) on the host."""
    from PIL import Image
    with Image.open(path) as im:
        if im.mode in ("L", "1", "P", "I;16", "I", "F", "LA"):
            # gray sources: the `image` crate's to_luma is the identity on Luma8 (no weights, no truncation)
            if im.mode == "L":
                return np.asarray(im).copy()
            if im.mode == "LA":
                return np.asarray(im)[..., 0].copy()
            if im.mode != "P":
                return np.asarray(im.convert("L")).copy()
        return to_luma_u8(np.asarray(im.convert("RGB")))


def extract_features(input_image_path, options=None, engine=None):
    """akaze::extract_features (lib.rs:167-194): returns (evolutions, keypoints, descriptors)."""
    eng = engine or default_engine()
    gray = load_gray(input_image_path)
    f = eng.extract_u8(gray, options or Config.default())
    return f.evolutions, f.keypoints, f.descriptors


def descriptor_match(descriptors_0, descriptors_1, distance_threshold=10000, lowes_ratio=0.86, engine=None):
    eng = engine or default_engine()
    return eng.descriptor_match(descriptors_0, descriptors_1, distance_threshold, lowes_ratio)


def match_features(keypoints_0, descriptors_0, keypoints_1, descriptors_1, lowes_ratio, ransac_trials,
                   ransac_epsilon_inliers, engine=None, seed=None, ransac="host"):
    """akaze::match_features (lib.rs:252-275): brute-force matching on the GPU, then the reference's RANSAC outlier removal:
    ransac="host" runs the numpy restatement (ransac.py), "gpu" / "gpu-advancing" run akz_remove_outliers (all trials in parallel,
    with the crate's sampling or with one random source advanced across the trials)."""
    output = descriptor_match(descriptors_0, descriptors_1, 10000, lowes_ratio, engine=engine)
    if ransac in ("gpu", "gpu-advancing"):
        eng = engine or default_engine()
        return eng.remove_outliers(keypoints_0, keypoints_1, output, ransac_trials, 0.05, ransac_epsilon_inliers,
                                   sampling="advancing" if ransac == "gpu-advancing" else "reference")
    from . import ransac as ransac_host
    return ransac_host.remove_outliers(keypoints_0, keypoints_1, output, ransac_trials, 0.05, ransac_epsilon_inliers, seed=seed)
