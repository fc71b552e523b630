"""Independent Python restatement of the reference's keypoint stages, written from the Rust sources (NOT from the C oracle):

    find_scale_space_extrema   akaze/src/ops/scale_space_extrema.rs:12-132
    do_subpixel_refinement     akaze/src/ops/scale_space_extrema.rs:141-189
    compute_main_orientation   akaze/src/ops/scale_space_extrema.rs:274-329
    get_mldb_descriptor        akaze/src/ops/descriptors.rs:37-175

All arithmetic is numpy float32 scalar arithmetic (IEEE, unfused, like rustc's); f32::cos / sin / atan2 / powf go to the
C library through ctypes, as Rust's std does. Used by tests/test_oracle_vs_numpy_keypoints.py to pin the C oracle's
keypoint stages with a second, independently written implementation that consumes the same evolution images.
"""
import ctypes
import ctypes.util

import numpy as np

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n, _k in (("cosf", 1), ("sinf", 1), ("atan2f", 2), ("powf", 2), ("roundf", 1), ("sqrtf", 1), ("ceilf", 1)):
    _f = getattr(_libm, _n)
    _f.restype = ctypes.c_float
    _f.argtypes = [ctypes.c_float] * _k


def cosf(x): return F(_libm.cosf(float(x)))
def sinf(x): return F(_libm.sinf(float(x)))
def atan2f(y, x): return F(_libm.atan2f(float(y), float(x)))
def powf(x, y): return F(_libm.powf(float(x), float(y)))
def roundf(x): return F(_libm.roundf(float(x)))  # f32::round: half away from zero


GAUSS25 = np.array([
    [0.02546481, 0.02350698, 0.01849125, 0.01239505, 0.00708017, 0.00344629, 0.00142946],
    [0.02350698, 0.02169968, 0.01706957, 0.01144208, 0.00653582, 0.00318132, 0.00131956],
    [0.01849125, 0.01706957, 0.01342740, 0.00900066, 0.00514126, 0.00250252, 0.00103800],
    [0.01239505, 0.01144208, 0.00900066, 0.00603332, 0.00344629, 0.00167749, 0.00069579],
    [0.00708017, 0.00653582, 0.00514126, 0.00344629, 0.00196855, 0.00095820, 0.00039744],
    [0.00344629, 0.00318132, 0.00250252, 0.00167749, 0.00095820, 0.00046640, 0.00019346],
    [0.00142946, 0.00131956, 0.00103800, 0.00069579, 0.00039744, 0.00019346, 0.00008024]], dtype=F)


def find_scale_space_extrema(ldets, levels, detector_threshold=0.001, derivative_factor=1.5):
    """scale_space_extrema.rs:12-132. ldets[i]: Ldet of level i (H x W f32); levels[i]: dict(octave, esigma).
    Keypoints are dicts with point (x, y), response, size, octave, class_id."""
    cache = []
    smax = F(10.0) * F(np.sqrt(F(2.0)))
    thr = F(detector_threshold)
    for e_id, (ldet, lv) in enumerate(zip(ldets, levels)):
        h, w = ldet.shape
        buf = ldet.reshape(-1)
        size = F(lv["esigma"] * derivative_factor)          # (esigma * derivative_factor) as f32  (:46)
        ratio = powf(2.0, F(lv["octave"]))
        sigma_size = roundf(size / ratio)
        # candidates in raster order; i runs over (w+1) .. (len - w - 1) exactly like the iterator code (:32)
        idx = np.arange(w + 1, buf.size - w - 1)
        v = buf[idx]
        cand = (idx % w != 0) & (v > thr) & (v > buf[idx + 1]) & (v > buf[idx - 1]) & (v > buf[idx - w]) & (v > buf[idx + w])
        for i in idx[cand]:
            x, y = int(i % w), int(i // w)
            resp = F(abs(buf[i]))
            px, py = F(x), F(y)
            # compare with the cache: same or previous class, first hit decides (:60-84)
            id_repeated, is_repeated, is_extremum = 0, False, True
            for k, pk in enumerate(cache):
                if e_id == pk["class_id"] or (e_id != 0 and e_id - 1 == pk["class_id"]):
                    dx = px * ratio - pk["x"]
                    dy = py * ratio - pk["y"]
                    dist = dx * dx + dy * dy
                    if dist <= size * size:
                        if resp > pk["response"]:
                            id_repeated, is_repeated = k, True
                        else:
                            is_extremum = False
                        break
            if not is_extremum:
                continue
            left_x = roundf(px - smax * sigma_size) - F(1)
            right_x = roundf(px + smax * sigma_size) + F(1)
            up_y = roundf(py - smax * sigma_size) - F(1)
            down_y = roundf(py + smax * sigma_size) + F(1)
            if left_x < 0 or right_x >= F(w) or up_y < 0 or down_y >= F(h):
                continue
            kp = dict(x=px * ratio + F(0.5) * (ratio - F(1.0)), y=py * ratio + F(0.5) * (ratio - F(1.0)), response=resp, size=size,
                      octave=int(lv["octave"]), class_id=e_id)
            if is_repeated:
                cache[id_repeated] = kp
            else:
                cache.append(kp)
    # filter with the upper scale level (:111-129): only later cache entries are looked at
    out = []
    for i, ki in enumerate(cache):
        rep = False
        for kj in cache[i:]:
            if ki["class_id"] + 1 == kj["class_id"]:
                dx, dy = ki["x"] - kj["x"], ki["y"] - kj["y"]
                if dx * dx + dy * dy <= ki["size"] * ki["size"]:
                    rep = True
                    break
        if not rep:
            out.append(ki)
    return out, len(cache)


def do_subpixel_refinement(kps, ldets):
    """:141-189. The LU solve's result is discarded by the reference, so b stays (-d_x, -d_y)."""
    res = []
    for kp in kps:
        ratio = powf(2.0, F(kp["octave"]))
        x = int(roundf(kp["x"] / ratio))
        y = int(roundf(kp["y"] / ratio))
        L = ldets[kp["class_id"]]
        d_x = F(0.5) * (L[y, x + 1] - L[y, x - 1])
        d_y = F(0.5) * (L[y + 1, x] - L[y - 1, x])
        b0, b1 = -d_x, -d_y
        if abs(b0) <= 1.0 and abs(b1) <= 1.0:
            q = dict(kp)
            q["x"] = (F(x) + b0) * ratio + F(0.5) * (ratio - F(1))
            q["y"] = (F(y) + b1) * ratio + F(0.5) * (ratio - F(1))
            res.append(q)
    return res


def compute_main_orientation(kp, lx, ly, octave):
    """:274-329, literally, including angs = atan2(res_y, res_y) and the never-reset sums."""
    ratio = F(1 << octave)
    s = roundf(F(0.5) * kp["size"] / ratio)
    xf, yf = kp["x"] / ratio, kp["y"] / ratio
    ids = [6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 5, 6]
    res_x, res_y, angs = [], [], []
    for i in range(-6, 7):
        for j in range(-6, 7):
            if i * i + j * j < 36:
                iy = int(roundf(yf + F(j) * s))
                ix = int(roundf(xf + F(i) * s))
                g = GAUSS25[ids[i + 6]][ids[j + 6]]
                rx, ry = g * lx[iy, ix], g * ly[iy, ix]
                res_x.append(rx)
                res_y.append(ry)
                angs.append(atan2f(ry, ry))
    PI = F(np.pi)
    ang1, sum_x, sum_y, mx, angle = F(0), F(0), F(0), F(0), F(0)
    while ang1 < F(2.0) * PI:
        ang2 = ang1 - F(5.0) * PI / F(3.0) if ang1 + PI / F(3.0) > F(2.0) * PI else ang1 + PI / F(3.0)
        ang1 = ang1 + F(0.15)
        for k in range(109):
            a = angs[k]
            if (ang1 < ang2 and ang1 < a and a < ang2) or (ang2 < ang1 and ((a > 0 and a < ang2) or (a > ang1 and a < F(2.0) * PI))):
                sum_x = sum_x + res_x[k]
                sum_y = sum_y + res_y[k]
        val = sum_x * sum_x + sum_y * sum_y
        if val > mx:
            mx = val
            angle = atan2f(sum_y, sum_x)
    return angle


def mldb_descriptor(kp, lt, lx, ly, channels=3, pattern=10):
    """descriptors.rs:37-175."""
    t = (6 + 36 + 120) * channels
    out = np.zeros((t + 7) // 8, np.uint8)
    ratio = F(1 << kp["octave"])
    scale = roundf(F(0.5) * kp["size"] / ratio)
    xf, yf = kp["x"] / ratio, kp["y"] / ratio
    co, si = cosf(kp["angle"]), sinf(kp["angle"])
    dpos = 0
    for lvl, mult in enumerate((F(1.0), F(2.0) / F(3.0), F(1.0) / F(2.0))):
        count = (lvl + 2) * (lvl + 2)
        step = int(_libm.ceilf(float(F(pattern) * mult)))
        values = []
        for i in range(-pattern, pattern, step):
            for j in range(-pattern, pattern, step):
                di, dx, dy, ns = F(0), F(0), F(0), 0
                for k in range(i, i + step):
                    for l in range(j, j + step):
                        lf, kf = F(l) + F(0.5), F(k) + F(0.5)
                        sy = yf + (lf * co * scale + kf * si * scale)
                        sx = xf + (-lf * si * scale + kf * co * scale)
                        y1, x1 = int(roundf(sy)), int(roundf(sx))
                        di = di + lt[y1, x1]
                        rx, ry = lx[y1, x1], ly[y1, x1]
                        if channels == 2:
                            dx = dx + F(np.sqrt(rx * rx + ry * ry))
                        elif channels == 3:
                            dx = dx + (-rx * si + ry * co)
                            dy = dy + (rx * co + ry * si)
                        ns += 1
                values.append((di / F(ns), dx / F(ns), dy / F(ns))[:channels])
        for pos in range(channels):
            for i in range(count):
                for j in range(i + 1, count):
                    if values[i][pos] > values[j][pos]:
                        out[dpos >> 3] |= np.uint8(1 << (dpos & 7))
                    dpos += 1
    return out


def find_scale_space_extrema_vec(ldets, levels, detector_threshold=0.001, derivative_factor=1.5):
    """Same semantics as find_scale_space_extrema, with the linear cache scan of every candidate done by numpy (first
    matching slot via argmax over the whole cache), so that the reference's full-size test images are affordable."""
    cap = 1 << 16
    cx, cy, cr, cs = (np.zeros(cap, F) for _ in range(4))
    cc = np.full(cap, -10, np.int64)
    co = np.zeros(cap, np.int64)
    n = 0
    smax = F(10.0) * F(np.sqrt(F(2.0)))
    thr = F(detector_threshold)
    for e_id, (ldet, lv) in enumerate(zip(ldets, levels)):
        h, w = ldet.shape
        buf = ldet.reshape(-1)
        size = F(lv["esigma"] * derivative_factor)
        ratio = powf(2.0, F(lv["octave"]))
        sigma_size = roundf(size / ratio)
        idx = np.arange(w + 1, buf.size - w - 1)
        v = buf[idx]
        cand = (idx % w != 0) & (v > thr) & (v > buf[idx + 1]) & (v > buf[idx - 1]) & (v > buf[idx - w]) & (v > buf[idx + w])
        s2 = size * size
        for i in idx[cand]:
            x, y = int(i % w), int(i // w)
            resp = F(abs(buf[i]))
            px, py = F(x), F(y)
            qx, qy = px * ratio, py * ratio
            id_repeated, is_repeated, is_extremum = 0, False, True
            if n:
                cls = cc[:n]
                dx = qx - cx[:n]
                dy = qy - cy[:n]
                hit = ((cls == e_id) | (cls == e_id - 1)) & (dx * dx + dy * dy <= s2)
                if hit.any():
                    k = int(np.argmax(hit))  # the FIRST slot that matches decides (:60-84)
                    if resp > cr[k]:
                        id_repeated, is_repeated = k, True
                    else:
                        is_extremum = False
            if not is_extremum:
                continue
            if (roundf(px - smax * sigma_size) - F(1) < 0 or roundf(px + smax * sigma_size) + F(1) >= F(w)
                    or roundf(py - smax * sigma_size) - F(1) < 0 or roundf(py + smax * sigma_size) + F(1) >= F(h)):
                continue
            k = id_repeated if is_repeated else n
            cx[k], cy[k] = qx + F(0.5) * (ratio - F(1.0)), qy + F(0.5) * (ratio - F(1.0))
            cr[k], cs[k], cc[k], co[k] = resp, size, e_id, int(lv["octave"])
            if not is_repeated:
                n += 1
    out = []
    for i in range(n):
        later = slice(i, n)
        dx, dy = cx[i] - cx[later], cy[i] - cy[later]
        if not ((cc[later] == cc[i] + 1) & (dx * dx + dy * dy <= cs[i] * cs[i])).any():
            out.append(dict(x=cx[i], y=cy[i], response=cr[i], size=cs[i], octave=int(co[i]), class_id=int(cc[i])))
    return out, n
