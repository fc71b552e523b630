"""Debug / visualisation outputs of the crate, host code on top of the engine's results:

    normalize, create_dynamic_image, save, random_color, draw_circle, draw_line   akaze/src/types/image.rs:148-210, 385-480
    write_evolutions, build_path                                                  akaze/src/types/evolution.rs:163-218
    draw_keypoints_to_image, draw_keypoints                                       akaze/src/types/keypoint.rs:52-72
    draw_matches                                                                  akaze/src/types/feature_match.rs:17-82

Never performance relevant: these exist so that a user of the crate finds the same helpers (the integration test dumps the
scale space and the keypoint / match overlays when AKAZE_SCALE_SPACE_DIR is set, akaze/tests/integration-test.rs:49-64,
94-121). Arithmetic follows the reference where it decides pixels: f32 normalisation, truncating casts, the circle's
half-open pixel ranges and `<=` radius test, the 50 % blend. Two behaviours of the reference are kept on purpose:
`build_path` calls `set_extension(".png")` on a name that already ends in ".png", which yields "Lt_00000..png"; and
`random_color` builds a fresh default `random` source (xorshift128+ seeded [42, 69], third-party and unpinned) on every
call, so every circle and line gets the same colour.
"""
import os

import numpy as np

from .ransac import _Xorshift128Plus

f32 = np.float32


def normalize(image):
    """types/image.rs:163-194: (pixel - min) / (max - min) in f32 (NaN / inf when the image is flat, like the reference)."""
    a = np.asarray(image, f32)
    if a.size == 0:
        return a.copy()
    mn, mx = f32(a.min()), f32(a.max())
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((a - mn) / f32(mx - mn)).astype(f32)


def create_dynamic_image(image):
    """types/image.rs:148-161: Luma8 with (value * 255f32) as u8 -- a saturating, truncating cast (NaN -> 0)."""
    a = np.asarray(image, f32) * f32(255.0)
    with np.errstate(invalid="ignore"):
        a = np.where(np.isnan(a), f32(0.0), np.clip(a, 0.0, 255.0))
    return a.astype(np.uint8)  # truncation toward zero on [0, 255]


def save(image, path):
    """types/image.rs:201-207: normalise and write; 0x0 images (level 0's Lflow / Lstep) are skipped."""
    a = np.asarray(image, f32)
    if a.ndim != 2 or a.shape[0] == 0 or a.shape[1] == 0:
        return False
    from PIL import Image
    Image.fromarray(create_dynamic_image(normalize(a)), mode="L").save(path, format="PNG")
    return True


def build_path(destination_dir, path_label, idx):
    """types/evolution.rs:163-168, including what set_extension(".png") does to a name that already ends in .png."""
    return os.path.join(str(destination_dir), "%s%05d..png" % (path_label, idx))


EVOLUTION_IMAGES = ("Lt", "Lsmooth", "Lx", "Ly", "Lxx", "Lyy", "Lxy", "Lflow", "Lstep", "Ldet")


def write_evolutions(evolutions, destination_dir):
    """types/evolution.rs:175-218: all ten images of every EvolutionStep as normalised PNGs. `evolutions` is what
    extract_features returns (the engine must keep evolutions for the six non-persistent images; missing ones are skipped)."""
    os.makedirs(str(destination_dir), exist_ok=True)
    written = []
    for i, e in enumerate(evolutions):
        for name in EVOLUTION_IMAGES:
            try:
                img = getattr(e, name)
            except Exception:  # noqa: BLE001 -- the engine's AkazeError for an image that is not resident
                continue
            p = build_path(destination_dir, name + "_", i)
            if img is not None and save(img, p):
                written.append(p)
    return written


def random_color():
    """types/image.rs:385-392: three u8 reads from a FRESH default source -> the same colour every call."""
    src = _Xorshift128Plus()
    return tuple(int(src.read_u64() & 0xFF) for _ in range(3))


def _blend(p1, p2):
    """types/image.rs:402-408."""
    return ((p1.astype(f32) + p2.astype(f32)) / f32(2.0)).astype(np.uint8)


def draw_circle(image, point, rgb, radius):
    """types/image.rs:419-445 on an (h, w, 3) uint8 array, in place. Like the reference it indexes the image without a bounds
    check of its own: pixels outside the image raise (the reference panics)."""
    px, py, r = f32(point[0]), f32(point[1]), f32(radius)
    cx, cy, ri = int(px), int(py), int(r)
    x0, x1 = max(cx - ri, 0), cx + ri  # saturating_sub .. saturating_add, upper bound exclusive
    y0, y1 = max(cy - ri, 0), cy + ri
    if x1 <= x0 or y1 <= y0:
        return
    h, w = image.shape[:2]
    if x1 > w or y1 > h:
        raise IndexError("draw_circle: pixel outside the image (the reference panics here)")
    xs = np.arange(x0, x1, dtype=f32)[None, :] - px
    ys = np.arange(y0, y1, dtype=f32)[:, None] - py
    inside = np.sqrt(xs * xs + ys * ys).astype(f32) <= r
    patch = image[y0:y1, x0:x1]
    patch[inside] = _blend(np.broadcast_to(np.asarray(rgb, np.uint8), patch.shape)[inside], patch[inside])


def draw_line(image, point_0, point_1, rgb, radius):
    """types/image.rs:455-480 (a circle per step along x; vertical lines degenerate exactly like the reference's)."""
    p0 = (f32(point_0[0]), f32(point_0[1]))
    p1 = (f32(point_1[0]), f32(point_1[1]))
    dx, dy = f32(p1[0] - p0[0]), f32(p1[1] - p0[1])
    if abs(dx) <= 1.0 and abs(dy) <= 1.0:
        draw_circle(image, p0, rgb, radius)
        return
    with np.errstate(divide="ignore", invalid="ignore"):
        m = f32(dy / dx)
        b = f32(p0[1] - m * p0[0])
    x_0, x_n = min(p0[0], p1[0]), max(p0[0], p1[0])
    num_points = max(max(abs(dx), abs(dy)), f32(2.0))
    x_step = f32(abs(dx) / num_points)
    x = f32(x_0)
    while x <= x_n:
        y = f32(m * x + b)
        if np.isfinite(y):
            draw_circle(image, (x, y), rgb, radius)
        if x_step == 0.0:  # vertical line: the reference's loop would never advance; one column of circles is all it can mean
            break
        x = f32(x + x_step)


def draw_keypoints_to_image(image, keypoints):
    """types/keypoint.rs:52-56: a circle of radius keypoint.size per keypoint, in place."""
    for k in keypoints:
        draw_circle(image, (k["x"], k["y"]), random_color(), k["size"])


def draw_keypoints(image, keypoints):
    """types/keypoint.rs:68-72: RGB copy of the image with the keypoints drawn."""
    a = np.asarray(image)
    rgb = np.stack([a, a, a], axis=-1).astype(np.uint8) if a.ndim == 2 else np.ascontiguousarray(a[..., :3], np.uint8).copy()
    draw_keypoints_to_image(rgb, keypoints)
    return rgb


def draw_matches(image_0, image_1, keypoints_0, keypoints_1, matches):
    """types/feature_match.rs:32-82: both images side by side, one line per match."""
    a0 = draw_keypoints(image_0, [])
    a1 = draw_keypoints(image_1, [])
    half = max(a0.shape[1], a1.shape[1])
    height = max(a0.shape[0], a1.shape[0])
    out = np.zeros((height, 2 * half, 3), np.uint8)
    out[:a0.shape[0], :a0.shape[1]] = a0
    out[:a1.shape[0], half:half + a1.shape[1]] = a1
    for m in matches:
        k0, k1 = keypoints_0[int(m["index_0"])], keypoints_1[int(m["index_1"])]
        draw_line(out, (k0["x"], k0["y"]), (f32(k1["x"]) + f32(half), k1["y"]), random_color(), f32(height) / f32(500.0))
    return out
