"""Independent numpy restatement of the stencil stages in CLOSED FORM (test infrastructure).

The C oracle (oracle/akaze_oracle.c) follows the reference's loop structure literally (flattened-buffer
sweeps, per-case border code). This module restates the same stages the way the CUDA kernels compute
them -- valid-interior correlation + clamp-replicated borders, antisymmetric edge fluxes
with zero flux across the image border -- so that tests can prove the two formulations are bit-identical
before the GPU is involved. All arithmetic is float32 with explicit operation order (numpy never fuses).
"""
import numpy as np

f32 = np.float32


def correlate_clamped(img, kernel, axis):
    img = np.asarray(img, f32)
    kernel = np.asarray(kernel, f32)
    hw = len(kernel) // 2
    H, W = img.shape
    ys = np.clip(np.arange(H), hw, H - 1 - hw)
    xs = np.clip(np.arange(W), hw, W - 1 - hw)
    cy, cx = np.meshgrid(ys, xs, indexing="ij")
    acc = np.zeros((H, W), f32)
    for i, kv in enumerate(kernel):
        d = i - hw
        if axis == 1:
            v = img[cy, cx + d]
        else:
            v = img[cy + d, cx]
        acc = acc + kv * v
    return acc


def gaussian_kernel(r, size):
    """types/image.rs:341-365 in float32."""
    r = f32(r)
    hw = size // 2
    k = np.zeros(size, f32)
    s = f32(0)
    for i in range(-hw, hw + 1):
        x = f32(i)
        a = f32(1) / (np.sqrt(f32(2) * f32(np.pi)) * r)
        # expf: glibc's is correctly rounded except in vanishingly rare cases; exp in f64 then one
        # rounding reproduces it (numpy's own float32 exp is a SIMD approximation, 1 ulp off here)
        e = f32(np.exp(np.float64(-(x * x) / (f32(2) * (r * r)))))
        val = f32(a * e)
        k[i + hw] = val
        s = f32(s + val)
    return (k / s).astype(f32)


def gaussian_blur(img, r):
    size = int(np.ceil(r)) * 2 + 1
    k = gaussian_kernel(r, size)
    return correlate_clamped(correlate_clamped(img, k, 1), k, 0)


def scharr_main(s):
    w = 10.0 / 3.0
    norm = 1.0 / (2.0 * float(s) * (w + 2.0))
    k = np.zeros(2 * s + 1, f32)
    k[0] = f32(norm)
    k[s] = f32(w * norm)
    k[-1] = f32(norm)
    return k


def scharr_off(s):
    k = np.zeros(2 * s + 1, f32)
    k[0] = -1
    k[-1] = 1
    return k


def scharr_x(img, s):
    """'x_order' (derivatives.rs:41-47): main-axis kernel along x, then difference along y."""
    return correlate_clamped(correlate_clamped(img, scharr_main(s), 1), scharr_off(s), 0)


def scharr_y(img, s):
    """'y_order' (derivatives.rs:59-65)."""
    return correlate_clamped(correlate_clamped(img, scharr_off(s), 1), scharr_main(s), 0)


def half_size(img):
    img = np.asarray(img, f32)
    H, W = img.shape
    h, w = H // 2, W // 2
    a = img[0:2 * h:2, 0:2 * w:2]
    b = img[1:2 * h:2, 0:2 * w:2]
    c = img[0:2 * h:2, 1:2 * w:2]
    d = img[1:2 * h:2, 1:2 * w:2]
    return ((((f32(0) + a) + b) + c) + d) / f32(4)


def pm_g2(lx, ly, k):
    lx = lx.astype(np.float64)
    ly = ly.astype(np.float64)
    inv = 1.0 / (k * k)
    return (1.0 / (1.0 + inv * (lx * lx + ly * ly))).astype(f32)


def fed_step(L, c, tau):
    """nonlinear_diffusion.rs:15-144 with one flux per edge: fE(x)=(c+cE)(LE-L), fW(x)=fE(x-1),
    flux across the image border = 0; Lstep = (0.5*tau)*(((fE-fW)+fS)-fN)."""
    L = np.asarray(L, f32)
    c = np.asarray(c, f32)
    H, W = L.shape
    fE = np.zeros((H, W), f32)
    fS = np.zeros((H, W), f32)
    fE[:, :-1] = (c[:, :-1] + c[:, 1:]) * (L[:, 1:] - L[:, :-1])
    fS[:-1, :] = (c[:-1, :] + c[1:, :]) * (L[1:, :] - L[:-1, :])
    fW = np.zeros((H, W), f32)
    fN = np.zeros((H, W), f32)
    fW[:, 1:] = fE[:, :-1]
    fN[1:, :] = fS[:-1, :]
    ht = f32(0.5) * f32(tau)
    step = ht * (((fE - fW) + fS) - fN)
    return (L + step).astype(f32), step.astype(f32)


def contrast_factor(img, percentile=0.7, nbins=300):
    """contrast_factor.rs:18-71."""
    g = gaussian_blur(img, 1.0)
    lx = scharr_x(g, 1).astype(np.float64)[1:-1, 1:-1]
    ly = scharr_y(g, 1).astype(np.float64)[1:-1, 1:-1]
    modg = np.sqrt(lx * lx + ly * ly)
    hmax = modg.max()
    nz = modg[modg != 0.0]
    bins = np.floor(nbins * (nz / hmax)).astype(np.int64)
    bins[bins == nbins] = nbins - 1
    hist = np.bincount(bins, minlength=nbins)
    thr = int(len(nz) * percentile)
    k = 0
    n = 0
    while n < thr and k < nbins:
        n += int(hist[k])
        k += 1
    if n >= thr:
        return hmax * float(k) / float(nbins)
    return 0.03


def detector(lsmooth, s):
    """detector_response.rs:8-14, 38-55."""
    lx = scharr_x(lsmooth, s)
    ly = scharr_y(lsmooth, s)
    lxx = scharr_x(lx, s)
    lyy = scharr_y(ly, s)
    lxy = scharr_y(lx, s)
    ldet = ((lxx * lyy) - (lxy * lxy)) * f32(s ** 4)
    return lx, ly, lxx, lyy, lxy, ldet.astype(f32)


def synthetic_image(h, w, seed, n_rect=None):
    """Deterministic textured, corner-rich u8 test image (numpy PCG64)."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.Generator(np.random.PCG64(seed))
    a = gaussian_filter(rng.standard_normal((h, w)), 1.5)
    b = gaussian_filter(rng.standard_normal((h, w)), 8.0)
    img = 128.0 + 40.0 * a / a.std() + 60.0 * b / b.std()
    if n_rect is None:
        n_rect = max(4, (h * w) // 10000)
    for _ in range(n_rect):
        rw, rh = rng.integers(8, 121, size=2)
        x0 = rng.integers(0, max(1, w - 8))
        y0 = rng.integers(0, max(1, h - 8))
        img[y0:y0 + rh, x0:x0 + rw] = rng.uniform(0, 255)
    img = gaussian_filter(img, 0.7)
    return np.clip(img, 0, 255).astype(np.uint8)


def natural_image(h, w, seed):
    """Photo-like density (bench.py's workload recipe in numpy): five octaves of blurred noise with
    amplitudes (4,8,16,32,48) at sigma (1.5,3,6,12,24), 48 rectangles per 2 Mpx, sigma-1 blur."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.Generator(np.random.PCG64(seed))
    img = np.full((h, w), 128.0)
    for a, s in zip((4, 8, 16, 32, 48), (1.5, 3, 6, 12, 24)):
        n = gaussian_filter(rng.standard_normal((h, w)), s)
        img += a * n / n.std()
    for _ in range(max(2, int(round(48 * h * w / (1920.0 * 1080.0))))):
        rw, rh = rng.integers(8, 121, size=2)
        x0 = rng.integers(0, max(1, w - 8))
        y0 = rng.integers(0, max(1, h - 8))
        img[y0:y0 + rh, x0:x0 + rw] = rng.uniform(0, 255)
    img = gaussian_filter(img, 1.0)
    return np.clip(img, 0, 255).astype(np.uint8)
