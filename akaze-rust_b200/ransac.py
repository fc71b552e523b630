"""Host-side RANSAC outlier removal, the second half of akaze::match_features (lib.rs:267-274).

Restates akaze/src/ops/estimate_fundamental_matrix.rs in numpy float32. This step stays on the host by
design (tiny serial work after the GPU matcher). Third-party pieces that are
NOT under /root/reference and therefore parity-unpinned: nalgebra ^0.16 SVD (numpy/LAPACK here; singular
vectors are only defined up to sign/ordering) and the `random` ^0.12 default source, assumed to be
xorshift128+ seeded [42, 69]. Like the reference, a fresh default source is created on every trial
(estimate_fundamental_matrix.rs:118), so every trial draws the same eight matches; the reference then
iterates them in HashSet order, which is process-random -- here they are used in ascending order.
"""
import numpy as np

_MASK = (1 << 64) - 1


class _Xorshift128Plus:
    def __init__(self, seed=(42, 69)):
        self.s = [seed[0] & _MASK, seed[1] & _MASK]

    def read_u64(self):
        x, y = self.s
        self.s[0] = y
        x ^= (x << 23) & _MASK
        x ^= x >> 17
        x ^= y ^ (y >> 26)
        self.s[1] = x
        return (x + y) & _MASK


def estimate_fundamental_matrix(kp0, kp1, matches, epsilon):
    """estimate_fundamental_matrix.rs:17-69 (8 matches -> 3x3 or None)."""
    a = np.zeros((8, 9), np.float32)
    for i, m in enumerate(matches):
        x0, y0 = np.float32(kp0["x"][m["index_0"]]), np.float32(kp0["y"][m["index_0"]])
        x1, y1 = np.float32(kp1["x"][m["index_1"]]), np.float32(kp1["y"][m["index_1"]])
        a[i] = (x0 * x1, x0 * y1, x0, y0 * x1, y0 * y1, y0, x1, y1, 1.0)
    try:
        _, s, vt = np.linalg.svd(a, full_matrices=False)
    except np.linalg.LinAlgError:
        return None
    if int(np.sum(s > np.float32(epsilon))) != 8:
        return None
    i = int(np.argmin(s[:8]))
    v = vt[i]
    # Matrix3::new takes its arguments row by row (:55-65)
    return np.array([[v[0], v[3], v[6]], [v[1], v[4], v[7]], [v[2], v[5], v[8]]], np.float32)


def _errors(model, kp0, kp1, matches):
    """evaluate_model (:79-83) for all matches: |p_r^T F p_l|."""
    n = len(matches)
    pl = np.ones((n, 3), np.float32)
    pr = np.ones((n, 3), np.float32)
    pl[:, 0] = kp0["x"][matches["index_0"]]
    pl[:, 1] = kp0["y"][matches["index_0"]]
    pr[:, 0] = kp1["x"][matches["index_1"]]
    pr[:, 1] = kp1["y"][matches["index_1"]]
    return np.abs(np.einsum("ni,ij,nj->n", pr, model, pl)).astype(np.float32)


def remove_outliers(kp0, kp1, matches, num_trials, epsilon_model, epsilon_inlier, seed=None):
    """remove_outliers (:99-165)."""
    if len(matches) < 8:
        return matches.copy()
    max_inlier_count = 0
    final_model = np.zeros((3, 3), np.float32)
    cache = {}
    for _ in range(int(num_trials)):
        src = _Xorshift128Plus(seed or (42, 69))  # random::default() per trial (:118)
        chosen = set()
        while len(chosen) < 8:
            chosen.add(src.read_u64() % len(matches))
        key = tuple(sorted(chosen))
        if key not in cache:
            model = estimate_fundamental_matrix(kp0, kp1, matches[list(key)], epsilon_model)
            count = int(np.sum(_errors(model, kp0, kp1, matches) < np.float32(epsilon_inlier))) if model is not None else -1
            cache[key] = (model, count)
        model, count = cache[key]
        if model is not None and count > max_inlier_count:
            max_inlier_count = count
            final_model = model
    err = _errors(final_model, kp0, kp1, matches)
    return matches[err < np.float32(epsilon_inlier)].copy()
