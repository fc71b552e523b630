"""Generates tests/golden/*.npz + summary.json by running the CPU oracle on the reference's two test images.

The reference itself cannot run here (no cargo/rustc), so these are ORACLE outputs, not outputs of the Rust
crate: they pin the oracle against regressions and give the GPU tests fixed targets. Independent evidence
that they are right: the numpy/numba scratch restatement made during the survey
found the same 19 660 candidates, 7 395 / 5 629 keypoints and k0 = 0.005138 on these images.
Inputs are decoded with PIL and converted with the `image` crate's to_luma formula (akaze_rust_b200.to_luma_u8).

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import akaze_oracle as O  # noqa: E402
import akaze_rust_b200 as A  # noqa: E402  (only for load_gray: host-side decode helper)


def main():
    summary = {}
    res = {}
    for name in ("1", "2"):
        gray = A.load_gray(os.path.join(HERE, name + ".jpg"))
        r = O.extract(O.unit_float_from_u8(gray), threads=8)
        assert r.status == 0
        res[name] = r
        np.savez_compressed(os.path.join(HERE, "features_%s.npz" % name), keypoints=r.keypoints, descriptors=r.descriptors)
        lt_hash = [hashlib.sha256(r.image(l, "Lt").tobytes()).hexdigest()[:16] for l in range(r.num_levels)]
        ldet_hash = [hashlib.sha256(r.image(l, "Ldet").tobytes()).hexdigest()[:16] for l in range(r.num_levels)]
        summary[name] = {
            "gray_sha256": hashlib.sha256(gray.tobytes()).hexdigest(),
            "shape": list(gray.shape),
            "contrast_factor_hex": float(r.contrast_factor).hex(),
            "num_candidates": int(r.num_candidates),
            "num_cache": int(r.num_cache),
            "num_keypoints": int(len(r.keypoints)),
            "n_steps": [int(l["n_steps"]) for l in r.levels],
            "Lt_sha256_16": lt_hash,
            "Ldet_sha256_16": ldet_hash,
        }
    m = O.descriptor_match(res["1"].descriptors, res["2"].descriptors, 10000, 0.86)
    np.savez_compressed(os.path.join(HERE, "matches_1_2.npz"), matches=m)
    summary["matches_1_2"] = {"count": int(len(m)), "mean_distance": float(m["distance"].mean())}
    with open(os.path.join(HERE, "summary.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    print(json.dumps({k: (v if k == "matches_1_2" else {kk: v[kk] for kk in ("num_candidates", "num_cache", "num_keypoints")}) for k, v in summary.items()}))


if __name__ == "__main__":
    main()
