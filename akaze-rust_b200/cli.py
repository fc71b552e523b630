"""The three akaze-util command-line tools on the B200 engine, argument for argument:

    python -m akaze_rust_b200.cli extract_features  INPUT OUTPUT [-d DIRECTORY] [-o PATH]      akaze-util/src/bin/extract_features.rs
    python -m akaze_rust_b200.cli match_features    EXTRACTIONS_0 EXTRACTIONS_1 OUTPUT [-t F]  akaze-util/src/bin/match_features.rs
    python -m akaze_rust_b200.cli extract_and_match INPUT_0 INPUT_1 OUTPUT_PREFIX [-m IMAGE]   akaze-util/src/bin/extract_and_match.rs

File formats are akaze-util's (formats.py: ".json" -> serde_json layout, anything else -> bincode 1.1), the option file
of `-o` is the serde_json form of `Config` and is WRITTEN with the defaults when it does not exist, like the reference
(extract_features.rs:69-84). The matcher parameters are the reference's constants (0.86, 1000 RANSAC trials, 3.0 px);
`-t` is parsed and logged but, as in the reference, not used (match_features.rs:57-78). `-d` writes the scale space as
normalised PNGs and the keypoint overlay, `-m` the match image (visualize.py mirrors the crate's drawing helpers; Lt and
Ldet are also written as float32 .npy for numeric inspection).
"""
import argparse
import json
import logging
import os
import sys

import numpy as np

log = logging.getLogger("akaze")

CONFIG_FIELDS = ("num_sublevels", "max_octave_evolution", "base_scale_offset", "initial_contrast", "contrast_percentile",
                 "contrast_factor_num_bins", "derivative_factor", "detector_threshold", "descriptor_channels", "descriptor_pattern_size")


def _api():
    import akaze_rust_b200 as A
    return A


def load_options(path, A=None):
    """extract_features.rs:66-87: read the JSON options file, or create it with the defaults if it does not exist."""
    A = A or _api()
    options = A.Config.default()
    if path is None:
        return options
    if os.path.exists(path):
        log.info("Reading options file from %s", path)
        with open(path) as fh:
            doc = json.load(fh)
        missing = [f for f in CONFIG_FIELDS if f not in doc]
        if missing:  # serde would fail on a missing field: so do we
            raise ValueError("options file %s lacks %s" % (path, ", ".join(missing)))
        for f in CONFIG_FIELDS:
            setattr(options, f, type(getattr(options, f))(doc[f]))
    else:
        with open(path, "w") as fh:
            json.dump({f: getattr(options, f) for f in CONFIG_FIELDS}, fh, separators=(",", ":"))
        log.info("Writing options file from %s", path)
    return options


def cmd_extract_features(args, A=None):
    A = A or _api()
    from akaze_rust_b200 import formats
    log.info("Input image path is %s, output extractions path is %s.", args.INPUT, args.OUTPUT)
    options = load_options(args.options, A)
    evolutions, keypoints, descriptors = A.extract_features(args.INPUT, options)
    formats.serialize_features_to_file(keypoints, descriptors, args.OUTPUT)
    log.info("Done, extracted %d features.", len(keypoints))
    if args.debug_path:
        # extract_features.rs:88-102: the scale space as normalised PNGs (write_evolutions) and the keypoint overlay
        from akaze_rust_b200 import visualize
        os.makedirs(args.debug_path, exist_ok=True)
        written = visualize.write_evolutions(evolutions, args.debug_path)
        for i, e in enumerate(evolutions):  # the float images themselves, for numeric inspection
            for name in ("Lt", "Ldet"):
                img = getattr(e, name, None)
                if img is not None and np.size(img):
                    np.save(os.path.join(args.debug_path, "%s_%02d.npy" % (name, i)), np.asarray(img, np.float32))
        from akaze_rust_b200 import load_gray  # host code: decode + to_luma
        from PIL import Image
        try:
            Image.fromarray(visualize.draw_keypoints(load_gray(args.INPUT), keypoints)).save(os.path.join(args.debug_path, "keypoints.png"))
        except IndexError as e:  # a circle left the image: the reference panics here; the dump above is already written
            log.warning("keypoint overlay not written: %s", e)
        log.info("Wrote %d scale-space images to %s.", len(written), args.debug_path)
    return 0


def _match(A, f0, f1):
    k0, d0 = f0
    k1, d1 = f1
    width = len(d0[0]) if len(d0) else (len(d1[0]) if len(d1) else 61)
    d0 = np.stack(d0) if len(d0) else np.zeros((0, width), np.uint8)
    d1 = np.stack(d1) if len(d1) else np.zeros((0, width), np.uint8)
    return A.match_features(k0, d0, k1, d1, 0.86, 1000, 3.0)


def cmd_match_features(args, A=None):
    A = A or _api()
    from akaze_rust_b200 import formats
    log.info("Input extractions: %s/%s, output matches: %s, threshold: %s.", args.INPUT_EXTRACTIONS_0, args.INPUT_EXTRACTIONS_1, args.OUTPUT,
             args.threshold)
    m = _match(A, formats.deserialize_features_from_file(args.INPUT_EXTRACTIONS_0), formats.deserialize_features_from_file(args.INPUT_EXTRACTIONS_1))
    formats.serialize_matches_to_file(m, args.OUTPUT)
    log.debug("Done, got %d matches.", len(m))
    return 0


def cmd_extract_and_match(args, A=None):
    A = A or _api()
    from akaze_rust_b200 import formats
    options = A.Config.default()
    paths = [args.OUTPUT_PREFIX + s for s in ("-extractions_0.cbor", "-extractions_1.cbor", "-matches.cbor")]  # the reference's names
    feats = []
    for i, inp in enumerate((args.INPUT_0, args.INPUT_1)):
        _e, k, d = A.extract_features(inp, options)
        formats.serialize_features_to_file(k, d, paths[i])
        log.info("Done, extracted %d features from image %d.", len(k), i)
        feats.append((k, d))
    m = A.match_features(feats[0][0], feats[0][1], feats[1][0], feats[1][1], 0.86, 1000, 3.0)
    log.info("Got %d matches.", len(m))
    formats.serialize_matches_to_file(m, paths[2])
    if args.match_image:  # extract_and_match.rs:110-121
        from akaze_rust_b200 import load_gray, visualize
        from PIL import Image
        img = visualize.draw_matches(load_gray(args.INPUT_0), load_gray(args.INPUT_1), feats[0][0], feats[1][0], m)
        Image.fromarray(img).save(args.match_image)
        log.info("Wrote the match image to %s.", args.match_image)
    return 0


def build_parser():
    ap = argparse.ArgumentParser(prog="akaze_rust_b200.cli", description="akaze-util's command line tools on the B200 engine")
    sub = ap.add_subparsers(dest="tool", required=True)
    p = sub.add_parser("extract_features", help="KAZE extractor.")
    p.add_argument("INPUT", help="The input image.")
    p.add_argument("OUTPUT", help="The output extractions. Extension can be JSON or CBOR.")
    p.add_argument("-d", "--debug_path", metavar="DIRECTORY", help="Sets a directory to write debug information to.")
    p.add_argument("-o", "--options", metavar="PATH", help="A JSON file containing options.")
    p.set_defaults(fn=cmd_extract_features)
    p = sub.add_parser("match_features", help="Match two extraction files.")
    p.add_argument("INPUT_EXTRACTIONS_0", help="The input extraction results for image 0.")
    p.add_argument("INPUT_EXTRACTIONS_1", help="The input extraction results for image 1.")
    p.add_argument("OUTPUT", help="The output matches.")
    p.add_argument("-t", "--threshold", metavar="FLOAT", type=float, default=10.0, help="The distance threshold for the matcher.")
    p.set_defaults(fn=cmd_match_features)
    p = sub.add_parser("extract_and_match", help="Extract and match KAZE image features.")
    p.add_argument("INPUT_0", help="The first input image.")
    p.add_argument("INPUT_1", help="The second input image.")
    p.add_argument("OUTPUT_PREFIX", help="The output prefix for all files.")
    p.add_argument("-m", "--match_image", metavar="IMAGE_FILE_PATH", help="Sets a path to write the match image to.")
    p.set_defaults(fn=cmd_extract_and_match)
    return ap


def main(argv=None, A=None):
    level = os.environ.get("AKAZE_LOG", "info").upper()  # the reference's env_logger variable
    logging.basicConfig(level=getattr(logging, level, logging.INFO), format="%(levelname)s %(message)s")
    args = build_parser().parse_args(argv)
    return args.fn(args, A)


if __name__ == "__main__":
    sys.exit(main())
