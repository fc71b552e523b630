"""CPU-only: the akaze-util command line mirror (akaze-rust_b200/cli.py). The engine needs a GPU,
so these tests hand the CLI a stand-in API whose extract_features / descriptor matching come from the CPU oracle (test
infrastructure); what is checked is the tools' own logic: arguments, option-file semantics, file names and formats."""
import json
import os
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture()
def fake_api(akz, oracle):
    from akaze_rust_b200 import ransac
    seen = {}

    def extract_features(path, options=None):
        seen["options"] = options.to_dict() if options is not None else None
        gray = akz.load_gray(path)[:480, :640]  # a crop keeps the CPU oracle fast
        cfg = oracle.default_config()
        if options is not None:
            for n, _ in cfg._fields_:
                setattr(cfg, n, getattr(options, n))
        r = oracle.extract(oracle.unit_float_from_u8(np.ascontiguousarray(gray)), cfg, threads=4)
        evo = [types.SimpleNamespace(Lt=r.image(l, "Lt"), Ldet=r.image(l, "Ldet")) for l in range(r.num_levels)]
        return evo, r.keypoints, r.descriptors

    def match_features(k0, d0, k1, d1, lowes, trials, eps):
        put = oracle.descriptor_match(np.ascontiguousarray(d0), np.ascontiguousarray(d1), 10000, lowes)
        seen["putative"] = put
        m = np.zeros(len(put), akz.MATCH_DTYPE)
        for f in m.dtype.names:
            m[f] = put[f]
        return ransac.remove_outliers(k0, k1, m, trials, 0.05, eps)

    return types.SimpleNamespace(Config=akz.Config, extract_features=extract_features, match_features=match_features, seen=seen)


def test_extract_features_tool(tmp_path, fake_api):
    from akaze_rust_b200 import cli, formats
    out = tmp_path / "f.bin"
    opt = tmp_path / "options.json"
    dbg = tmp_path / "dbg"
    # 1st run: the options file does not exist -> it is written with Config::default() (extract_features.rs:79-84)
    assert cli.main(["extract_features", os.path.join(GOLD, "1.jpg"), str(out), "-o", str(opt), "-d", str(dbg)], fake_api) == 0
    doc = json.loads(opt.read_text())
    assert doc == {"num_sublevels": 4, "max_octave_evolution": 4, "base_scale_offset": 1.6, "initial_contrast": 0.001,
                   "contrast_percentile": 0.7, "contrast_factor_num_bins": 300, "derivative_factor": 1.5, "detector_threshold": 0.001,
                   "descriptor_channels": 3, "descriptor_pattern_size": 10}
    k, d = formats.deserialize_features_from_file(str(out))
    assert len(k) == len(d) > 300 and all(len(v) == 61 for v in d)
    assert sorted(os.listdir(dbg))[0].startswith("Ldet_00") and len(os.listdir(dbg)) >= 16
    n_default = len(k)
    # 2nd run: the file exists -> it is read; a higher threshold must reach the extractor and give fewer keypoints
    doc["detector_threshold"] = 0.01
    opt.write_text(json.dumps(doc))
    out2 = tmp_path / "f.json"
    assert cli.main(["extract_features", os.path.join(GOLD, "1.jpg"), str(out2), "--options", str(opt)], fake_api) == 0
    assert fake_api.seen["options"]["detector_threshold"] == 0.01
    k2, _ = formats.deserialize_features_from_file(str(out2))  # .json extension -> the serde_json layout
    assert 0 < len(k2) < n_default
    assert set(json.loads(out2.read_text())) == {"keypoints", "descriptors"}
    # a field is missing -> error like serde's
    opt.write_text(json.dumps({"num_sublevels": 4}))
    with pytest.raises(ValueError):
        cli.main(["extract_features", os.path.join(GOLD, "1.jpg"), str(out2), "-o", str(opt)], fake_api)


def test_extract_and_match_and_match_features_tools(tmp_path, fake_api):
    from akaze_rust_b200 import cli, formats
    prefix = str(tmp_path / "run")
    assert cli.main(["extract_and_match", os.path.join(GOLD, "1.jpg"), os.path.join(GOLD, "2.jpg"), prefix, "-m", str(tmp_path / "m.png")],
                    fake_api) == 0
    names = sorted(os.listdir(tmp_path))
    assert names == ["m.png", "run-extractions_0.cbor", "run-extractions_1.cbor", "run-matches.cbor"]  # extract_and_match.rs:66-71, -m
    from PIL import Image
    mi = Image.open(tmp_path / "m.png")
    assert mi.size == (2 * 2016, 1512) and mi.mode == "RGB"  # draw_matches: both images side by side (feature_match.rs:44-47)
    m = formats.deserialize_matches_from_file(prefix + "-matches.cbor")
    put = fake_api.seen["putative"]
    assert 0 < len(m) <= len(put)
    assert set(zip(m["index_0"], m["index_1"])).issubset(set(zip(put["index_0"], put["index_1"])))
    # match_features on the two extraction files gives matches from the same putative set
    out = str(tmp_path / "matches.json")
    assert cli.main(["match_features", prefix + "-extractions_0.cbor", prefix + "-extractions_1.cbor", out, "-t", "10"], fake_api) == 0
    m2 = formats.deserialize_matches_from_file(out)
    assert 0 < len(m2) <= len(put) and set(zip(m2["index_0"], m2["index_1"])).issubset(set(zip(put["index_0"], put["index_1"])))


def test_cli_rejects_missing_arguments():
    from akaze_rust_b200 import cli
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(["extract_features", "only_one_argument"])
