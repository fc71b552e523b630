"""On-disk formats of akaze-util (akaze-util/src/lib.rs:11-67): `Features` and `Vec<Match>` as bincode 1.1
(default options: little endian, u64 lengths, usize as u64) or as serde_json, chosen by the file extension
exactly like the reference (".json" -> JSON, anything else -> bincode).

Layouts: Keypoint = (f32, f32), f32, f32, u64, u64, f32 = 36 bytes;
Features = u64 n + n x 36 B + u64 n + n x (u64 len + bytes); Match = u64, u64, f64 = 24 bytes.
The same writer exists in C++ (include/akaze_b200.hpp, namespace akaze_util); tests check that the two
produce identical bytes. JSON numbers are the shortest strings that round-trip the f32/f64 value (what
serde_json's ryu backend prints for all but exponent formatting, which cannot be checked here without cargo).
"""
import json
import os

import numpy as np

KP_BIN = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("size", "<f4"), ("octave", "<u8"), ("class_id", "<u8"),
                   ("angle", "<f4")])
MATCH_BIN = np.dtype([("index_0", "<u8"), ("index_1", "<u8"), ("distance", "<f8")])
assert KP_BIN.itemsize == 36 and MATCH_BIN.itemsize == 24


def _is_json(path):
    return os.path.splitext(str(path))[1].lower() == ".json"


def _descriptor_rows(descriptors, n, desc_len):
    """Descriptor.vector of every keypoint as its own u8 array. The reference serialises each vector with its own
    length (akaze-util/src/lib.rs:11-14: Vec<Descriptor>), i.e. (162*channels+7)/8 = 21 / 41 / 61 bytes for
    descriptor_channels 1 / 2 / 3; desc_len=None takes the width of the array that is passed."""
    if n == 0:
        return []
    if isinstance(descriptors, np.ndarray) and descriptors.ndim == 2:
        d = np.asarray(descriptors, np.uint8)
        return list(d[:, :desc_len] if desc_len else d)
    rows = [np.asarray(r, np.uint8).ravel() for r in descriptors]
    return [r[:desc_len] for r in rows] if desc_len else rows


def features_to_bytes(keypoints, descriptors, desc_len=None):
    """keypoints: structured array with x, y, response, size, octave, class_id, angle; descriptors: (n, len) u8 array or a
    list of u8 arrays."""
    n = len(keypoints)
    k = np.zeros(n, KP_BIN)
    for f in KP_BIN.names:
        k[f] = keypoints[f]
    rows = _descriptor_rows(descriptors, n, desc_len)
    if len(rows) != n:
        raise ValueError("%d keypoints but %d descriptors" % (n, len(rows)))
    lens = {len(r) for r in rows}
    if len(lens) == 1:  # the usual case: one dense block
        ln = lens.pop()
        blk = np.zeros(n, np.dtype([("len", "<u8"), ("bytes", "u1", (ln,))]))
        blk["len"] = ln
        blk["bytes"] = np.stack(rows) if ln else np.zeros((n, 0), np.uint8)
        body = blk.tobytes()
    else:
        body = b"".join(np.uint64(len(r)).tobytes() + np.ascontiguousarray(r).tobytes() for r in rows)
    return np.uint64(n).tobytes() + k.tobytes() + np.uint64(n).tobytes() + body


def features_from_bytes(b):
    n = int(np.frombuffer(b, "<u8", 1, 0)[0])
    k = np.frombuffer(b, KP_BIN, n, 8).copy()
    at = 8 + 36 * n
    nd = int(np.frombuffer(b, "<u8", 1, at)[0])
    at += 8
    desc = []
    for _ in range(nd):
        ln = int(np.frombuffer(b, "<u8", 1, at)[0])
        at += 8
        desc.append(np.frombuffer(b, np.uint8, ln, at).copy())
        at += ln
    if at != len(b):
        raise ValueError("bincode: %d trailing bytes" % (len(b) - at))
    return k, desc


def matches_to_bytes(matches):
    m = np.zeros(len(matches), MATCH_BIN)
    for f in MATCH_BIN.names:
        m[f] = matches[f]
    return np.uint64(len(m)).tobytes() + m.tobytes()


def matches_from_bytes(b):
    n = int(np.frombuffer(b, "<u8", 1, 0)[0])
    if len(b) != 8 + 24 * n:
        raise ValueError("bincode: size mismatch")
    return np.frombuffer(b, MATCH_BIN, n, 8).copy()


def _f32(v):
    return json.loads(str(np.float32(v)))  # shortest decimal that round-trips the f32


def serialize_features_to_file(keypoints, descriptors, path, desc_len=None):
    """akaze-util serialize_features_to_file (lib.rs:17-30)."""
    if _is_json(path):
        descriptors = _descriptor_rows(descriptors, len(keypoints), desc_len)
        doc = {"keypoints": [{"point": [_f32(k["x"]), _f32(k["y"])], "response": _f32(k["response"]), "size": _f32(k["size"]),
                              "octave": int(k["octave"]), "class_id": int(k["class_id"]), "angle": _f32(k["angle"])} for k in keypoints],
               "descriptors": [{"vector": [int(v) for v in d]} for d in descriptors]}
        with open(path, "w") as fh:
            json.dump(doc, fh, separators=(",", ":"))
    else:
        with open(path, "wb") as fh:
            fh.write(features_to_bytes(keypoints, descriptors, desc_len))


def deserialize_features_from_file(path):
    """akaze-util deserialize_features_from_file (lib.rs:33-42) -> (keypoints KP_BIN array, list of u8 arrays)."""
    if _is_json(path):
        with open(path) as fh:
            doc = json.load(fh)
        k = np.zeros(len(doc["keypoints"]), KP_BIN)
        for i, e in enumerate(doc["keypoints"]):
            k[i] = (e["point"][0], e["point"][1], e["response"], e["size"], e["octave"], e["class_id"], e["angle"])
        return k, [np.asarray(d["vector"], np.uint8) for d in doc["descriptors"]]
    with open(path, "rb") as fh:
        return features_from_bytes(fh.read())


def serialize_matches_to_file(matches, path):
    """akaze-util serialize_matches_to_file (lib.rs:45-55)."""
    if _is_json(path):
        with open(path, "w") as fh:
            json.dump([{"index_0": int(m["index_0"]), "index_1": int(m["index_1"]), "distance": float(m["distance"])} for m in matches], fh,
                      separators=(",", ":"))
    else:
        with open(path, "wb") as fh:
            fh.write(matches_to_bytes(matches))


def deserialize_matches_from_file(path):
    """akaze-util deserialize_matches_from_file (lib.rs:58-67)."""
    if _is_json(path):
        with open(path) as fh:
            doc = json.load(fh)
        m = np.zeros(len(doc), MATCH_BIN)
        for i, e in enumerate(doc):
            m[i] = (e["index_0"], e["index_1"], e["distance"])
        return m
    with open(path, "rb") as fh:
        return matches_from_bytes(fh.read())
