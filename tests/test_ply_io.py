"""PLY files in the reference's layout (scene/gaussian_model.py:286-325, 357-398): header text, property order, the
channel-major SH packing, and a bit-exact round trip.  CPU only."""
import os
import tempfile

import numpy as np

import helpers  # noqa: F401  (sys.path)
from b200gs import ply_io


def _scene(P=37, seed=3):
    r = np.random.default_rng(seed)
    return dict(xyz=r.normal(size=(P, 3)).astype(np.float32), shs=r.normal(size=(P, 16, 3)).astype(np.float32),
                opacity=r.normal(size=(P, 1)).astype(np.float32), scaling=r.normal(size=(P, 3)).astype(np.float32),
                rotation=r.normal(size=(P, 4)).astype(np.float32), feature=r.normal(size=(P, 3)).astype(np.float32))


def test_header_and_layout_follow_the_reference():
    s = _scene()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "point_cloud", "iteration_7", "point_cloud.ply")
        ply_io.save_ply(path, **s)
        raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    props = [l.split()[-1] for l in lines[3:]]
    assert all(l.startswith("property float ") for l in lines[3:])
    assert props == (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)]
                     + ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3",
                        "languagefeature_0", "languagefeature_1", "languagefeature_2"])
    t = np.frombuffer(body, "<f4").reshape(37, len(props))
    np.testing.assert_array_equal(t[:, 0:3], s["xyz"])
    assert not t[:, 3:6].any()  # normals are zeros
    np.testing.assert_array_equal(t[:, 6:9], s["shs"][:, 0, :])
    # f_rest is channel-major: f_rest_{c*15+k} = shs[:, 1+k, c]   (features.transpose(1, 2).flatten(start_dim=1))
    np.testing.assert_array_equal(t[:, 9 + 1 * 15 + 4], s["shs"][:, 1 + 4, 1])
    np.testing.assert_array_equal(t[:, 54], s["opacity"][:, 0])


def test_round_trip_is_bit_exact_with_and_without_the_feature_head():
    for with_feature in (True, False):
        s = _scene(P=101, seed=9)
        if not with_feature:
            s["feature"] = None
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "pc.ply")
            ply_io.save_ply(path, **s)
            back = ply_io.load_ply(path)
        for k, v in s.items():
            if v is None:
                assert back[k] is None
            else:
                np.testing.assert_array_equal(back[k], v)
