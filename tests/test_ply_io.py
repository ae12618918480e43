"""PLY reader / writer against the layout of GaussianModel.save_ply / load_ply (scene/gaussian_model.py:363-397,515-559)."""
import os
import struct
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
from rade_gs_b200 import ply_io  # noqa: E402


def _model(P, deg, seed=0):
    r = np.random.default_rng(seed)
    M = (deg + 1) ** 2
    return dict(xyz=r.normal(size=(P, 3)), features_dc=r.normal(size=(P, 1, 3)), features_rest=r.normal(size=(P, M - 1, 3)),
                opacity=r.normal(size=(P, 1)), scaling=r.normal(size=(P, 3)), rotation=r.normal(size=(P, 4)), filter_3D=r.random(size=(P, 1)))


@pytest.mark.parametrize("deg", [0, 1, 3])
def test_round_trip_and_reference_column_layout(tmp_path, deg):
    m = {k: v.astype(np.float32) for k, v in _model(37, deg).items()}
    path = str(tmp_path / "sub" / "point_cloud.ply")
    ply_io.save_gaussian_ply(path, **m)
    back = ply_io.load_gaussian_ply(path, max_sh_degree=deg)
    assert back["sh_degree"] == deg
    for k, v in m.items():
        assert back[k].dtype == np.float32 and np.array_equal(back[k], v), k
    # header and body exactly as the reference writes them: float32 columns in construct_list_of_attributes order,
    # SH blocks channel-major (features.transpose(1, 2).flatten(1))
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    names = [l.split()[2] for l in lines[3:]]
    assert names == ply_io.attribute_names((deg + 1) ** 2 - 1)
    assert all(l.startswith("property float ") for l in lines[3:])
    table = np.frombuffer(body, dtype="<f4").reshape(37, len(names))
    assert np.array_equal(table[:, 0:3], m["xyz"]) and not table[:, 3:6].any()
    assert np.array_equal(table[:, 6:9], m["features_dc"][:, 0, :])
    n_rest = (deg + 1) ** 2 - 1
    if n_rest:
        # column f_rest_{c * n_rest + k} holds coefficient k + 1 of colour channel c
        assert np.array_equal(table[:, 9 + 1 * n_rest + 2], m["features_rest"][:, 2, 1])
    assert np.array_equal(table[:, names.index("opacity")], m["opacity"][:, 0])
    assert np.array_equal(table[:, names.index("rot_3")], m["rotation"][:, 3])
    assert np.array_equal(table[:, -1], m["filter_3D"][:, 0])


def test_plain_3dgs_file_without_filter_column_and_wrong_degree(tmp_path):
    m = {k: v.astype(np.float32) for k, v in _model(5, 1).items()}
    m["filter_3D"] = None
    path = str(tmp_path / "a.ply")
    ply_io.save_gaussian_ply(path, **m)
    back = ply_io.load_gaussian_ply(path)
    assert not back["filter_3D"].any() and back["filter_3D"].shape == (5, 1)
    with pytest.raises(ValueError):
        ply_io.load_gaussian_ply(path, max_sh_degree=3)  # the reference asserts on the f_rest count (:532)


def test_reader_handles_other_orders_types_and_encodings(tmp_path):
    # big-endian doubles in a shuffled property order, with a comment line and a trailing face element
    P = 4
    r = np.random.default_rng(1)
    cols = {n: r.normal(size=P) for n in ["z", "x", "y", "opacity", "f_dc_1", "f_dc_0", "f_dc_2", "scale_2", "scale_0", "scale_1",
                                          "rot_0", "rot_1", "rot_2", "rot_3", "filter_3D"]}
    path = str(tmp_path / "be.ply")
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_big_endian 1.0\ncomment made by hand\nelement vertex %d\n" % P).encode())
        for n in cols:
            f.write(("property double %s\n" % n).encode())
        f.write(b"element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        for i in range(P):
            f.write(struct.pack(">%dd" % len(cols), *[cols[n][i] for n in cols]))
    back = ply_io.load_gaussian_ply(path)
    assert back["sh_degree"] == 0 and back["features_rest"].shape == (P, 0, 3)
    assert np.allclose(back["xyz"], np.stack([cols["x"], cols["y"], cols["z"]], 1).astype(np.float32))
    assert np.allclose(back["scaling"][:, 2], cols["scale_2"].astype(np.float32))
    assert np.allclose(back["features_dc"][:, 0, 1], cols["f_dc_1"].astype(np.float32))
    # ascii body
    path2 = str(tmp_path / "ascii.ply")
    with open(path2, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 2\n" + "".join(f"property float {n}\n" for n in cols) + "end_header\n")
        for i in range(2):
            f.write(" ".join(repr(float(cols[n][i])) for n in cols) + "\n")
    back2 = ply_io.load_gaussian_ply(path2)
    assert np.allclose(back2["opacity"][:, 0], cols["opacity"][:2].astype(np.float32))


def test_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.ply"
    p.write_bytes(b"plx\n")
    with pytest.raises(ValueError):
        ply_io.read_ply_vertices(str(p))
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nend_header\n\x00\x00")
    with pytest.raises(ValueError):
        ply_io.read_ply_vertices(str(p))
