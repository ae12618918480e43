"""Binary PLY reader / writer for Gaussian models, without the `plyfile` dependency.

Mirrors the on-disk layout of ``GaussianModel.save_ply`` / ``load_ply`` (reference scene/gaussian_model.py:363-397,
515-559): one ``vertex`` element of float32 properties

    x y z  nx ny nz  f_dc_0..2  f_rest_0..(3(M-1)-1)  opacity  scale_0..2  rot_0..3  filter_3D

where the SH blocks are stored channel-major (``features.transpose(1, 2).flatten(1)``).  The loader accepts any
property order, ``float`` / ``double`` / integer property types, both byte orders and ASCII bodies, and returns the
arrays in the layouts the model keeps in memory (features as [P, coeffs, 3]).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def attribute_names(n_rest_coeffs: int, with_filter: bool = True):
    """``construct_list_of_attributes`` (scene/gaussian_model.py:363-377) for a model with 1 + n_rest_coeffs SH coefficients."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(3 * n_rest_coeffs)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(3)]
    names += [f"rot_{i}" for i in range(4)]
    if with_filter:
        names.append("filter_3D")
    return names


def save_gaussian_ply(path: str, xyz, features_dc, features_rest, opacity, scaling, rotation, filter_3D=None) -> None:
    """Arrays (numpy or torch, any device) in the model's layouts: features_dc [P,1,3], features_rest [P,M-1,3], opacity
    [P,1], scaling [P,3], rotation [P,4], filter_3D [P,1] or None (then the column is left out)."""
    def arr(a):
        if hasattr(a, "detach"):
            a = a.detach().cpu().numpy()
        return np.asarray(a, dtype=np.float32)

    xyz, features_dc, features_rest = arr(xyz), arr(features_dc), arr(features_rest)
    P = xyz.shape[0]
    cols = [xyz, np.zeros_like(xyz),
            features_dc.transpose(0, 2, 1).reshape(P, -1), features_rest.transpose(0, 2, 1).reshape(P, -1),
            arr(opacity).reshape(P, 1), arr(scaling).reshape(P, 3), arr(rotation).reshape(P, 4)]
    if filter_3D is not None:
        cols.append(arr(filter_3D).reshape(P, 1))
    table = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    names = attribute_names(features_rest.shape[1], with_filter=filter_3D is not None)
    assert table.shape[1] == len(names)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {P}"] + [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(table.tobytes())


def read_ply_vertices(path: str) -> Dict[str, np.ndarray]:
    """All scalar properties of the ``vertex`` element as a dict name -> 1-D array."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, current = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                current = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(current)
            elif tok[0] == "property":
                if tok[1] == "list":
                    if current["name"] == "vertex" or not elements or elements[0]["name"] != "vertex" or current is elements[0]:
                        raise ValueError(f"{path}: list properties are only tolerated after the vertex element")
                    current["props"].append(("__list__", None))
                else:
                    if tok[1] not in _PLY_TYPES:
                        raise ValueError(f"{path}: unknown PLY type {tok[1]}")
                    current["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        if not elements or elements[0]["name"] != "vertex":
            raise ValueError(f"{path}: the first element must be `vertex`")
        v = elements[0]
        if fmt == "ascii":
            data = np.loadtxt(f, dtype=np.float64, max_rows=v["count"], ndmin=2)
            if data.shape != (v["count"], len(v["props"])):
                raise ValueError(f"{path}: vertex table has shape {data.shape}")
            return {n: data[:, i].astype(t) for i, (n, t) in enumerate(v["props"])}
        order = "<" if fmt == "binary_little_endian" else ">"
        dtype = np.dtype([(n, order + t) for n, t in v["props"]])
        raw = f.read(dtype.itemsize * v["count"])
        if len(raw) != dtype.itemsize * v["count"]:
            raise ValueError(f"{path}: truncated vertex data")
        rec = np.frombuffer(raw, dtype=dtype, count=v["count"])
        return {n: np.ascontiguousarray(rec[n]) for n, _ in v["props"]}


def _numbered(props: Dict[str, np.ndarray], prefix: str):
    names = sorted((n for n in props if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    return names


def load_gaussian_ply(path: str, max_sh_degree: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Returns float32 arrays ``xyz`` [P,3], ``features_dc`` [P,1,3], ``features_rest`` [P,M-1,3], ``opacity`` [P,1],
    ``scaling`` [P,3], ``rotation`` [P,4], ``filter_3D`` [P,1] (zeros when the file has no such column, i.e. a plain
    3DGS model) and ``sh_degree``.  With ``max_sh_degree`` given the SH count is checked as the reference does."""
    p = read_ply_vertices(path)
    for need in ("x", "y", "z", "opacity", "f_dc_0", "f_dc_1", "f_dc_2"):
        if need not in p:
            raise ValueError(f"{path}: missing property {need}")
    f32 = lambda a: np.asarray(a, dtype=np.float32)  # noqa: E731
    xyz = np.stack([f32(p["x"]), f32(p["y"]), f32(p["z"])], axis=1)
    P = xyz.shape[0]
    dc = np.stack([f32(p["f_dc_0"]), f32(p["f_dc_1"]), f32(p["f_dc_2"])], axis=1).reshape(P, 3, 1)
    rest_names = _numbered(p, "f_rest_")
    if len(rest_names) % 3 != 0:
        raise ValueError(f"{path}: {len(rest_names)} f_rest_* columns is not a multiple of 3")
    n_rest = len(rest_names) // 3
    deg = int(round((n_rest + 1) ** 0.5)) - 1
    if (deg + 1) ** 2 != n_rest + 1:
        raise ValueError(f"{path}: {n_rest + 1} SH coefficients is not a square number")
    if max_sh_degree is not None and len(rest_names) != 3 * (max_sh_degree + 1) ** 2 - 3:
        raise ValueError(f"{path}: expected SH degree {max_sh_degree}, file holds degree {deg}")
    rest = np.stack([f32(p[n]) for n in rest_names], axis=1).reshape(P, 3, n_rest) if n_rest else np.zeros((P, 3, 0), np.float32)
    scale_names, rot_names = _numbered(p, "scale_"), _numbered(p, "rot")
    if len(scale_names) != 3 or len(rot_names) != 4:
        raise ValueError(f"{path}: need scale_0..2 and rot_0..3")
    return {
        "xyz": xyz,
        "features_dc": np.ascontiguousarray(dc.transpose(0, 2, 1)),
        "features_rest": np.ascontiguousarray(rest.transpose(0, 2, 1)),
        "opacity": f32(p["opacity"]).reshape(P, 1),
        "scaling": np.stack([f32(p[n]) for n in scale_names], axis=1),
        "rotation": np.stack([f32(p[n]) for n in rot_names], axis=1),
        "filter_3D": f32(p["filter_3D"]).reshape(P, 1) if "filter_3D" in p else np.zeros((P, 1), np.float32),
        "sh_degree": deg,
    }
