"""Minimal `plyfile` stand-in (test infrastructure): binary little-endian vertex tables only, which is all the reference
reads and writes (scene/dataset_readers.py:156-190 fetchPly / storePly, scene/gaussian_model.py:380-397, 515-559)."""
import numpy as np


class _Property:
    def __init__(self, name, dtype):
        self.name, self.val_dtype = name, dtype


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))

    @property
    def properties(self):
        return tuple(_Property(n, self.data.dtype[n]) for n in self.data.dtype.names)

    @property
    def count(self):
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]


_PLY_TYPES = {"f4": "float", "f8": "double", "u1": "uchar", "i1": "char", "u2": "ushort", "i2": "short", "u4": "uint", "i4": "int"}
_NP_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
             "ushort": "u2", "uint16": "u2", "short": "i2", "int16": "i2", "uint": "u4", "uint32": "u4", "int": "i4", "int32": "i4"}


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def write(self, path):
        with open(path, "wb") as f:
            head = ["ply", "format binary_little_endian 1.0"]
            for e in self.elements:
                head.append(f"element {e.name} {len(e.data)}")
                for n in e.data.dtype.names:
                    t = e.data.dtype[n]
                    head.append(f"property {_PLY_TYPES[t.str[1:]]} {n}")
            head.append("end_header")
            f.write(("\n".join(head) + "\n").encode("ascii"))
            for e in self.elements:
                le = e.data.astype(e.data.dtype.newbyteorder("<"), copy=False)
                f.write(np.ascontiguousarray(le).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            if f.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt, elements, cur = None, [], None
            while True:
                line = f.readline()
                if not line:
                    raise ValueError("unterminated PLY header")
                tok = line.decode("ascii").split()
                if not tok or tok[0] == "comment":
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "element":
                    cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                    elements.append(cur)
                elif tok[0] == "property":
                    if tok[1] == "list":
                        raise ValueError("list properties are not supported by this stand-in")
                    cur["props"].append((tok[2], _NP_TYPES[tok[1]]))
                elif tok[0] == "end_header":
                    break
            if fmt not in ("binary_little_endian", "binary_big_endian"):
                raise ValueError("only binary PLY is supported by this stand-in")
            order = "<" if fmt == "binary_little_endian" else ">"
            out = []
            for e in elements:
                dt = np.dtype([(n, order + t) for n, t in e["props"]])
                raw = f.read(dt.itemsize * e["count"])
                out.append(PlyElement(e["name"], np.frombuffer(raw, dtype=dt, count=e["count"]).copy()))
            return PlyData(out)
