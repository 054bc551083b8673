"""Small stand-ins for third-party packages the reference imports but this image does not ship
(munch, plyfile, natsort).  Only what the tracking path needs."""
import re

import numpy as np


class Munch(dict):
    """dict with attribute access (what the reference uses munch.munchify for: main.py:24-25)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def munchify(x):
    if isinstance(x, dict):
        return Munch({k: munchify(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return type(x)(munchify(v) for v in x)
    return x


def natsorted(seq):
    key = lambda s: [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", str(s))]
    return sorted(seq, key=key)


_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


def read_ply_vertices(path):
    """Vertex element of a binary-little-endian or ascii PLY as a numpy structured array."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, props, count, in_vertex = None, [], 0, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list properties are not supported in the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "binary_little_endian":
            dt = np.dtype([(n, "<" + t) for n, t in props])
            return np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=count, ndmin=2)
            out = np.empty(count, dtype=[(n, t) for n, t in props])
            for i, (n, _) in enumerate(props):
                out[n] = data[:, i]
            return out
        raise ValueError(f"{path}: unsupported PLY format {fmt}")


def write_ply_vertices(path, arr):
    """Writes a structured array as the vertex element of a binary-little-endian PLY."""
    names = {"f4": "float", "f8": "double", "u1": "uchar", "i4": "int", "u4": "uint", "i2": "short", "u2": "ushort", "i1": "char"}
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {arr.shape[0]}\n".encode())
        for n in arr.dtype.names:
            f.write(f"property {names[arr.dtype[n].str[1:]]} {n}\n".encode())
        f.write(b"end_header\n")
        f.write(np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes())
