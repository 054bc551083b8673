"""TEST INFRASTRUCTURE — minimal stand-in for `plyfile` (binary little-endian vertex element only),
enough for the reference's GaussianModel.load_ply / save_ply (gaussian_model.py:173-327)."""
import numpy as np

_T = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1",
      "short": "i2", "ushort": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}
_N = {v: k for k, v in {"float": "f4", "double": "f8", "uchar": "u1", "char": "i1", "short": "i2", "ushort": "u2",
                        "int": "i4", "uint": "u4"}.items()}


class PlyProperty:
    def __init__(self, name, dtype):
        self.name, self.dtype = name, dtype


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data
        self.properties = [PlyProperty(n, data.dtype[n].str[1:]) for n in data.dtype.names]

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))

    def __getitem__(self, k):
        return self.data[k]

    def __len__(self):
        return self.data.shape[0]


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"ply"
            fmt, elems, cur = None, [], None
            while True:
                tok = f.readline().decode("ascii").split()
                if not tok:
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "element":
                    cur = [tok[1], int(tok[2]), []]
                    elems.append(cur)
                elif tok[0] == "property":
                    cur[2].append((tok[2], "<" + _T[tok[1]]))
                elif tok[0] == "end_header":
                    break
            assert fmt == "binary_little_endian", fmt
            out = []
            for name, count, props in elems:
                dt = np.dtype(props)
                out.append(PlyElement(name, np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)))
            return PlyData(out)

    def write(self, path):
        with open(path, "wb") as f:
            f.write(b"ply\nformat binary_little_endian 1.0\n")
            for e in self.elements:
                f.write(f"element {e.name} {len(e)}\n".encode())
                for n in e.data.dtype.names:
                    f.write(f"property {_N[e.data.dtype[n].str[1:]]} {n}\n".encode())
            f.write(b"end_header\n")
            for e in self.elements:
                f.write(np.ascontiguousarray(e.data).tobytes())
