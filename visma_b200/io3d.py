"""The file formats VISMA's two callers of the ICP path read and write (host-side glue, numpy only):
PLY point clouds (Open3D's ReadPointCloudFromPLY / WritePointCloud: test.klg.ply, fragments/*.ply), OBJ meshes
(igl::readOBJ: the CAD database), and JSON with // comments (cfg/tool.json is parsed by jsoncpp / folly's
stripComments: src/annotation.cpp:99-100, core/utils.h LoadJson).  Matrices in JSON are flat row-major lists
(WriteMatrixToJson / GetMatrixFromJson, core/utils.h:300-339)."""
import json
import re

import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4",
              "float": "f4", "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2",
              "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def read_ply(path):
    """-> (points N x 3 float64, normals N x 3 float64 or None).  ascii and binary_little_endian, scalar vertex
    properties of any type; other elements are ignored."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply", "not a PLY file"
        fmt, n_vertex, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line == "end_header":
                break
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n_vertex = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list property on vertices")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
        if fmt == "ascii":
            a = np.loadtxt(f, max_rows=n_vertex, ndmin=2)
            cols = {name: a[:, i] for i, (name, _) in enumerate(props)}
        elif fmt == "binary_little_endian":
            dt = np.dtype([(name, "<" + t) for name, t in props])
            rec = np.frombuffer(f.read(dt.itemsize * n_vertex), dtype=dt, count=n_vertex)
            cols = {name: rec[name] for name, _ in props}
        else:
            raise ValueError("PLY format %r not supported" % fmt)
    pts = np.stack([cols["x"], cols["y"], cols["z"]], 1).astype(np.float64)
    nrm = np.stack([cols["nx"], cols["ny"], cols["nz"]], 1).astype(np.float64) if "nx" in cols else None
    return np.ascontiguousarray(pts), None if nrm is None else np.ascontiguousarray(nrm)


def write_ply(path, points, normals=None):
    """binary_little_endian, double x y z (+ nx ny nz): what Open3D's WritePointCloud(..., write_ascii=false) emits."""
    pts = np.asarray(points, np.float64)
    cols = [pts] if normals is None else [pts, np.asarray(normals, np.float64)]
    names = ["x", "y", "z"] + ([] if normals is None else ["nx", "ny", "nz"])
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(pts)).encode())
        for n in names:
            f.write(("property double %s\n" % n).encode())
        f.write(b"end_header\n")
        np.ascontiguousarray(np.concatenate(cols, 1), "<f8").tofile(f)


def read_obj(path):
    """-> (V n x 3 float64, F m x 3 int32): `v x y z [r g b]` and `f a[/..] b[/..] c[/..] ...` (polygons are fanned)."""
    V, F = [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                V.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] == "f":
                idx = [int(x.split("/")[0]) for x in t[1:]]
                idx = [i - 1 if i > 0 else len(V) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    F.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(V, np.float64).reshape(-1, 3), np.asarray(F, np.int32).reshape(-1, 3)


def write_obj(path, V, F):
    with open(path, "w") as f:
        for v in np.asarray(V, np.float64):
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        for t in np.asarray(F) + 1:
            f.write("f %d %d %d\n" % tuple(t))


def load_json(path):
    """JSON with // and /* */ comments, as jsoncpp's reader and folly::json::stripComments accept."""
    s = open(path).read()
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    s = re.sub(r'("(?:\\.|[^"\\])*")|//[^\n]*', lambda m: m.group(1) or "", s)
    return json.loads(s)


def save_json(obj, path):
    with open(path, "w") as f:
        json.dump(obj, f, indent=2)


def matrix_to_json(m):
    """WriteMatrixToJson (core/utils.h:333-339): flat, row-major."""
    return [float(x) for x in np.asarray(m, np.float64).reshape(-1)]


def matrix_from_json(v, rows, cols):
    return np.asarray(v, np.float64).reshape(rows, cols)
