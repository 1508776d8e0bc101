"""Minimal reader for the 'DATA binary' float32 PCD files Open3D ships as ICP test data
(FIELDS x y z rgb normal_x normal_y normal_z curvature).  Used only by the fixture
generator scripts; mirrors Open3D's reader behaviour of dropping non-finite points
(thirdparty/Open3D/src/IO/FileFormat/FilePCD.cpp:757)."""
import numpy as np


def read_pcd(path):
    with open(path, "rb") as f:
        fields, count, npts = None, None, None
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("FIELDS"):
                fields = line.split()[1:]
            elif line.startswith("POINTS"):
                npts = int(line.split()[1])
            elif line.startswith("DATA"):
                assert line.split()[1] == "binary", line
                break
        raw = np.frombuffer(f.read(npts * 4 * len(fields)), dtype=np.float32)
    raw = raw.reshape(npts, len(fields))
    col = {n: i for i, n in enumerate(fields)}
    xyz = raw[:, [col["x"], col["y"], col["z"]]]
    nrm = raw[:, [col["normal_x"], col["normal_y"], col["normal_z"]]]
    ok = np.isfinite(xyz).all(1)
    return xyz[ok].copy(), nrm[ok].copy()
