"""Converts the reference's only mesh fixture, misc/hermanmiller_aeron.obj (2492 v with per-vertex rgb,
4999 `v//vn` faces), into visma_b200/data/chair_mesh.npz (float32 V, int32 F) so bench.py and the GPU tests
can use it on the GPU box, where /root/reference does not exist.  Mirrors feh::LoadMesh's
`V.leftCols(3)` (core/utils.cpp:125-135).  Run once in the dev container:  python scripts/make_mesh_fixture.py
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
V, F = [], []
with open(os.path.join(REF, "misc", "hermanmiller_aeron.obj")) as f:
    for line in f:
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            V.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            F.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
V = np.asarray(V, np.float32)
F = np.asarray(F, np.int32)
assert V.shape == (2492, 3) and F.shape == (4999, 3), (V.shape, F.shape)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "visma_b200", "data", "chair_mesh.npz")
np.savez_compressed(out, V=V, F=F)
print("wrote", out, V.min(0), V.max(0))
