"""The reference's on-disk format for 2-D maps — feh::SaveMat<T> (core/utils.h:359-373): int32 rows, int32 cols,
then rows*cols values of T, row-major — as read by misc/show_2Dmap.py:16-21.  depthmap.bin holds float32 window
depth, mask.bin uint8 (render/tools/render_depth.cpp:63-77)."""
import numpy as np


def SaveMat(filename, mat):
    mat = np.ascontiguousarray(mat)
    if mat.ndim != 2:
        raise ValueError("SaveMat writes 2-D maps")
    with open(filename, "wb") as f:
        np.array(mat.shape, np.int32).tofile(f)
        mat.tofile(f)


def LoadMat(filename, dtype):
    with open(filename, "rb") as f:
        h, w = np.frombuffer(f.read(8), dtype=np.int32)
        return np.frombuffer(f.read(), dtype=dtype).reshape(int(h), int(w)).copy()
