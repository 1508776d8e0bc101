#!/usr/bin/env python
"""render_depth <cfg.json> — the reference tool (render/tools/render_depth.cpp:11-79) on the CUDA rasteriser: same
JSON keys and defaults (misc/render_depth.json), same call sequence — including the tool's quirk of passing fy as
cy (:31) — and the same output files: depthmap.bin (+ mask.bin) in feh::SaveMat's format.  Dev/demo tool: the
mesh is read from the cfg's path when it exists (OBJ `v`/`f` lines), else the packaged chair fixture is used."""
import json
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from visma_b200 import io2d, renderer, synth  # noqa: E402


def load_json_with_comments(path):
    return json.loads(re.sub(r"//[^\n]*", "", open(path).read()))  # jsoncpp accepts // comments


def load_obj(path):
    V, F = [], []
    for line in open(path):
        t = line.split()
        if t[:1] == ["v"]:
            V.append([float(x) for x in t[1:4]])       # LoadMesh keeps the first 3 columns (core/utils.cpp:132)
        elif t[:1] == ["f"]:
            F.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.asarray(V, np.float32), np.asarray(F, np.int32)


def main(argv):
    cfg = load_json_with_comments(argv[1])
    H, W = int(cfg.get("image_height", 480)), int(cfg.get("image_width", 640))
    zn, zf = float(cfg.get("z_near", 0.05)), float(cfg.get("z_far", 10.0))
    fx, fy = float(cfg.get("fx", 400)), float(cfg.get("fy", 400))
    cx, cy = float(cfg.get("cx", 320)), float(cfg.get("cy", 240))  # cy is read and then ignored, as in the tool
    ren = renderer.Renderer(H, W, int(cfg.get("major_version", 4)), int(cfg.get("minor_version", 3)))
    ren.SetCamera(zn, zf, fx, fy, cx, fy)                       # render_depth.cpp:31 passes fy where cy belongs
    ren.SetCamera(np.eye(4, dtype=np.float32))                  # :32
    mesh = cfg.get("mesh", "misc/hermanmiller_aeron.obj")
    V, F = load_obj(mesh) if os.path.exists(mesh) else synth.load_chair()
    ren.SetMesh(V, F)
    model = np.eye(4, dtype=np.float32)
    model[:3, 3] = cfg.get("translation", [0, 0, 0])            # SE3f(SO3f(), translation), :44
    depth = ren.RenderDepth(model)
    if cfg.get("save", False):
        out = cfg.get("output_path", ".")
        io2d.SaveMat(os.path.join(out, "depthmap.bin"), depth)
        if cfg.get("mask", False):
            io2d.SaveMat(os.path.join(out, "mask.bin"), ren.RenderMask(model))
    print("rendered %dx%d, %d covered pixels" % (H, W, int((depth < 1).sum())))


if __name__ == "__main__":
    main(sys.argv)
