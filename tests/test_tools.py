"""tools/render_depth (C++, over visma_b200/host/renderer_b200.h): the reference tool's argument, JSON keys and
output files (render/tools/render_depth.cpp:11-79, misc/render_depth.json, core/utils.h:359-373).  The binary is
built by `make -C tools` where Eigen headers exist (the dev container); it travels to the GPU box prebuilt."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "tools", "bin", "render_depth")
needs_bin = pytest.mark.skipif(not os.path.exists(BIN), reason="tools/bin/render_depth not built (`make -C tools`)")


def write_obj(path, V, F):
    with open(path, "w") as f:
        f.write("# chair fixture\n")
        for v in V:
            f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in v))
        for t in F:
            f.write("f %d//%d %d//%d %d//%d\n" % (t[0] + 1, t[0] + 1, t[1] + 1, t[1] + 1, t[2] + 1, t[2] + 1))


def run_tool(cfg_path, cwd):
    return subprocess.run([BIN, str(cfg_path)], cwd=str(cwd), capture_output=True, text=True, timeout=300)


@needs_bin
def test_usage_and_config_errors(tmp_path):
    """Host-side behaviour that needs no GPU: usage, unreadable / malformed config, missing mesh."""
    out = subprocess.run([BIN], capture_output=True, text=True)
    assert out.returncode == 2 and "usage: render_depth" in out.stderr
    out = run_tool(tmp_path / "nope.json", tmp_path)
    assert out.returncode == 1 and "failed to read file" in out.stderr
    (tmp_path / "bad.json").write_text('{"fx": 400, "translation": [0, 0, 1}')
    out = run_tool(tmp_path / "bad.json", tmp_path)
    assert out.returncode == 1 and "Line 1" in out.stderr   # jsoncpp's own diagnostics (the reference's parser)
    (tmp_path / "nomesh.json").write_text('{ // jsoncpp-style comment\n "mesh": "absent.obj", /* block */ "translation": [0, 0, 1]}')
    out = run_tool(tmp_path / "nomesh.json", tmp_path)
    assert out.returncode == 1 and "failed to load mesh absent.obj" in out.stderr


@needs_bin
@pytest.mark.gpu
def test_render_depth_tool_outputs(vb, oracle, tmp_path):
    """misc/render_depth.json's keys (with comments, as jsoncpp accepts): depthmap.bin and mask.bin equal the
    Python mirror's maps and the oracle's, bit for bit — including the tool's cy := fy quirk (:31)."""
    from visma_b200 import io2d
    V, F = vb.synth.load_chair()
    write_obj(tmp_path / "chair.obj", V, F)
    cfg = {"major_version": 3, "minor_version": 3, "fx": 400, "fy": 400, "z_far": 10, "mesh": "chair.obj",
           "translation": [0, 0, 1], "show": False, "save": True, "output_path": ".", "mask": True,
           "cy": 240}  # cy is read and ignored by the tool
    (tmp_path / "cfg.json").write_text("// render_depth config\n" + json.dumps(cfg, indent=1))
    out = run_tool(tmp_path / "cfg.json", tmp_path)
    assert out.returncode == 0, out.stderr
    assert "2492 vertices, 4999 faces" in out.stdout
    depth = io2d.LoadMat(tmp_path / "depthmap.bin", np.float32)
    mask = io2d.LoadMat(tmp_path / "mask.bin", np.uint8)
    assert depth.shape == (480, 640) and mask.shape == (480, 640)
    # the OBJ text round-trips the float32 vertices exactly (%.9g), so the maps must match bit for bit
    ren = vb.ren.Renderer(480, 640, 3, 3)
    ren.SetCamera(0.05, 10.0, 400.0, 400.0, 320.0, 400.0)
    ren.SetCamera(np.eye(4, dtype=np.float32))
    ren.SetMesh(V, F)
    model = vb.synth.make_T(np.eye(3), [0, 0, 1.0])
    assert (depth == ren.RenderDepth(model)).all()
    assert (mask == ren.RenderMask(model)).all()
    P = oracle.projection(0.05, 10.0, 400.0, 400.0, 320.0, 400.0, 480, 640)
    oz, od = oracle.render_depth(V, F, np.asarray(model, np.float32).T.reshape(-1),
                                 oracle.view(np.eye(4, dtype=np.float32).reshape(-1)), P, 480, 640)
    assert (depth == od).all()
    # reference polarity: background 255 (the GL clear colour, render/renderer.cpp:411-422), covered pixels 0
    assert ((mask == 0) == (oz < oracle.ZMAX24)).all() and (mask == 0).sum() > 10000
    # saving disabled: nothing written
    (tmp_path / "o2").mkdir()
    cfg2 = dict(cfg, save=False, mesh=str(tmp_path / "chair.obj"))
    (tmp_path / "o2" / "cfg.json").write_text(json.dumps(cfg2))
    out = run_tool(tmp_path / "o2" / "cfg.json", tmp_path / "o2")
    assert out.returncode == 0 and not (tmp_path / "o2" / "depthmap.bin").exists()
