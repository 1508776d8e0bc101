"""GPU parity: depth rasteriser (bit-exact integer z-buffer) and voxel down-sample (bit-exact averages)
against the oracle.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CAM = dict(zn=0.05, zf=10.0, fx=400.0, fy=400.0, cx=320.0, cy=240.0, H=480, W=640)


def oracle_render(oracle, V, F, model, pose=np.eye(4), cam=CAM):
    P = oracle.projection(cam["zn"], cam["zf"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["H"], cam["W"])
    Vw = oracle.view(np.asarray(pose, np.float32).T.reshape(-1))
    return oracle.render_depth(V, F, np.asarray(model, np.float32).T.reshape(-1), Vw, P, cam["H"], cam["W"])


def gpu_renderer(vb, cam=CAM, pose=np.eye(4)):
    r = vb.ren.Renderer(cam["H"], cam["W"])
    r.SetCamera(cam["zn"], cam["zf"], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    r.SetCamera(pose)
    return r


def test_render_depth_tool_config(vb, oracle):
    """misc/render_depth.json through render_depth.cpp: chair at t=[0,0,1], and the tool's cy:=fy quirk
    (render/tools/render_depth.cpp:31)."""
    V, F = vb.synth.load_chair()
    model = vb.synth.make_T(np.eye(3), [0, 0, 1.0])
    for cy in (240.0, 400.0):
        cam = dict(CAM, cy=cy)
        r = gpu_renderer(vb, cam)
        r.SetMesh(V, F)
        depth, z24 = r.RenderDepthBatch([model], want_z24=True)
        oz, od = oracle_render(oracle, V, F, model, cam=cam)
        assert (z24[0] == oz).all()
        assert (depth[0] == od).all()
        covered = (oz < oracle.ZMAX24).sum()
        assert covered > 20000
        # metric depth of the covered pixels is around 1 m
        lin = vb.ren.LinearizeDepth(depth[0][oz < oracle.ZMAX24].astype(np.float64), cam["zn"], cam["zf"])
        assert 0.6 < lin.min() and lin.max() < 1.5


def test_render_depth_against_an_independent_pinhole_rasteriser__parity_unpinned_by_the_reference(vb):
    """R5 rests on more than the restated GL pipeline: the GPU z-buffer against the float64 rasteriser derived from
    the closed-form pinhole mapping (tests/indep_raster.py) — identical coverage away from triangle edges, depth
    within a few 24-bit units (plus the 1/256-pixel snapping on steep triangles)."""
    from indep_raster import compare, render_depth_indep
    V, F = vb.synth.load_chair()
    pose = vb.synth.make_T(vb.synth.rot_xyz(0.05, -0.1, 0.02), [0.03, 0.02, 0.1])
    for cam, model, ps in ((dict(CAM, cy=400.0), vb.synth.make_T(np.eye(3), [0, 0, 1.0]), np.eye(4)),
                           (CAM, vb.synth.make_T(vb.synth.rot_y(0.8), [0.2, -0.1, 1.6]), pose)):
        r = gpu_renderer(vb, cam, ps)
        r.SetMesh(V, F)
        _, z24 = r.RenderDepthBatch([model], want_z24=True)
        depth, covered, uncertain, slope = render_depth_indep(V, F, model, ps, cam["zn"], cam["zf"], cam["fx"], cam["fy"],
                                                              cam["cx"], cam["cy"], cam["H"], cam["W"])
        c = compare(z24[0], depth, covered, uncertain, slope)
        assert c["n_covered"] > 5000 and c["coverage_mismatch_sure"] == 0 and c["depth_over_tol"] == 0, c


def test_render_batch_poses(vb, oracle):
    V, F = vb.synth.load_chair()
    poses = vb.synth.render_poses(12)
    r = gpu_renderer(vb)
    r.SetMesh(V, F)
    depth, z24 = r.RenderDepthBatch(list(poses), want_z24=True)
    for i in range(len(poses)):
        oz, od = oracle_render(oracle, V, F, poses[i])
        assert (z24[i] == oz).all(), i
        assert (depth[i] == od).all(), i


def test_render_clipping_and_big_triangles(vb, oracle):
    """Cube straddling the near plane and filling the image (large pixel boxes -> the queued path),
    camera pose not identity."""
    V, F = vb.synth.cube_mesh()
    pose = vb.synth.make_T(vb.synth.rot_xyz(0.1, -0.2, 0.05), [0.05, -0.02, 0.1])
    for t, s in (([0, 0, 0.3], 1.0), ([0.1, 0.05, 0.0], 1.0), ([0, 0, 9.8], 1.0), ([0, 0, 2.0], 0.01), ([0, 0, -3.0], 1.0)):
        model = vb.synth.make_T(vb.synth.rot_xyz(0.3, 0.5, -0.2) * s, t)
        r = gpu_renderer(vb, pose=pose)
        r.SetMesh(V, F)
        depth, z24 = r.RenderDepthBatch([model], want_z24=True)
        oz, od = oracle_render(oracle, V, F, model, pose=pose)
        assert (z24[0] == oz).all(), t
        assert (depth[0] == od).all(), t
    # mixed meshes in one batch, including an empty one
    chair = vb.synth.load_chair()
    meshes = [chair, (V, F), (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))]
    models = [vb.synth.make_T(np.eye(3), [0, 0, 1.5]), vb.synth.make_T(np.eye(3), [0, 0, 2.0]), np.eye(4)]
    r = gpu_renderer(vb)
    depth, z24 = r.RenderDepthBatch(models, meshes=meshes, want_z24=True)
    for i, ((v, f), m) in enumerate(zip(meshes, models)):
        oz, od = oracle_render(oracle, v, f, m)
        assert (z24[i] == oz).all(), i
    assert (z24[2] == oracle.ZMAX24).all() and (depth[2] == 1.0).all()


def test_render_other_resolution(vb, oracle):
    cam = dict(zn=0.1, zf=5.0, fx=480.0, fy=470.0, cx=479.5, cy=249.5, H=500, W=960)  # the reference's 500x960 target
    V, F = vb.synth.load_chair()
    model = vb.synth.make_T(vb.synth.rot_y(1.0), [0.2, 0.1, 1.2])
    r = gpu_renderer(vb, cam)
    r.SetMesh(V, F)
    out = np.zeros((500, 960), np.float32)
    r.RenderDepth(model, out)
    oz, od = oracle_render(oracle, V, F, model, cam=cam)
    assert (out == od).all()
    with pytest.raises(ValueError):
        r.RenderDepth(model, np.zeros((10, 10), np.float32))  # the reference CHECKs the size


def test_voxel_downsample_bitexact(vb, oracle, unit_rand):
    from conftest import small_scene
    d = small_scene(n_scene=80000, n_objects=2, m=100)
    xyz, nrm = d["scene_xyz"], d["scene_nrm"].copy()
    nrm[::97] = np.nan  # NaN normals are skipped (DownSample.cpp:52-56)
    for voxel in (0.05, 0.02):
        g = vb.reg.VoxelDownSample(vb.reg.PointCloud(xyz, nrm), voxel)
        o_p, o_n = oracle.voxel_downsample(xyz, voxel, nrm)
        assert len(g.points_) == len(o_p)
        assert (g.points_ == o_p).all()
        assert np.array_equal(g.normals_, o_n, equal_nan=True)
    g = vb.reg.VoxelDownSample(xyz, 0.05)
    assert (g.points_ == oracle.voxel_downsample(xyz, 0.05)).all() and g.normals_ is None
    # the reference's golden input (UnitTest PointCloud.cpp:677-782): 20 points, each its own voxel
    r, rmax = unit_rand
    pts = (r[:60] * (1000.0 / rmax)).reshape(20, 3)
    g = vb.reg.VoxelDownSample(pts, 0.5)
    assert len(g.points_) == 20
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert (key(g.points_) == key(pts)).all()
    assert len(vb.reg.VoxelDownSample(pts, 0.0).points_) == 0  # voxel_size <= 0 -> empty cloud


def test_render_device_resident_outputs(vb, oracle):
    """vb200_render_depth_batch_ex writing straight into caller-owned DEVICE buffers."""
    torch = pytest.importorskip("torch")
    V, F = vb.synth.load_chair()
    poses = vb.synth.render_poses(6, seed=5)
    r = gpu_renderer(vb)
    r.SetMesh(V, F)
    d_depth = torch.empty((6, 480, 640), dtype=torch.float32, device="cuda")
    d_z = torch.empty((6, 480, 640), dtype=torch.int32, device="cuda")
    ms = r.RenderDepthBatchDevice(list(poses), d_depth.data_ptr(), d_z.data_ptr())
    assert ms > 0
    for i in range(6):
        oz, od = oracle_render(oracle, V, F, poses[i])
        assert (d_z[i].cpu().numpy().view(np.uint32) == oz).all()
        assert (d_depth[i].cpu().numpy() == od).all()


def test_render_edge_and_mask(vb, oracle):
    """RenderEdge / RenderMask (render/renderer.cpp:353-433, edge_detection.frag) against the restated shader."""
    V, F = vb.synth.load_chair()
    poses = vb.synth.render_poses(5, seed=8)
    r = gpu_renderer(vb)
    r.SetMesh(V, F)
    edge, mask = r.RenderEdgeMaskBatch(list(poses))
    for i in range(len(poses)):
        oz, _ = oracle_render(oracle, V, F, poses[i])
        oe, om = oracle.render_edge(oz), oracle.render_mask(oz)
        assert (edge[i] == oe).all(), i
        assert (mask[i] == om).all(), i
        assert oe.max() == 255 and (oe[:5] == 0).all() and (oe[:, -5:] == 0).all()
        assert 0 < (om == 255).sum() < om.size
    assert (r.RenderEdge(poses[0]) == edge[0]).all() and (r.RenderMask(poses[0]) == mask[0]).all()
    # other shader uniforms than the reference's fixed 0.05 / 2.0
    e2, _ = r.RenderEdgeMaskBatch([poses[1]], edge_z_near=0.05, edge_z_far=10.0)
    oz, _ = oracle_render(oracle, V, F, poses[1])
    assert (e2[0] == oracle.render_edge(oz, 0.05, 10.0)).all()
