"""An INDEPENDENT depth rasteriser for checking the renderer's conventions (test infrastructure).

oracle/raster_oracle.c restates the reference's OpenGL pipeline step by step — glm::frustum with the flipped
top/bottom, view = diag(1,-1,-1,1), viewport transform, bottom-up read-back (render/renderer.cpp:238-351) — and the
CUDA rasteriser is bit-exact against it; but both were written by the same hand from the same reading of the
code.  This file takes the other route: the closed-form pinhole mapping the chain of GL conventions must collapse
to (derived from renderer.cpp:259-267,343):

    c = pose * model * [p, 1]            camera frame, +Z forward, +Y down
    u = fx X / Z + cx,  v = fy Y / Z + cy
    x_w = u W / (W - 1),  y_w = v H / (H - 1)         (the frustum spans pixels 0 .. W-1, the viewport is W wide)
    z_w = ((zf + zn) / (zf - zn) - 2 zf zn / ((zf - zn) Z) + 1) / 2
    output pixel (col i, row j; row 0 = image top) is covered iff (i + 1/2, j + 1/2) lies in the triangle (x_w, y_w);
    z_w interpolated linearly in window space; depth = min over triangles of round(z_w (2^24 - 1)).

in float64 numpy with plain barycentric coordinates — no matrices, no clip space, no sub-pixel snapping, no fill
rule.  Pixels whose centre lies within `edge_tol` pixels of a triangle edge are reported as uncertain (there the
1/256-pixel snapping and the top-left rule of the canonical rasteriser decide).  Meshes must lie between the near
and far planes (no clipping here)."""
import numpy as np

ZMAX24 = (1 << 24) - 1


def render_depth_indep(V, F, model, pose, zn, zf, fx, fy, cx, cy, H, W, edge_tol=0.02):
    V = np.asarray(V, np.float64)
    T = np.asarray(pose, np.float64).reshape(4, 4) @ np.asarray(model, np.float64).reshape(4, 4)
    c = V @ T[:3, :3].T + T[:3, 3]
    assert (c[:, 2] > zn).all() and (c[:, 2] < zf).all(), "independent rasteriser does not clip"
    xw = (fx * c[:, 0] / c[:, 2] + cx) * W / (W - 1.0)
    yw = (fy * c[:, 1] / c[:, 2] + cy) * H / (H - 1.0)
    zw = 0.5 * ((zf + zn) / (zf - zn) - 2.0 * zf * zn / ((zf - zn) * c[:, 2]) + 1.0)
    depth = np.full((H, W), float(ZMAX24))
    slope = np.zeros((H, W))  # |dz/dx| + |dz/dy| of the winning triangle, in 24-bit units per pixel
    covered = np.zeros((H, W), bool)
    uncertain = np.zeros((H, W), bool)
    for tri in np.asarray(F):
        x, y, z = xw[tri], yw[tri], zw[tri]
        area = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0])
        i0, i1 = int(np.floor(x.min() - 1)), int(np.ceil(x.max() + 1))
        j0, j1 = int(np.floor(y.min() - 1)), int(np.ceil(y.max() + 1))
        i0, j0, i1, j1 = max(i0, 0), max(j0, 0), min(i1, W - 1), min(j1, H - 1)
        if i0 > i1 or j0 > j1:
            continue
        px = np.arange(i0, i1 + 1)[None, :] + 0.5
        py = np.arange(j0, j1 + 1)[:, None] + 0.5
        # signed distances (pixels) from the three edges, positive inside
        dist = []
        for a, b in ((0, 1), (1, 2), (2, 0)):
            ex, ey = x[b] - x[a], y[b] - y[a]
            ln = np.hypot(ex, ey)
            if ln < 1e-12:
                dist = None
                break
            s = ((px - x[a]) * ey - (py - y[a]) * ex) / ln
            dist.append(s if area < 0 else -s)
        if dist is None or abs(area) < 1e-12:
            # degenerate in the image: covers nothing, but its neighbourhood is uncertain
            uncertain[j0:j1 + 1, i0:i1 + 1] |= (np.abs(px - x.mean()) < 1) & (np.abs(py - y.mean()) < 1)
            continue
        dmin = np.minimum(np.minimum(dist[0], dist[1]), dist[2])
        inside = dmin > 0
        uncertain[j0:j1 + 1, i0:i1 + 1] |= np.abs(dmin) <= edge_tol
        if not inside.any():
            continue
        # barycentric interpolation of z_w (linear in window space)
        w1 = ((px - x[0]) * (y[2] - y[0]) - (py - y[0]) * (x[2] - x[0])) / area
        w2 = ((x[1] - x[0]) * (py - y[0]) - (y[1] - y[0]) * (px - x[0])) / area
        zz = z[0] + w1 * (z[1] - z[0]) + w2 * (z[2] - z[0])
        q = np.rint(zz * ZMAX24)
        gx = ((z[1] - z[0]) * (y[2] - y[0]) - (z[2] - z[0]) * (y[1] - y[0])) / area
        gy = ((x[1] - x[0]) * (z[2] - z[0]) - (x[2] - x[0]) * (z[1] - z[0])) / area
        sub, ssub = depth[j0:j1 + 1, i0:i1 + 1], slope[j0:j1 + 1, i0:i1 + 1]
        win = inside & (q < sub)
        sub[win] = q[win]
        ssub[win] = (abs(gx) + abs(gy)) * ZMAX24
        covered[j0:j1 + 1, i0:i1 + 1] |= inside
    return depth, covered, uncertain, slope


def compare(z24, depth, covered, uncertain, slope, base_units=4):
    """-> dict of what must hold between a canonical z-buffer and the independent one.  The canonical rasteriser snaps
    vertices to 1/256 pixel before interpolating, so on a triangle whose depth changes by s units per pixel the two
    may differ by s/128 (both axes, both roundings) on top of a few units of f32-vs-f64 arithmetic."""
    z24 = np.asarray(z24, np.int64)
    cov_c = z24 != ZMAX24
    sure = ~uncertain
    both = cov_c & covered & sure
    dz = np.abs(z24[both] - depth[both]) - slope[both] / 128.0
    max_units = base_units
    return dict(coverage_mismatch_sure=int((cov_c != covered)[sure].sum()), n_sure=int(sure.sum()),
                n_uncertain=int(uncertain.sum()), n_covered=int(covered.sum()),
                depth_over_tol=int((dz > max_units).sum()), depth_p999=float(np.quantile(dz, 0.999)) if len(dz) else 0.0,
                depth_max=float(dz.max()) if len(dz) else 0.0)
