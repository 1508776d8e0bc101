"""vb200_sample_mesh (feh::SamplePointCloudFromMesh, include/geometry.h:29-64): bit-exact against the seeded
restatement, and — because the reference seeds from the clock — statistical checks of what the reference's
algorithm is meant to produce: face frequencies proportional to area, samples on their triangles."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sample_mesh_matches_restatement(vb, oracle):
    V, F = vb.synth.load_chair()
    for seed, n in ((0, 50000), (12345678901, 1000), (7, 1)):
        p, nr = vb.reg.SamplePointCloudFromMesh(V, F, n, seed=seed, with_normals=True)
        op, on, _ = oracle.sample_mesh(V, F, n, seed)
        assert p.shape == (n, 3) and (p == op).all()
        assert np.allclose(nr, on, atol=1e-12)
    # reproducible, seed-dependent, prefix-stable (counter-based stream keyed by sample index)
    a = vb.reg.SamplePointCloudFromMesh(V, F, 2000, seed=3)
    b = vb.reg.SamplePointCloudFromMesh(V, F, 4000, seed=3)
    c = vb.reg.SamplePointCloudFromMesh(V, F, 2000, seed=4)
    assert (a == b[:2000]).all() and not (a == c).all()
    assert len(vb.reg.SamplePointCloudFromMesh(V, F, 0)) == 0


def test_sample_mesh_statistics(vb, oracle):
    V, F = vb.synth.load_chair()
    n = 400000
    p = vb.reg.SamplePointCloudFromMesh(V, F, n, seed=99)
    _, _, f = oracle.sample_mesh(V, F, n, 99)
    Vd = V.astype(np.float64)
    v0, e1, e2 = Vd[F[f, 0]], Vd[F[f, 1]] - Vd[F[f, 0]], Vd[F[f, 2]] - Vd[F[f, 0]]
    # every sample lies on its triangle: solve p = v0 + a e1 + b e2
    d = p - v0
    A = np.stack([e1, e2], 2)
    ab = np.einsum("nij,nj->ni", np.linalg.pinv(A), d)
    assert (ab > -1e-9).all() and (ab.sum(1) < 1 + 1e-9).all()
    assert np.abs(np.einsum("nij,nj->ni", A, ab) - d).max() < 1e-9
    # face histogram ~ area (chi-square per degree of freedom close to 1)
    area = 0.5 * np.linalg.norm(np.cross(Vd[F[:, 1]] - Vd[F[:, 0]], Vd[F[:, 2]] - Vd[F[:, 0]]), axis=1)
    expect = n * area / area.sum()
    counts = np.bincount(f, minlength=len(F))
    big = expect > 20
    chi2 = ((counts[big] - expect[big]) ** 2 / expect[big]).sum() / big.sum()
    assert 0.8 < chi2 < 1.2, chi2
    # same first moments as the numpy sampler the synthetic workloads use
    q, _ = vb.synth.sample_mesh(V, F, n, np.random.default_rng(5))
    assert np.abs(p.mean(0) - q.mean(0)).max() < 3e-3 and np.abs(p.std(0) - q.std(0)).max() < 3e-3
