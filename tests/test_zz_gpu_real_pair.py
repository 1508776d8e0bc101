"""A second real-data case on the GPU: cloud_bin_2 -> cloud_bin_1 from the pairwise initial alignment the
reference ships (examples/TestData/ICP/init.log, entry "1 2").  Expected outputs are the unmodified
reference's (tests/golden/icp_pair12.npz, written by scripts/make_golden.py; the oracle restatement reproduces
them to 1e-11 in tests/test_oracle_golden.py).  Added after the round's GPU budget was spent: first run is the
round-end one (the file sorts last so that it cannot mask another test under -x)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_second_real_pair(vb, kat, pair12):
    s = pair12["src"].astype(np.float64)
    t, tn = kat["tgt"].astype(np.float64), kat["tgt_nrm"].astype(np.float64)
    sc = vb.reg.Scene(vb.reg.PointCloud(t, tn), 0.02)
    src = vb.reg.PointCloud(s, s)  # normals: presence only (Registration.cpp:152-157)
    m = len(s)
    ev = vb.reg.EvaluateRegistration(src, sc, 0.02, pair12["init"])
    # the initial pass has no arithmetic before the search: identical correspondences
    assert len(ev.correspondence_set_) == int(pair12["ref_eval"][2])
    assert ev.fitness_ == pair12["ref_eval"][0] and abs(ev.inlier_rmse_ - pair12["ref_eval"][1]) < 1e-12
    for est, key in ((vb.reg.TransformationEstimationPointToPoint(), "ref_p2p"),
                     (vb.reg.TransformationEstimationPointToPlane(), "ref_p2l")):
        r = vb.reg.RegistrationICP(src, sc, 0.02, pair12["init"], est)
        fit, rmse, nc = pair12[key]
        # poses agree far inside the 1e-4 rad / 1e-3 m bar; a correspondence sitting on the radius can flip
        # with the last bits of the transform, hence the slack of two on the count
        assert np.allclose(r.transformation_, pair12[key + "_T"], atol=1e-8), key
        assert abs(len(r.correspondence_set_) - int(nc)) <= 2, key
        assert abs(r.fitness_ - fit) <= 2.0 / m and abs(r.inlier_rmse_ - rmse) < 1e-7, key
