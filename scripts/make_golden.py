"""Generates the committed golden fixtures under tests/golden/ from the reference checkout.
Run in the dev container (needs /root/reference and oracle/_ref/libvisma_ref.so):

    make -C oracle ref && python scripts/make_golden.py

  icp_kat.npz    Open3D's ICP tutorial data (examples/TestData/ICP/cloud_bin_{0,1}.pcd, float32 as stored),
                 the tutorial's trans_init (examples/Python/Basic/icp_registration.py:22-27), the numbers the
                 docs publish (docs/tutorial/Basic/icp_registration.rst:56-58,91-98,154-161) and the outputs of
                 the unmodified reference compiled here (full precision).
  unit_rand.npz  UnitTest::Rand's stream (srand(0) + glibc rand(), src/UnitTest/UnitTest.cpp:75-94) that the
                 reference's golden tests for 1-NN distances and VoxelDownSample draw their inputs from.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from read_pcd import read_pcd  # noqa: E402
from oracle import pyref  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

d = os.path.join(REF, "thirdparty/Open3D/examples/TestData/ICP/")
s, sn = read_pcd(d + "cloud_bin_0.pcd")
t, tn = read_pcd(d + "cloud_bin_1.pcd")
init = np.array([[0.862, 0.011, -0.507, 0.5], [-0.139, 0.967, -0.215, 0.7], [0.487, 0.255, 0.835, -1.4],
                 [0.0, 0.0, 0.0, 1.0]])
s64, t64, tn64, sn64 = (a.astype(np.float64) for a in (s, t, tn, sn))
ev = pyref.evaluate_registration(s64, t64, 0.02, init)
p2p = pyref.registration_icp(s64, t64, 0.02, init, pyref.P2P)
p2l = pyref.registration_icp(s64, t64, 0.02, init, pyref.P2PLANE, src_nrm=sn64, tgt_nrm=tn64)
kd = pyref.KDTree(t64)
tr_p2p = kd.icp_trace(s64, 0.02, init, pyref.P2P)
tr_p2l = kd.icp_trace(s64, 0.02, init, pyref.P2PLANE, src_nrm=sn64, tgt_nrm=tn64)
assert np.allclose(tr_p2p[-1, 3:].reshape(4, 4), p2p["T"], atol=1e-12)
np.savez_compressed(
    os.path.join(OUT, "icp_kat.npz"), src=s, tgt=t, tgt_nrm=tn, init=init,
    # published (docs) values: fitness, rmse, ncorr
    doc_eval=np.array([0.174723, 0.011771, 34741]), doc_p2p=np.array([0.372450, 0.007760, 74056]),
    doc_p2l=np.array([0.620972, 0.006581, 123471]),
    doc_p2p_T=np.array([[0.83924644, 0.01006041, -0.54390867, 0.64639961],
                        [-0.15102344, 0.96521988, -0.21491604, 0.75166079],
                        [0.52191123, 0.2616952, 0.81146378, -1.50303533], [0, 0, 0, 1.0]]),
    doc_p2l_T=np.array([[0.84023324, 0.00618369, -0.54244126, 0.64720943],
                        [-0.14752342, 0.96523919, -0.21724508, 0.81018928],
                        [0.52132423, 0.26174429, 0.81182576, -1.48366001], [0, 0, 0, 1.0]]),
    ref_eval=np.array([ev["fitness"], ev["rmse"], ev["ncorr"]]),
    ref_p2p=np.array([p2p["fitness"], p2p["rmse"], p2p["ncorr"]]), ref_p2p_T=p2p["T"],
    ref_p2l=np.array([p2l["fitness"], p2l["rmse"], p2l["ncorr"]]), ref_p2l_T=p2l["T"],
    ref_trace_p2p=tr_p2p, ref_trace_p2l=tr_p2l)
print("icp_kat:", ev, p2p["fitness"], p2l["fitness"], len(tr_p2p), len(tr_p2l))

libc = ctypes.CDLL("libc.so.6")
RAND_MAX = 2147483647
libc.srand(0)
stream = np.array([libc.rand() for _ in range(300)], np.int64)
np.savez_compressed(os.path.join(OUT, "unit_rand.npz"), rand=stream, rand_max=np.int64(RAND_MAX))
print("unit_rand:", stream[:3])

# A second real-data pair: cloud_bin_2 -> cloud_bin_1 from the pairwise initial alignment the reference ships
# (examples/TestData/ICP/init.log, entry "1 2": the transformation that brings fragment 2 into fragment 1's
# frame).  No published numbers exist for it: the expected outputs are the unmodified reference's (oracle/_ref).
s2, s2n = read_pcd(d + "cloud_bin_2.pcd")
log = open(d + "init.log").read().split()
k = [i for i in range(0, len(log), 19) if log[i] == "1" and log[i + 1] == "2"][0]
init12 = np.array([float(x) for x in log[k + 3:k + 19]]).reshape(4, 4)
s2_64, s2n_64 = s2.astype(np.float64), s2n.astype(np.float64)
ev = pyref.evaluate_registration(s2_64, t64, 0.02, init12)
p2p = pyref.registration_icp(s2_64, t64, 0.02, init12, pyref.P2P)
p2l = pyref.registration_icp(s2_64, t64, 0.02, init12, pyref.P2PLANE, src_nrm=s2n_64, tgt_nrm=tn64)
np.savez_compressed(
    os.path.join(OUT, "icp_pair12.npz"), src=s2, init=init12,
    ref_eval=np.array([ev["fitness"], ev["rmse"], ev["ncorr"]]),
    ref_p2p=np.array([p2p["fitness"], p2p["rmse"], p2p["ncorr"]]), ref_p2p_T=p2p["T"],
    ref_p2l=np.array([p2l["fitness"], p2l["rmse"], p2l["ncorr"]]), ref_p2l_T=p2l["T"])
print("icp_pair12:", ev["fitness"], ev["ncorr"], p2p["fitness"], p2p["ncorr"], p2l["fitness"], p2l["ncorr"])
