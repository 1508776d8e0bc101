"""dev tool: launches of the exhaustive KNN (2 and 8 queries against 1e7 points) and of the warp-per-query grid search, for ncu."""
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from visma_b200 import _lib, registration as reg, synth
L=_lib.lib()
N=10_000_000
t=torch.rand(N,3,dtype=torch.float64,device="cuda")*6
for nb in (2,8):
    q=t[:nb].clone()+0.001
    i=torch.empty(nb,dtype=torch.int32,device="cuda"); d=torch.empty(nb,dtype=torch.float64,device="cuda")
    for _ in range(1):
        _lib.check(L.vb200_knn1_bruteforce_device(C.c_void_p(t.data_ptr()),N,C.c_void_p(q.data_ptr()),nb,0.075,0,C.c_void_p(i.data_ptr()),C.c_void_p(d.data_ptr()),None))
    torch.cuda.synchronize()

d = synth.make_room_scene(1_000_000, 8, 10)
sc = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), 0.075)
qq = synth.knn_queries(d["scene_xyz"], 10000)
for _ in range(2):
    sc.SearchHybrid1(qq, 0.075)
