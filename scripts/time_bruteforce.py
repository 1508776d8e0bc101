"""dev tool: vb200_knn1_bruteforce_device at Q = 10 000 over the sweep sizes (CUDA events on the library stream, L2 flushed),
checked against the grid search bit for bit.   python scripts/time_bruteforce.py [N ...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from visma_b200 import registration as reg, synth, _lib
Q, R = 10_000, 0.075
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
for N in [int(float(a)) for a in sys.argv[1:]] or [100_000, 1_000_000, 10_000_000]:
    d = synth.make_room_scene(N, 8, 10)
    scene = reg.Scene(reg.PointCloud(d["scene_xyz"], d["scene_nrm"]), R)
    q = torch.from_numpy(synth.knn_queries(d["scene_xyz"], Q)).to(dev)
    idx = torch.empty(Q, dtype=torch.int32, device=dev); d2 = torch.empty(Q, dtype=torch.float64, device=dev)
    bidx = torch.empty_like(idx); bd2 = torch.empty_like(d2)
    _lib.check(L.vb200_knn1_device(scene.handle, C.c_void_p(q.data_ptr()), Q, R, C.c_void_p(idx.data_ptr()), C.c_void_p(d2.data_ptr())))
    tgt = torch.from_numpy(d["scene_xyz"]).to(dev)
    stream = torch.cuda.ExternalStream(scene.stream(), device=dev)
    fn = lambda: _lib.check(L.vb200_knn1_bruteforce_device(C.c_void_p(tgt.data_ptr()), N, C.c_void_p(q.data_ptr()), Q, R, 0,
                                                             C.c_void_p(bidx.data_ptr()), C.c_void_p(bd2.data_ptr()), C.c_void_p(scene.stream())))
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    same = bool((bidx == idx).all().item()) and bool((bd2 == d2).all().item())
    print(json.dumps({"N": N, "Q": Q, "bruteforce_ms": ms, "pairs_per_s": N * Q / ms * 1e3,
                      "fp32_frac": 3.0 * N * Q / (ms * 1e-3) / (128 * 148 * 1.965e9), "equals_grid": same}))
    scene.close()
