"""dev tool: pinned / pageable host<->device copy rates of the box (what bounds the e2e legs)."""
import time, torch, numpy as np
dev = torch.device("cuda", 0)
for mb in (19, 38, 157):
    n = mb << 20
    hp = torch.empty(n, dtype=torch.uint8).pin_memory()
    hq = torch.empty(n, dtype=torch.uint8)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    res = []
    for name, src, dst in (("H2D pinned", hp, d), ("D2H pinned", d, hp), ("H2D pageable", hq, d), ("D2H pageable", d, hq)):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        res.append("%s %.1f GB/s" % (name, n / min(ts) / 1e9))
    print("%d MB: " % mb + ", ".join(res))
