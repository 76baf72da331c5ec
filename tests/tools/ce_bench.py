import sys, os, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from conftest import sub
L = sub("_lib"); fn = sub("functional")
dev = torch.device("cuda", 0)
x = torch.randn(8, 19, 512, 1024, device=dev)
lab = torch.randint(0, 20, (8, 512, 1024), device=dev).float(); lab[lab == 19] = 255
def run():
    return fn.cross_entropy(x, lab)
for mode in ("regs", "generic"):
    if mode == "generic": os.environ["S2R_CE_GENERIC"] = "1"
    l = run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5): l = run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(mode, "loss %.6f" % float(l), "%.1f us per call (forward: loss + unscaled gradient)" % (e0.elapsed_time(e1) / 5 * 1e3))
