"""Device time / achieved HBM bandwidth of the BatchNorm streaming kernels on shapes of the B=8, 512x1024 step
(CUDA-graph replay of 10 launches; S2R_BENCH_EAGER=1: plain launches for ncu).  GPU box: python tests/tools/bn_bench.py"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
eng = sub("engine"); L = sub("_lib")
dev = torch.device("cuda", 0)
cx = eng.Ctx(dev, True)
PEAK = 6544.3
SHAPES = [(1048576, 96), (262144, 144), (65536, 192), (16384, 384), (16384, 960), (262144, 24), (1048576, 16), (262144, 256), (16384, 64)]
vp = lambda t: C.c_void_p(t.data_ptr())


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    if os.environ.get("S2R_BENCH_EAGER"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 2 * 1e3
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


flt = sys.argv[1] if len(sys.argv) > 1 else ""
print("%-16s %-12s %9s %8s %6s" % ("P x C", "op", "us", "GB/s", "frac"))
for P, Cc in SHAPES:
    if flt and flt != str(Cc): continue
    x = (torch.randn(P, Cc, device=dev)).to(torch.bfloat16)
    dy = (torch.randn(P, Cc, device=dev)).to(torch.bfloat16)
    y = torch.empty_like(x)
    ss = torch.cat([torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev)]).contiguous()
    mi = torch.cat([torch.randn(Cc, device=dev) * 0.1, torch.rand(Cc, device=dev) + 0.5]).contiguous()
    sums = torch.zeros(2 * Cc, device=dev, dtype=torch.float64)
    ops = {
        "apply": (lambda: L.call("s2r_bn_apply_act", vp(x), P, Cc, Cc, 0, vp(ss), L.ACT_RELU6, None, 0.0, 0, None, vp(y), Cc, 0, cx.stream), 4),
        "apply+res": (lambda: L.call("s2r_bn_apply_act", vp(x), P, Cc, Cc, 0, vp(ss), L.ACT_NONE, vp(dy), 0.0, 0, None, vp(y), Cc, 0, cx.stream), 6),
        "sums": (lambda: L.call("s2r_channel_sums_bf16", vp(x), P, Cc, Cc, 0, vp(sums), cx.stream), 2),
        "bwd_reduce": (lambda: L.call("s2r_bn_bwd_reduce", vp(dy), Cc, 0, vp(x), Cc, 0, vp(mi), vp(ss), L.ACT_RELU6, 0.0, 0, None, P, Cc, vp(sums), cx.stream), 4),
        "bwd_apply": (lambda: L.call("s2r_bn_bwd_apply", vp(dy), Cc, 0, vp(x), Cc, 0, vp(mi), vp(ss), L.ACT_RELU6, 0.0, 0, None, vp(sums), float(P), P, Cc, vp(y), Cc, 0, None, None, 0, 0, 0, cx.stream), 6),
    }
    for name, (fn, bpe) in ops.items():
        t = timed(fn)
        b = P * Cc * bpe
        print("%-16s %-12s %9.1f %8.0f %6.3f" % ("%dx%d" % (P, Cc), name, t, b / t / 1e3, b / t / 1e3 / PEAK), flush=True)
