"""Device time and achieved HBM bandwidth of the depthwise 3x3 entry points on the real block shapes of the
B=8, 512x1024 step (CUDA-graph replay of 10 launches).  Algorithmic bytes: forward = in + out (bf16);
backward (dgrad+wgrad) = dy + x + g.  GPU box: python tests/tools/dw_bench.py [filter]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
eng = sub("engine"); L = sub("_lib")
dev = torch.device("cuda", 0)
cx = eng.Ctx(dev, True)
PEAK = 6544.3
# name, H, W, C, stride, dil, halo, count per G pass
SHAPES = [("f1 32 256x512", 256, 512, 32, 1, 1, False, 1), ("f2 96 256x512 s2", 256, 512, 96, 2, 1, True, 1),
          ("f3 144 128x256", 128, 256, 144, 1, 1, True, 1), ("f4 144 128x256 s2", 128, 256, 144, 2, 1, True, 1),
          ("f5 192 64x128", 64, 128, 192, 1, 1, True, 2), ("f7 192 64x128 s2", 64, 128, 192, 2, 1, True, 1),
          ("f8 384 32x64", 32, 64, 384, 1, 1, True, 4), ("f12 576 32x64", 32, 64, 576, 1, 1, True, 3),
          ("f15 960 32x64", 32, 64, 960, 1, 1, True, 2), ("f17 960 32x64 d2", 32, 64, 960, 1, 2, True, 1)]
flt = sys.argv[1] if len(sys.argv) > 1 else ""
N = 8


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
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


vp = lambda t: C.c_void_p(t.data_ptr())
tot = {}
print("%-22s %-8s %9s %8s %6s" % ("block", "op", "us", "GB/s", "frac"))
for name, H, W, Cc, s, d, halo, cnt in SHAPES:
    if flt and flt not in name: continue
    x = eng.Act((torch.randn(N, H, W, Cc, device=dev) * 2).to(torch.bfloat16))
    ss = torch.cat([torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev)]).contiguous()
    mi = torch.cat([torch.randn(Cc, device=dev) * 0.1, torch.rand(Cc, device=dev) + 0.5]).contiguous()
    st = eng.BNState(ss, mi, 1.0, False)
    w = torch.randn(Cc, 1, 3, 3, device=dev) * 0.3
    Ho, Wo = eng.conv_out_hw(H, W, 3, 3, s, d, d)
    stats = cx.f64(2 * Cc)
    ext = d if halo else 0
    dy = eng.Act(torch.randn(N, Ho, Wo, Cc, device=dev).to(torch.bfloat16))
    GEXT = bool(os.environ.get('S2R_BENCH_GEXT'))   # 1: extended gradient layout (border stored)
    g = cx.new(N, H + 2 * ext, W + 2 * ext, Cc) if GEXT else cx.new(N, H, W, Cc)   # default: unextended, as the engine requests it
    bs = cx.f64(2 * Cc); dw = torch.zeros_like(w)
    for mode in ("new", "generic"):
        if mode == "generic":
            if s != 1 or d != 1: continue
            os.environ["S2R_DW_GENERIC"] = "1"
        else:
            os.environ.pop("S2R_DW_GENERIC", None)
        tf = timed(lambda: eng.dw_fwd(cx, x, st, L.ACT_RELU6, halo, w, s, d, d, stats))
        tb = timed(lambda: L.call("s2r_dwconv3x3_bwd", dy.vp(), vp(w), x.vp(), vp(ss), vp(mi), L.ACT_RELU6, 1 if halo else 0,
                                  0 if GEXT else 1, g.vp(), vp(bs), vp(dw), N, H, W, Cc, s, d, d, cx.stream))
        bf_ = (N * H * W * Cc + N * Ho * Wo * Cc) * 2
        bb = (N * Ho * Wo * Cc + 2 * N * H * W * Cc) * 2
        print("%-22s %-8s %9.1f %8.0f %6.3f" % (name, "fwd/" + mode[:3], tf, bf_ / tf / 1e3, bf_ / tf / 1e3 / PEAK))
        print("%-22s %-8s %9.1f %8.0f %6.3f" % (name, "bwd/" + mode[:3], tb, bb / tb / 1e3, bb / tb / 1e3 / PEAK))
        if mode == "new":
            tot["fwd"] = tot.get("fwd", 0) + tf * cnt; tot["bwd"] = tot.get("bwd", 0) + tb * cnt
os.environ.pop("S2R_DW_GENERIC", None)
print("per G pass (x2 per step): fwd %.0f us, bwd %.0f us" % (tot.get("fwd", 0), tot.get("bwd", 0)))
