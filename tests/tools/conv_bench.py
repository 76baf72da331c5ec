"""Device time of the dense-conv entry points on the real layer shapes of the B=8, 512x1024 step
(CUDA-graph replay of 10 launches, so host overhead is excluded).  GPU box: python tests/tools/conv_bench.py [filter]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
eng = sub("engine"); L = sub("_lib")
dev = torch.device("cuda", 0)
cx = eng.Ctx(dev, True)

# name, N, H, W, Cin, Cout, R, stride, pad, dil, what
SHAPES = [
    ("stem 3x3s2 3->32", 8, 512, 1024, 3, 32, 3, 2, 1, 1),
    ("f1 pw 32->16", 8, 256, 512, 32, 16, 1, 1, 0, 1),
    ("f2 pw 16->96", 8, 256, 512, 16, 96, 1, 1, 0, 1),
    ("f2 pw 96->24", 8, 128, 256, 96, 24, 1, 1, 0, 1),
    ("f3 pw 24->144", 8, 128, 256, 24, 144, 1, 1, 0, 1),
    ("f3 pw 144->24", 8, 128, 256, 144, 24, 1, 1, 0, 1),
    ("f5 pw 32->192", 8, 64, 128, 32, 192, 1, 1, 0, 1),
    ("f8 pw 64->384", 8, 32, 64, 64, 384, 1, 1, 0, 1),
    ("f8 pw 384->64", 8, 32, 64, 384, 64, 1, 1, 0, 1),
    ("f15 pw 160->960", 8, 32, 64, 160, 960, 1, 1, 0, 1),
    ("f17 pw 960->320", 8, 32, 64, 960, 320, 1, 1, 0, 1),
    ("aspp 3x3d12 320->256", 8, 32, 64, 320, 256, 3, 1, 12, 12),
    ("aspp 1x1 1280->256", 8, 32, 64, 1280, 256, 1, 1, 0, 1),
    ("dec 3x3 304->256", 8, 128, 256, 304, 256, 3, 1, 1, 1),
    ("dec 3x3 256->256", 8, 128, 256, 256, 256, 3, 1, 1, 1),
    ("dec 1x1 256->19", 8, 128, 256, 256, 19, 1, 1, 0, 1),
    ("D 4x4s2 19->64", 8, 512, 1024, 19, 64, 4, 2, 1, 1),
    ("D 4x4s2 64->128", 8, 256, 512, 64, 128, 4, 2, 1, 1),
    ("D 4x4s2 128->256", 8, 128, 256, 128, 256, 4, 2, 1, 1),
    ("D 4x4s2 256->512", 8, 64, 128, 256, 512, 4, 2, 1, 1),
]
flt = sys.argv[1] if len(sys.argv) > 1 else ""


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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("%-24s %-6s %9s %9s %8s %8s" % ("layer", "op", "tc us", "mma us", "GB/s", "TF/s"))
for name, N, H, W, Cin, Cout, R, s, p, d in SHAPES:
    if flt and flt not in name: continue
    Cp, Op = eng.round_up(Cin, 8), eng.round_up(Cout, 8)
    x = eng.Act(torch.randn(N, H, W, Cp, device=dev).to(torch.bfloat16)); x.C = Cin
    w = torch.nn.Parameter(torch.randn(Cout, Cin, R, R, device=dev) * 0.05)
    OH, OW = eng.conv_out_hw(H, W, R, R, s, p, d)
    out = cx.new(N, OH, OW, Op); out.C = Cout
    dy = eng.Act(torch.randn(N, OH, OW, Op, device=dev).to(torch.bfloat16)); dy.C = Cout
    dx = cx.new(N, H, W, Cp); dx.C = Cin
    stats = cx.f64(2 * Cout)
    flops = 2.0 * N * OH * OW * Cout * Cin * R * R
    byts = 2.0 * (N * H * W * Cin + N * OH * OW * Cout)
    for op in ("fwd", "dgrad", "wgrad"):
        res = []
        for force in (False, True):
            if op == "fwd": fn = lambda: eng.conv_fwd(cx, x, w, out, s, p, d, stats=None if os.environ.get("S2R_BENCH_NOSTATS") else stats, force_mma=force)
            elif op == "dgrad":
                if Cin < 8: res.append(float('nan')); continue
                fn = lambda: eng.conv_dgrad(cx, dy, w, dx, s, p, d, force_mma=force)
            else:
                if force: res.append(float('nan')); continue
                fn = lambda: eng.conv_wgrad(cx, x, dy, w, s, p, d)
            res.append(timed(fn))
        t = res[0]
        print("%-24s %-6s %9.1f %9.1f %8.0f %8.1f" % (name, op, res[0], res[1], byts / t / 1e3, flops / t / 1e6), flush=True)
