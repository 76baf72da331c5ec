"""tcgen05 conv path vs the mma.sync path vs torch fp32 on a list of shapes (GPU box)."""
import os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
torch.backends.cudnn.allow_tf32 = False
eng = sub("engine"); L = sub("_lib")
cx = eng.Ctx(torch.device("cuda", 0), True)

def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
def bf(t): return t.to(torch.bfloat16).float()
def act_of(t):
    N, C, H, W = t.shape
    buf = torch.zeros((N, H, W, eng.round_up(C, 8)), dtype=torch.bfloat16, device="cuda")
    buf[..., :C] = t.permute(0, 2, 3, 1).to(torch.bfloat16)
    a = eng.Act(buf); a.C = C
    return a
def nchw(a): return a.t[..., :a.C].float().permute(0, 3, 1, 2)

CASES = [(2, 16, 128, 64, 64, 1, 1, 0, 1), (2, 17, 23, 32, 16, 1, 1, 0, 1), (1, 32, 64, 320, 256, 3, 1, 6, 6),
         (2, 20, 28, 304, 256, 3, 1, 1, 1), (2, 32, 48, 24, 64, 4, 2, 1, 1), (2, 17, 25, 64, 128, 4, 2, 1, 1),
         (2, 33, 47, 8, 32, 3, 2, 1, 1), (2, 16, 24, 256, 19, 1, 1, 0, 1), (1, 9, 12, 1024, 1024, 3, 1, 1, 1),
         (8, 128, 256, 304, 256, 3, 1, 1, 1)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for ci, (N, H, W, Cin, Cout, R, s, p, d) in enumerate(CASES):
    if only is not None and ci != only: continue
    g = torch.Generator(device="cuda").manual_seed(ci)
    x = bf(torch.randn(N, Cin, H, W, device="cuda", generator=g))
    w = torch.nn.Parameter(torch.randn(Cout, Cin, R, R, device="cuda", generator=g) * (2.0 / (Cin * R * R)) ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv2d(x, bf(w.detach()), b, s, p, d)
    OH, OW = ref.shape[2:]
    xa = act_of(x)
    res = {}
    for name, force in (("mma", True), ("tc", False)):
        out = cx.new(N, OH, OW, eng.round_up(Cout, 8), zero=True); out.C = Cout
        st = cx.f64(2 * Cout)
        eng.conv_fwd(cx, xa, w, out, s, p, d, bias=b, stats=st, force_mma=force)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3): eng.conv_fwd(cx, xa, w, out, s, p, d, bias=b, force_mma=force)
        t1.record(); torch.cuda.synchronize()
        sref = torch.stack([ref.double().sum((0, 2, 3)), (ref.double() ** 2).sum((0, 2, 3))])
        res[name] = (rel(nchw(out), ref), rel(st.view(2, Cout), sref), t0.elapsed_time(t1) / 3)
    fl = 2.0 * N * OH * OW * Cout * Cin * R * R
    print("case %d %s  mma: err %.2e stats %.2e %.3f ms (%.0f TF/s) | tc: err %.2e stats %.2e %.3f ms (%.0f TF/s)" % (
        ci, (N, H, W, Cin, Cout, R, s, p, d), res["mma"][0], res["mma"][1], res["mma"][2], fl / res["mma"][2] / 1e9,
        res["tc"][0], res["tc"][1], res["tc"][2], fl / res["tc"][2] / 1e9), flush=True)
