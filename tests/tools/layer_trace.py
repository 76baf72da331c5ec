"""Per-layer error profile of the B200 path against the fp32 oracle and the bf16-emulating oracle.
Run on a GPU box:  python tests/tools/layer_trace.py [H W N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub  # noqa: E402
from emul import emulate_bf16, traced  # noqa: E402
from oracle import ref_port as O  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    H, W, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (65, 97, 2)
    eng = sub("engine")
    for train in (True, False):
        torch.manual_seed(1)
        m = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
        m._s2r_no_dropout = True
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        x = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(0))
        t_ref, t_emu = [], []
        with traced(t_ref):
            o_ref = O.deeplab_forward({k: v.clone() for k, v in sd.items()}, x, O.BNCfg(train), 16, drop=False)
        with emulate_bf16(t_emu):
            o_emu = O.deeplab_forward({k: v.clone() for k, v in sd.items()}, x, O.BNCfg(train), 16, drop=False)
        m.cuda().train(train)
        eng.TRACE = []
        with torch.no_grad():
            out = m(x.cuda())
        mine = {n: a.t[..., a.off:a.off + a.C].float().permute(0, 3, 1, 2) for n, a in eng.TRACE}
        eng.TRACE = None
        print("mode", "train" if train else "eval", "size", (N, H, W))
        print("%-12s %10s %10s %10s" % ("layer", "mine/fp32", "mine/emul", "emul/fp32"))
        for (n, r), (_, e) in zip(t_ref, t_emu):
            print("%-12s %10.4f %10.4f %10.4f" % (n, rel(mine[n], r), rel(mine[n], e), rel(e, r)))
        print("%-12s %10.4f %10.4f %10.4f" % ("logits", rel(out, o_ref), rel(out, o_emu), rel(o_emu, o_ref)))


if __name__ == "__main__":
    main()
