"""CPU experiment behind DESIGN.md section 4 ("mIoU within 0.1 points after a fixed-seed short run"): the oracle's
adaptation step (oracle/ref_port.adapt_step, bit-identical to the reference) is run for a short fixed-seed schedule on
a learnable synthetic task -- images whose pixels are a class colour + noise, one class per image (`const`) or two
(`halves`), a colour-shifted target domain -- once in fp32 and once with its tensors rounded to bf16 where the B200
path stores them (tests/emul.py), then evaluated in eval mode on fixed validation images.  The two runs share weights,
data and schedule; their mIoU difference is what the operand format alone does to a short run from random init.

    python tests/tools/shortrun_miou_cpu.py const 60 2e-3     # mode, steps, lr   (about a minute on 8 cores)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import ref_port as O
from emul import emulate_bf16
import contextlib

COLORS = torch.tensor([[a,b,c] for a in (-1.5,0.,1.5) for b in (-1.5,0.,1.5) for c in (-1.5,0.,1.5)][:19])

def batch(seed, n, H, W, shift=0.0, mode='halves'):
    g = torch.Generator().manual_seed(seed)
    lab = torch.empty(n, H, W)
    for k in range(n):
        if mode == 'const':
            lab[k] = float(torch.randint(0, 19, (1,), generator=g))
        else:
            c = torch.randint(0, 19, (2,), generator=g)
            cut = int(torch.randint(W // 4, 3 * W // 4, (1,), generator=g))
            lab[k, :, :cut] = float(c[0]); lab[k, :, cut:] = float(c[1])
    img = COLORS[lab.long()].permute(0, 3, 1, 2) + 0.3 * torch.randn(n, 3, H, W, generator=g) + shift
    lab = lab.clone()
    lab[torch.rand(n, H, W, generator=g) < 0.02] = 255
    return img.contiguous(), lab

def run(emul, steps=40, lr=1e-2, n=4, H=65, W=129, mode='halves', seed=1):
    g_sd, d_sd = O.init_deeplab(seed=seed), O.init_discriminator(seed=seed+1)
    for sd in (g_sd, d_sd):
        for v in O.leaf_params(sd).values(): v.requires_grad_(True)
    one, ten = O.split_lr_groups(list(O.leaf_params(g_sd).keys()))
    opt = torch.optim.SGD([{'params': [g_sd[k] for k in one], 'lr': lr}, {'params': [g_sd[k] for k in ten], 'lr': 10*lr}], momentum=0.9, weight_decay=5e-4)
    opt_d = torch.optim.Adam(list(O.leaf_params(d_sd).values()), lr=1e-4, betas=(0.9, 0.99))
    ctx = emulate_bf16() if emul else contextlib.nullcontext()
    cfg = O.BNCfg(True)
    with ctx:
        for it in range(steps):
            cur = O.poly_lr(lr, it, steps)
            opt.param_groups[0]['lr'] = cur; opt.param_groups[1]['lr'] = cur * 10
            for gp in opt_d.param_groups: gp['lr'] = cur          # train_adapt.py: scheduler overwrites D's lr (group 0)
            src, lab = batch(1000 + it, n, H, W, 0.0, mode)
            tgt, _ = batch(5000 + it, n, H, W, 0.3, mode)
            losses = O.adapt_step(g_sd, d_sd, opt, opt_d, src, lab, tgt, cfg, drop=False)
        cm = np.zeros((19, 19), np.int64)
        with torch.no_grad():
            for k in range(4):
                x, lab = batch(9000 + k, 4, H, W, 0.0, mode)
                out = O.deeplab_forward(g_sd, x, O.BNCfg(False), 16, drop=False)
                cm += O.confusion_matrix(lab.numpy(), out.argmax(1).numpy(), 19)
    m = O.evaluator_metrics(cm)
    return losses, m['mIoU'], m['PA']

if __name__ == "__main__":
    mode, steps, lr = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    t = time.time()
    a, b = run(False, steps, lr, mode=mode), run(True, steps, lr, mode=mode)
    print(mode, steps, lr, "fp32 (losses, mIoU, PA)", a, "bf16-emulated", b, "|d mIoU| = %.3f points" % (100 * abs(a[1] - b[1])),
          "%.0f s" % (time.time() - t), flush=True)
