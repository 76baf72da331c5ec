"""Run-to-run reproducibility of ONE eager adaptation step from the same seed in the same process (GPU box):
losses, flat gradients and BN buffers of runs 1..n against run 0.  Differences beyond atomics-order noise point at
reads of uninitialised memory (first run: fresh cudaMalloc blocks; later runs: recycled blocks) or stream races.
    python tests/tools/determinism_check.py [iterations] [B] [H] [W]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (4, 64, 96)


def inputs(it):
    g = torch.Generator().manual_seed(1000 + it)
    src = torch.randn(B, 3, H, W, generator=g); tgt = torch.randn(B, 3, H, W, generator=g)
    lab = torch.randint(0, 20, (B, H, W), generator=g).float(); lab[lab == 19] = 255
    return src.cuda(), lab.cuda(), tgt.cuda()


def run():
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    G._s2r_no_dropout = True
    G.cuda().train(); D.cuda().train()
    step = sub("steps").AdaptStep(G, D, lr=5e-4, epochs=1, iters_per_epoch=10)
    hist = []
    for it in range(n_it):
        out = step(*inputs(it), i=it, epoch=0)
        torch.cuda.synchronize()
        hist.append(dict(losses=torch.stack([out[k] for k in ('loss_seg', 'loss_adv', 'loss_D_src', 'loss_D_tgt')]).double().cpu(),
                         gG=step.optimizer.flat_grad.clone().cpu(), gD=step.optimizer_D.flat_grad.clone().cpu(),
                         bn0=G.backbone.features[0][1].running_mean.clone().cpu(), bnL=G.decoder.last_conv[5].running_var.clone().cpu(),
                         views={id(p): (k, o, n) for (p, o, n), k in zip(step.optimizer._views, [k for k, _ in G.named_parameters()])}))
    names = [(k, o, n) for k, o, n in hist[0]['views'].values()]
    global DNAMES
    DNAMES = [(k, p.numel()) for k, p in D.named_parameters()]
    global DSHAPES
    DSHAPES = [(k, p.numel(), tuple(p.shape)) for k, p in D.named_parameters()]
    return hist, names


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


import gc
MODES = os.environ.get("DET_MODES", "plain,two-streams-off,wgrad-stream-off").split(",")
for mode in MODES:
    if mode == "two-streams-off":
        os.environ["S2R_OVERLAP"] = "0"
    if mode == "wgrad-stream-off":
        sub("engine").WGRAD_STREAM = False
    ref, names = run()
    for r in range(1, int(os.environ.get("DET_RUNS", "3"))):
        if os.environ.get("DET_GC"):
            gc.collect(); torch.cuda.empty_cache()
        cur, _ = run()
        for it in range(n_it):
            a, b = cur[it], ref[it]
            worst = sorted(((rel(a['gG'][o:o + n], b['gG'][o:o + n]), k) for k, o, n in names), reverse=True)[:3]
            if os.environ.get("DET_DETAIL") and rel(a['gD'], b['gD']) > 1e-4:
                print("   loss diffs", (a['losses'] - b['losses']).tolist())
                off = 0
                for k_, n_, shp in DSHAPES:
                    if len(shp) == 4 and rel(a['gD'][off:off + n_], b['gD'][off:off + n_]) > 1e-4:
                        ga, gb = a['gD'][off:off + n_].view(shp).double(), b['gD'][off:off + n_].view(shp).double()
                        d_ = (ga - gb)
                        print("   %s per-tap rel:" % k_, [round(float(d_[:, :, i // shp[3], i % shp[3]].norm() / (gb[:, :, i // shp[3], i % shp[3]].norm() + 1e-30)), 4) for i in range(shp[2] * shp[3])])
                        print("   %s per-64-cout rel:" % k_, [round(float(d_[i:i + 64].norm() / (gb[i:i + 64].norm() + 1e-30)), 4) for i in range(0, shp[0], 64)])
                        print("   %s per-32-cin rel:" % k_, [round(float(d_[:, i:i + 32].norm() / (gb[:, i:i + 32].norm() + 1e-30)), 4) for i in range(0, shp[1], 32)])
                        nz = (d_.abs() > 1e-6 * gb.abs().max()).double().mean()
                        print("   %s fraction of differing elements %.4f; ratio a/b median %.4f" % (k_, float(nz), float((ga / (gb + 1e-30)).median())))
                    off += n_
                off = 0
                for k_, n_ in DNAMES:
                    print("   D.%s grad rel %.3e norm %.3e" % (k_, rel(a['gD'][off:off + n_], b['gD'][off:off + n_]), float(b['gD'][off:off + n_].norm())))
                    off += n_
            print("%s run %d it %d: losses max|d| %.3e  gradG %.3e gradD %.3e  bn0 %.3e bnL %.3e  worst %s" % (
                mode, r, it, float((a['losses'] - b['losses']).abs().max()), rel(a['gG'], b['gG']), rel(a['gD'], b['gD']),
                rel(a['bn0'], b['bn0']), rel(a['bnL'], b['bnL']), [(round(v, 5), k) for v, k in worst]))
