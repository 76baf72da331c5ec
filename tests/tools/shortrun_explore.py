import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from conftest import sub
from oracle import ref_port as O
import test_gpu_shortrun as T
for lr, steps in ((1e-3, 300), (2e-3, 300), (4e-3, 300), (2e-3, 600)):
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False).cuda().train()
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19).cuda().train()
    step = sub("steps").AdaptStep(G, D, lr=lr, epochs=1, iters_per_epoch=steps)
    tr = []
    for it in range(steps):
        src, lab = T.batch(1000 + it, 8, 128, 256); tgt, _ = T.batch(5000 + it, 8, 128, 256, 0.3)
        out = step(src.cuda(), lab.cuda(), tgt.cuda(), i=it, epoch=0)
        if it % 50 == 0 or it == steps - 1: tr.append(round(float(out['loss_seg']), 3))
    G.eval(); vstep = sub("steps").ValStep(G, 19)
    for k in range(4):
        x, lab = T.batch(9000 + k, 4, 128, 256); vstep(x.cuda(), lab.cuda())
    print(lr, steps, tr, "mIoU", float(vstep.evaluator.Mean_Intersection_over_Union()[0]), flush=True)
