"""Throughput of the two other loop bodies of the hot path on one B200 (eager launches, CUDA events):
  FeatureStep  train.py:163-216      (BASELINE config 4): backbone + ASPP + decoder + DomainClassifer, B=8, 512x1024
  ValStep      val_adapt.py:122-135  (BASELINE config 5): eval forward at 1x3x1024x2048 + fused argmax/confusion matrix
GPU box:  python tests/tools/other_steps_bench.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
import bench
dev = torch.device("cuda", 0)
nn = torch.nn


def timed(fn, n, warm=2):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


torch.manual_seed(1)
bb = sub("modeling.backbone.mobilenet").MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d).to(dev).train()
aspp = sub("modeling.assp").ASPP('mobilenet', 16, nn.BatchNorm2d).to(dev).train()
dec = sub("modeling.decoder").Decoder(19, 'mobilenet', nn.BatchNorm2d).to(dev).train()
dc = sub("modeling.domian").DomainClassifer('mobilenet', nn.BatchNorm2d).to(dev).train()
fstep = sub("steps").FeatureStep(bb, aspp, dec, dc, lr=5e-4, epochs=1, iters_per_epoch=100)
src, lab, tgt = (t.to(dev) for t in bench.synth(1000, 8, 512, 1024))
ms = timed(lambda i: fstep(src, lab, tgt, i=i), 5)
print("feature step (train.py:163-216), B=8 512x1024: %.1f ms/step = %.1f img-pairs/s (eager)" % (ms, 8 / ms * 1e3))
fstep.capture(src, lab, tgt, warmup=1)
ms = timed(lambda i: fstep.replay(src, lab, tgt, i=i), 10)
print("feature step, CUDA graph: %.1f ms/step = %.1f img-pairs/s" % (ms, 8 / ms * 1e3))

G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False).to(dev).eval()
vstep = sub("steps").ValStep(G, 19)
g = torch.Generator().manual_seed(5)
img = torch.randn(1, 3, 1024, 2048, generator=g).to(dev)
tl = torch.randint(0, 19, (1, 1024, 2048), generator=g).float()
tl[torch.rand(1, 1024, 2048, generator=g) < 0.05] = 255
tl = tl.to(dev)
ms = timed(lambda i: vstep(img, tl), 20, warm=3)
cm = vstep.evaluator.confusion_matrix
print("val step (val_adapt.py:122-135), 1x3x1024x2048, eager: %.2f ms/img = %.1f img/s; confusion-matrix total %d = %d valid pixels x %d images"
      % (ms, 1e3 / ms, int(cm.sum()), int((tl != 255).sum()), 23))
vstep.evaluator.reset()
vstep.capture(img, tl)
ms = timed(lambda i: vstep.replay(img, tl), 50, warm=3)
cm2 = vstep.evaluator.confusion_matrix
print("val step, CUDA graph: %.2f ms/img = %.1f img/s; confusion matrix identical per image: %s"
      % (ms, 1e3 / ms, bool((cm2 / 53 == cm / 23).all())))
