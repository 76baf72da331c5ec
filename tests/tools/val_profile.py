"""Per-call device time of ONE eager validation image (1x3x1024x2048, eval forward + fused argmax / confusion matrix;
val_adapt.py:122-135): CUDA events around every C-ABI call, aggregated by (entry point, shape signature).
GPU box:  python tests/tools/val_profile.py [topN]"""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
L = sub("_lib")
dev = torch.device("cuda", 0)
torch.manual_seed(1)
G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False).to(dev).eval()
val = sub("steps").ValStep(G)
g = torch.Generator().manual_seed(3)
img = torch.randn(1, 3, 1024, 2048, generator=g).to(dev)
lab = torch.randint(0, 20, (1, 1024, 2048), generator=g).float()
lab[lab == 19] = 255
lab = lab.to(dev)
for i in range(2):
    val(img, lab)
torch.cuda.synchronize()
L.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); val(img, lab); e1.record()
torch.cuda.synchronize()
prof, L.PROFILE = L.PROFILE, None
agg = collections.defaultdict(lambda: [0, 0.0]); byname = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for name, sig, a, b in prof:
    t = a.elapsed_time(b) * 1e3
    agg[(name, sig)][0] += 1; agg[(name, sig)][1] += t
    byname[name][0] += 1; byname[name][1] += t; tot += t
print("image %.2f ms wall (eager, event-instrumented); sum of per-call device time %.2f ms over %d calls" % (e0.elapsed_time(e1), tot / 1e3, len(prof)))
for k, (n, t) in sorted(byname.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-34s n=%4d %9.1f us %5.1f%%" % (k, n, t, 100 * t / tot))
print()
for (name, sig), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[1]) if len(sys.argv) > 1 else 60]:
    print("%-22s %-50s n=%3d %8.1f us (%.1f each)" % (name.replace("s2r_", ""), sig, n, t, t / n))
