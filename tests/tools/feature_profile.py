"""Per-call device time of one eager feature-adaptation step (train.py:163-216, BASELINE config 4; CUDA events around
every C-ABI call, one stream), aggregated by (entry point, shape signature).  GPU box: python tests/tools/feature_profile.py [topN]"""
import collections, os, sys
os.environ["S2R_OVERLAP"] = "0"
os.environ["S2R_WGRAD_STREAM"] = "0"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
import bench
L = sub("_lib")
dev = torch.device("cuda", 0)
nn = torch.nn
torch.manual_seed(1)
BN = nn.BatchNorm2d
bb = sub("modeling.backbone.mobilenet").MobileNetV2(output_stride=16, BatchNorm=BN).to(dev).train()
aspp = sub("modeling.assp").ASPP('mobilenet', 16, BN).to(dev).train()
dec = sub("modeling.decoder").Decoder(19, 'mobilenet', BN).to(dev).train()
dc = sub("modeling.domian").DomainClassifer('mobilenet', BN).to(dev).train()
step = sub("steps").FeatureStep(bb, aspp, dec, dc, lr=5e-4, optimizer='Adam', epochs=1, iters_per_epoch=100)
src, lab, tgt = (t.to(dev) for t in bench.synth(1000, 8, 512, 1024))
for i in range(2):
    step(src, lab, tgt, i=i)
torch.cuda.synchronize()
L.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(src, lab, tgt, i=2); e1.record()
torch.cuda.synchronize()
prof, L.PROFILE = L.PROFILE, None
agg = collections.defaultdict(lambda: [0, 0.0]); byname = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for name, sig, a, b in prof:
    t = a.elapsed_time(b) * 1e3
    agg[(name, sig)][0] += 1; agg[(name, sig)][1] += t
    byname[name][0] += 1; byname[name][1] += t; tot += t
print("step %.1f ms wall (eager, event-instrumented); sum of per-call device time %.1f ms over %d calls" % (e0.elapsed_time(e1), tot / 1e3, len(prof)))
for k, (n, t) in sorted(byname.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%-34s n=%4d %9.1f us %5.1f%%" % (k, n, t, 100 * t / tot))
print()
for (name, sig), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[1]) if len(sys.argv) > 1 else 45]:
    print("%-22s %-50s n=%3d %8.1f us (%.1f each)" % (name.replace("s2r_", ""), sig, n, t, t / n))
