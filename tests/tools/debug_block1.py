import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
from emul import emulate_bf16
from oracle import ref_port as O
eng = sub("engine")
torch.manual_seed(1)
m = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
m._s2r_no_dropout = True
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
x = torch.randn(2, 3, 65, 97, generator=torch.Generator().manual_seed(0))
t = []
with emulate_bf16(t):
    O.deeplab_forward(sd, x, O.BNCfg(True), 16, drop=False)
m.cuda().train()
eng.TRACE = []
with torch.no_grad():
    m(x.cuda())
mine = {n: a.t[..., a.off:a.off + a.C].float().permute(0, 3, 1, 2).cpu() for n, a in eng.TRACE}
msd = m.state_dict()
for k in ['backbone.features.0.1', 'backbone.features.1.conv.1', 'backbone.features.1.conv.4', 'backbone.features.2.conv.1']:
    for s in ('running_mean', 'running_var'):
        a, b = msd[k + '.' + s].cpu().double(), sd[k + '.' + s].double()
        print(k, s, 'rel', float((a - b).norm() / b.norm()), 'maxabs', float((a - b).abs().max()))
e = dict(t)
d = (mine['block1'] - e['block1']).abs()
print('block1 diff: frac nonzero', float((d > 0).float().mean()), 'max', float(d.max()))
inner = d[:, :, 2:-2, 2:-2]
print(' interior frac nonzero', float((inner > 0).float().mean()), 'border frac', float((d > 0).float().sum() - (inner > 0).float().sum()) / (d.numel() - inner.numel()))
print(' per-channel nonzero frac', [round(float((d[:, c] > 0).float().mean()), 3) for c in range(16)])
