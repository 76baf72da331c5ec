"""How well does the generator's forward pass (a chain of ~110 mostly small kernels) overlap with the discriminator's
two training passes (a chain of large GEMMs) when both are captured in one CUDA graph on two streams?  The answer
bounds what software pipelining across steps (D training of iteration k beside G forward of iteration k+1) can buy.
GPU box:  python tests/tools/overlap_probe.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(1)
G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False).to(dev).train()
D = sub("modeling.discriminator").FCDiscriminator(num_classes=19).to(dev).train()
st = sub("steps")
fn = sub("functional")
src, lab, tgt = (t.to(dev) for t in bench.synth(1000, 8, 512, 1024))
crit = sub("utils.loss").SegmentationLosses().build_loss('ce')
logits_a = torch.randn(8, 19, 512, 1024, device=dev)
logits_b = torch.randn(8, 19, 512, 1024, device=dev)


def g_forward():
    out = G(src)
    return crit(out, lab)


def g_fwd_bwd():
    loss = g_forward()
    loss.backward()


def d_train():
    for lg, t in ((logits_a, 0), (logits_b, 1)):
        l = fn.bce_with_logits(st._disc_on_softmax0(D, lg), t)
        l.backward()


def timed_graph(body, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body(); body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


side = torch.cuda.Stream()


def both(first):
    def body():
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            d_train()
        first()
        cur.wait_stream(side)
    return body


for p in list(G.parameters()) + list(D.parameters()):
    p.grad = None
t_f = timed_graph(g_forward)
t_fb = timed_graph(g_fwd_bwd)
t_d = timed_graph(d_train)
t_fd = timed_graph(both(g_forward))
t_fbd = timed_graph(both(g_fwd_bwd))
print("G forward + CE alone            %.2f ms" % t_f)
print("G forward + CE + backward alone %.2f ms" % t_fb)
print("D training passes (src, tgt)    %.2f ms" % t_d)
print("G forward || D training         %.2f ms  (sum %.2f, max %.2f)" % (t_fd, t_f + t_d, max(t_f, t_d)))
print("G forward+backward || D training %.2f ms  (sum %.2f, max %.2f)" % (t_fbd, t_fb + t_d, max(t_fb, t_d)))
