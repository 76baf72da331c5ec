"""Device time of the discriminator's input stage on the step's shape (B = 8, 19 classes, 512x1024): batch-axis softmax
of the fp32 NCHW logits into the zero-padded NHWC bf16 buffer, and its backward (CUDA-graph replay of 5 launches; the
640 MB working set exceeds the L2).  GPU box: python tests/tools/softmax0_bench.py"""
import sys, os, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from conftest import sub
L = sub("_lib")
dev = torch.device("cuda", 0)
B, Cc, H, W, Cp = 8, 19, 512, 1024, 24
x = torch.randn(B, Cc, H, W, device=dev)
xp = torch.empty(B, H + 2, W + 2, Cp, dtype=torch.bfloat16, device=dev)
gp = torch.randn(B, H + 2, W + 2, Cp, device=dev).to(torch.bfloat16)
dx = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
vp = lambda t: C.c_void_p(t.data_ptr())
ops = {
    "softmax0_nchw_to_nhwc_pad": (lambda s: L.call("s2r_softmax0_nchw_to_nhwc_pad", vp(x), B, Cc, H, W, 1, vp(xp), Cp, s), x.numel() * 4 + xp.numel() * 2),
    "softmax0_nhwc_pad_bwd": (lambda s: L.call("s2r_softmax0_nhwc_pad_bwd", vp(x), vp(gp), B, Cc, H, W, Cp, 1, vp(dx), s), x.numel() * 8 + gp.numel() * 2),
}
for name, (fn, nbytes) in ops.items():
    fn(st); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(5): fn(s)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5 * 1e3)
    print("%-28s %7.1f us  %6.0f GB/s  (%.0f MB)" % (name, best, nbytes / best / 1e3, nbytes / 1e6))
