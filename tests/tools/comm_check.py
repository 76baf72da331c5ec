"""Correctness + latency of the NVLink peer-memory all-reduce (csrc/comm.cu) against NCCL.
    torchrun --nproc-per-node N tests/tools/comm_check.py"""
import ctypes as C, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
eng = sub("engine"); L = sub("_lib")
assert eng.init_peer_exchange(None), "peer exchange not available"
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for it in range(300):
    n = [2, 64, 1920, 4096, 37][it % 5]
    x = torch.randn(n, device=dev, dtype=torch.float64, generator=g)
    ref = x.clone(); dist.all_reduce(ref)
    L.call("s2r_allreduce_small_f64", C.c_void_p(x.data_ptr()), n, st)
    torch.cuda.synchronize()
    if not torch.allclose(x, ref, rtol=1e-12, atol=1e-12):
        ok = False; print("rank %d mismatch at it %d n %d: %g" % (rank, it, n, float((x - ref).abs().max())), flush=True); break
# the same through a CUDA graph (device-side sequence counter), 50 exchanges per replay
buf = torch.ones(1920, device=dev, dtype=torch.float64)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(50):
        L.call("s2r_allreduce_small_f64", C.c_void_p(buf.data_ptr()), 1920, C.c_void_p(torch.cuda.current_stream().cuda_stream))
dist.barrier(); torch.cuda.synchronize()
for rep in range(3):
    buf.fill_(1.0); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    t_peer = e0.elapsed_time(e1) / 50 * 1e3
    expect = float(world) ** 50
    if abs(float(buf[0]) / expect - 1) > 1e-9: ok = False; print("rank %d graph result %g != %g" % (rank, float(buf[0]), expect), flush=True)
# NCCL latency for comparison
buf2 = torch.ones(1920, device=dev, dtype=torch.float64)
for _ in range(5): dist.all_reduce(buf2)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): dist.all_reduce(buf2)
e1.record(); torch.cuda.synchronize()
t_nccl = e0.elapsed_time(e1) / 50 * 1e3
# two channels on two streams at once: every rank issues the same sequence PER CHANNEL, but interleaves the channels
# differently (rank-dependent delays), as the two streams of steps.AdaptStep do
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
a = torch.full((1920,), 1.0, device=dev, dtype=torch.float64)
b = torch.full((640,), 2.0, device=dev, dtype=torch.float64)
spin = torch.empty(1 << 22, device=dev)
torch.cuda.synchronize(); dist.barrier()
for it in range(40):
    with torch.cuda.stream(sA):
        if (it + rank) % 3 == 0: spin.normal_()          # skew the streams against each other
        L.call("s2r_allreduce_small_f64_ch", C.c_void_p(a.data_ptr()), 1920, 0, C.c_void_p(sA.cuda_stream))
        a.mul_(1.0 / world)
    with torch.cuda.stream(sB):
        if (it + 2 * rank) % 4 == 0: spin.normal_()
        L.call("s2r_allreduce_small_f64_ch", C.c_void_p(b.data_ptr()), 640, 1, C.c_void_p(sB.cuda_stream))
        b.mul_(1.0 / world)
torch.cuda.synchronize()
if abs(float(a[7]) - 1.0) > 1e-9 or abs(float(b[5]) - 2.0) > 1e-9 or abs(float(a.sum()) - 1920.0) > 1e-6:
    ok = False; print("rank %d two-channel result %g %g" % (rank, float(a[7]), float(b[5])), flush=True)
err = L.lib().s2r_comm_error()
flag = torch.tensor([1 if (ok and err == 0) else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("peer all-reduce of 1920 fp64: %.1f us per exchange (CUDA graph); NCCL eager: %.1f us; world %d; %s"
          % (t_peer, t_nccl, world, "OK" if int(flag.item()) == 1 else "FAILED"), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0 if int(flag.item()) == 1 else 1)
