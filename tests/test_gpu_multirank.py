"""The PRODUCT's data-parallel path with two ranks, on the GPU box the driver's `-m gpu` run uses (one B200 is
enough): two processes -- on two GPUs when the box has them, otherwise time-sliced on the same one -- meet through
torch.distributed (gloo: plumbing only) and run

  * the synchronised-BatchNorm exchange -- csrc/comm.cu (CUDA-IPC mapped peer memory) when every rank has its own
    GPU, the torch.distributed collective path of the same engine code on a single GPU (see _worker) -- inside
    engine.bn_plan / engine.bn_backward, against
    modeling/sync_batchnorm/batchnorm.py:90-125 of the reference evaluated by the oracle on the GATHERED batch
    (O.batch_norm(..., sync_clamp=True): clamp(var, eps)^-1/2, unbiased running variance, global element count);
  * the global-batch mean of the cross entropy (functional.GLOBAL_BATCH_MEAN: train_adapt.py:87-88,144-145 evaluates
    the criterion on the gathered batch) with DIFFERENT numbers of valid pixels per rank;
  * the gradient all-reduce + fused SGD step of optim.py, after which both ranks must hold bit-identical weights and
    BatchNorm buffers.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu
PKG = "synthetic-to-real-semantic-segmentation_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_batch():
    g = torch.Generator().manual_seed(42)
    x = torch.randn(4, 3, 64, 96, generator=g)
    lab = torch.randint(0, 19, (4, 64, 96), generator=g).float()
    lab[:2][torch.rand(2, 64, 96, generator=g) < 0.05] = 255       # rank 0: ~5 % ignored
    lab[2:][torch.rand(2, 64, 96, generator=g) < 0.50] = 255       # rank 1: ~50 % ignored -> unequal valid counts
    return x, lab


def _worker(rank, world, port, q):
    import importlib
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    # Two GPUs: the NVLink peer-memory exchange of csrc/comm.cu (the product path).  ONE GPU: the two ranks are
    # time-sliced on it, and kernels that spin on a flag the OTHER process writes must not be launched there (nothing
    # guarantees that both run at the same time; the B200 profiling notes report context-switch timeouts, Xid 109, for
    # exactly this) -- the statistics then travel through torch.distributed (S2R_COMM=nccl selects the collective
    # path; the process group here is gloo), which waits on the host, never on the device.
    if torch.cuda.device_count() >= world:
        os.environ.pop("S2R_COMM", None)
    else:
        os.environ["S2R_COMM"] = "nccl"
    res = {}
    try:
        dev = torch.device("cuda", rank % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sub = lambda n: importlib.import_module(PKG + "." + n)      # noqa: E731
        eng = sub("engine")
        torch.manual_seed(1)
        G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=True)
        G._s2r_no_dropout = True
        sd0 = {k: v.detach().clone() for k, v in G.state_dict().items()}
        G.to(dev).train()
        x, lab = _global_batch()
        per = x.shape[0] // world
        xl, ll = x[rank * per:(rank + 1) * per].to(dev), lab[rank * per:(rank + 1) * per].to(dev)
        opt = sub("optim").FusedSGD([{'params': list(G.get_1x_lr_params()), 'lr': 5e-4},
                                     {'params': list(G.get_10x_lr_params()), 'lr': 5e-3}], lr=5e-4, momentum=0.9,
                                    weight_decay=5e-4)
        out = G(xl)
        if os.environ.get("S2R_COMM") != "nccl":
            assert eng.PEER["world"] == world, "the NVLink/IPC peer exchange was not set up: %r" % (eng.PEER,)
        res["exchange"] = "peer memory (csrc/comm.cu)" if eng.PEER["world"] == world else "torch.distributed collectives (one GPU)"
        loss = sub("utils.loss").SegmentationLosses().build_loss('ce')(out, ll)
        loss.backward()
        opt.all_reduce_grads()
        cls_grad = (G.decoder.last_conv[8].weight.grad.detach() * opt.grad_scale).double().cpu()
        bufs = {k: v.detach().double().cpu() for k, v in G.state_dict().items() if 'running_' in k and '_level_features.' not in k}
        opt.step()
        torch.cuda.synchronize(dev)
        assert sub("_lib").lib().s2r_comm_error() == 0
        # bit-equality across the ranks: every parameter and buffer after the step
        flat = torch.cat([v.detach().reshape(-1).float() for v in G.state_dict().values() if v.dtype.is_floating_point])
        digest = (int(flat.view(torch.int32).to(torch.int64).sum().item()), float(flat.double().abs().sum().item()))
        digests = [None] * world
        dist.all_gather_object(digests, digest)
        res["digests"] = digests
        res["loss"] = float(loss.item())
        if rank == 0:
            from oracle import ref_port as O
            from emul import emulate_bf16

            def oracle(emulate):
                sd = {k: v.clone() for k, v in sd0.items()}
                for v in O.leaf_params(sd).values():
                    v.requires_grad_(True)
                cfg = O.BNCfg(True, sync_clamp=True)
                if emulate:
                    with emulate_bf16():
                        o = O.seg_cross_entropy(O.deeplab_forward(sd, x, cfg, 16, drop=False), lab)
                        o.backward()
                else:
                    o = O.seg_cross_entropy(O.deeplab_forward(sd, x, cfg, 16, drop=False), lab)
                    o.backward()
                return sd, float(o)

            sd32, loss32 = oracle(False)
            sde, losse = oracle(True)
            res["oracle_loss"] = (loss32, losse)
            errs = {}
            for k, v in bufs.items():
                if k.endswith('running_var'):
                    errs[k] = float((v - sde[k].double()).norm() / ((sde[k].double() - 0.9).norm() + 1e-30))
            res["bn_var_err_emul"] = errs
            early = ('backbone.features.0.1.running_var', 'backbone.features.1.conv.1.running_var',
                     'backbone.features.2.conv.1.running_var')
            res["bn_early_vs_fp32"] = {k: float((bufs[k] - sd32[k].double()).norm() / ((sd32[k].double() - 0.9).norm() + 1e-30))
                                       for k in early}
            res["bn_mean_early_vs_fp32"] = float((bufs['backbone.features.0.1.running_mean'] -
                                                  sd32['backbone.features.0.1.running_mean'].double()).norm() /
                                                 sd32['backbone.features.0.1.running_mean'].double().norm())
            g32 = sd32['decoder.last_conv.8.weight'].grad.double()
            res["cls_grad_err"] = float((cls_grad - g32).norm() / g32.norm())
        q.put((rank, "ok", res))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "%s: %s\n%s" % (type(e).__name__, e, traceback.format_exc()), res))
    finally:
        try:
            importlib.import_module(PKG + "._lib").lib().s2r_comm_destroy()
            dist.destroy_process_group()
        except Exception:   # noqa: BLE001
            pass


def test_two_ranks_sync_bn_global_ce_and_identical_weights(built_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = [q.get(timeout=420) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    res = {r: (status, info) for r, status, info in got}
    assert res[0][0] == "ok" and res[1][0] == "ok", res
    r0 = res[0][1]
    print("BN-statistics exchange:", r0.get("exchange"))
    print("two ranks on %d GPU(s): loss %.5f (oracle fp32 %.5f, bf16-emulated %.5f); classifier gradient err %.4f; "
          "early BN vs fp32 %s mean %.2e" % (min(2, torch.cuda.device_count()), r0["loss"], r0["oracle_loss"][0],
                                             r0["oracle_loss"][1], r0["cls_grad_err"], r0["bn_early_vs_fp32"],
                                             r0["bn_mean_early_vs_fp32"]))
    # identical weights and buffers on both ranks, bit for bit
    assert r0["digests"][0] == r0["digests"][1] == res[1][1]["digests"][0], r0["digests"]
    # both ranks return the GLOBAL mean cross entropy (unequal valid counts: a mean of per-rank means would differ)
    assert res[0][1]["loss"] == res[1][1]["loss"]
    assert abs(r0["loss"] - r0["oracle_loss"][0]) <= 1e-2 * r0["oracle_loss"][0]
    # synchronised statistics: the first layers see (almost) the reference's activations -> tight against fp32
    assert max(r0["bn_early_vs_fp32"].values()) <= 1e-2 and r0["bn_mean_early_vs_fp32"] <= 1e-2, r0
    errs = sorted(r0["bn_var_err_emul"].values())
    print("BN running_var vs bf16-emulated global-batch oracle: median %.4f max %.4f over %d layers" % (errs[len(errs) // 2], errs[-1], len(errs)))
    assert len(errs) == 60 and errs[len(errs) // 2] <= 2e-2 and errs[-1] <= 2.5e-1, errs[-5:]
    # gradient of the global-mean loss (all-reduced and divided by the world size) at the classifier
    assert r0["cls_grad_err"] <= 0.2, r0["cls_grad_err"]
