"""world_size-2 checks of the data-parallel protocol on the CPU (gloo): what the ranks exchange and how it is
combined.  The product kernels need a GPU; here the per-rank pieces are restated in torch and run through real
`torch.distributed` collectives, against the reference algorithm evaluated on the gathered global batch:
  * synchronised BatchNorm forward: all-reduce of the fp64 per-channel (sum, sum of squares) + element count x world
    -> modeling/sync_batchnorm/batchnorm.py:113-125 (clamp(var, eps)^-1/2, unbiased running variance);
  * synchronised BatchNorm backward: all-reduce of (sum dy, sum dy*xhat) -> the autograd of the same formula;
  * gradient all-reduce: per-rank mean losses, gradients summed and divided by the world size (engine/optim.py);
  * data-parallel validation: utils.metrics.Evaluator.all_reduce (the product's method, on counts accumulated per rank
    with the oracle's bincount) -> the reference's Evaluator metrics on the whole set, bit-exact."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ref_port as O
        torch.manual_seed(0)
        Cc, eps, mom = 6, 1e-5, 0.1
        x_all = torch.randn(4, Cc, 5, 7, dtype=torch.float64) * 2 + 0.5
        dy_all = torch.randn(4, Cc, 5, 7, dtype=torch.float64)
        gamma = torch.rand(Cc, dtype=torch.float64) + 0.5
        beta = torch.randn(Cc, dtype=torch.float64)
        x, dy = x_all[2 * rank:2 * rank + 2], dy_all[2 * rank:2 * rank + 2]      # this rank's shard
        # ---- forward exchange
        sums = torch.stack([x.sum((0, 2, 3)), (x * x).sum((0, 2, 3))])
        dist.all_reduce(sums)
        count = x.numel() // Cc * world
        mean = sums[0] / count
        sumvar = sums[1] - sums[0] * mean
        invstd = (sumvar / count).clamp(eps) ** -0.5
        y = (x - mean.view(1, -1, 1, 1)) * (invstd * gamma).view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
        run_var = (1 - mom) * torch.ones(Cc, dtype=torch.float64) + mom * sumvar / (count - 1)
        # reference on the gathered batch
        sd = {'bn.weight': gamma.clone().requires_grad_(True), 'bn.bias': beta.clone().requires_grad_(True),
              'bn.running_mean': torch.zeros(Cc, dtype=torch.float64), 'bn.running_var': torch.ones(Cc, dtype=torch.float64)}
        xr = x_all.clone().requires_grad_(True)
        y_ref = O.batch_norm(sd, 'bn', xr, O.BNCfg(True, mom, eps, sync_clamp=True))
        assert torch.allclose(y, y_ref[2 * rank:2 * rank + 2].detach(), rtol=1e-10, atol=1e-10)
        assert torch.allclose(run_var, sd['bn.running_var'], rtol=1e-10)
        # ---- backward exchange
        xhat = (x - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
        bs = torch.stack([dy.sum((0, 2, 3)), (dy * xhat).sum((0, 2, 3))])
        dist.all_reduce(bs)
        dx = (gamma * invstd).view(1, -1, 1, 1) * (dy - (bs[0] / count).view(1, -1, 1, 1) - xhat * (bs[1] / count).view(1, -1, 1, 1))
        y_ref.backward(dy_all)
        assert torch.allclose(dx, xr.grad[2 * rank:2 * rank + 2], rtol=1e-8, atol=1e-10)
        # parameter gradients: local sums, completed by the gradient all-reduce (sum over ranks)
        dgamma_local = (dy * xhat).sum((0, 2, 3))
        dist.all_reduce(dgamma_local)
        assert torch.allclose(dgamma_local, sd['bn.weight'].grad, rtol=1e-8, atol=1e-10)
        # ---- gradient all-reduce of a mean loss: sum of per-rank gradients / world == gradient of the global mean
        w = torch.randn(Cc, dtype=torch.float64, requires_grad=True)
        (x.mean((0, 2, 3)) * w).sum().backward()
        g = w.grad.clone()
        dist.all_reduce(g)
        g /= world
        assert torch.allclose(g, x_all.mean((0, 2, 3)), rtol=1e-10)
        # ---- validation: per-rank confusion matrices summed by Evaluator.all_reduce == the matrix of the whole set
        import importlib
        import numpy as np
        metrics = importlib.import_module("synthetic-to-real-semantic-segmentation_b200.utils.metrics")
        rng = np.random.RandomState(11)
        gt_all = rng.randint(0, 20, (6, 33, 47)).astype(np.float32)
        gt_all[gt_all == 19] = 255
        pred_all = rng.randint(0, 19, (6, 33, 47)).astype(np.int64)
        ev = metrics.Evaluator(19)
        mine = slice(rank, None, world)                                   # rank r evaluates images r, r + world, ...
        ev._counts = torch.from_numpy(O.confusion_matrix(gt_all[mine], pred_all[mine], 19).astype(np.int64))
        ev._bad = torch.zeros(1, dtype=torch.int64)
        assert ev.all_reduce() is ev
        want = O.confusion_matrix(gt_all, pred_all, 19)
        assert np.array_equal(ev.confusion_matrix, want.astype(np.float64))
        m = O.evaluator_metrics(want)
        miou, iou = ev.Mean_Intersection_over_Union()
        assert miou == m['mIoU'] and np.array_equal(iou, m['IoU']) and ev.Pixel_Accuracy() == m['PA']
        assert ev.Pixel_Accuracy_Class() == m['mPA'] and ev.Frequency_Weighted_Intersection_over_Union() == m['fwIoU']
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        q.put((rank, "%s: %s" % (type(e).__name__, e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sync_bn_grad_allreduce_and_evaluator_protocol():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
