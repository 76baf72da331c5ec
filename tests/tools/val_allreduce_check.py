"""Data-parallel validation on N GPUs: every rank runs ValStep on its shard of a fixed set of synthetic images, then
Evaluator.all_reduce sums the int64 confusion matrices over NCCL; rank 0 also evaluates the whole set alone and the two
matrices must be identical.     torchrun --nproc-per-node N tests/tools/val_allreduce_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()

torch.manual_seed(1)
model = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False).to(dev).eval()
steps = sub("steps")
n_images = 4 * world + 1                                   # ragged: the shards differ in size


def sample(k):
    g = torch.Generator().manual_seed(700 + k)
    x = torch.randn(1, 3, 256, 512, generator=g)
    lab = torch.randint(0, 20, (1, 256, 512), generator=g).float()
    lab[lab == 19] = 255
    return x.to(dev), lab.to(dev)


val = steps.ValStep(model, 19)
for k in range(rank, n_images, world):                     # rank r takes images r, r + world, ...
    val(*sample(k))
cm = val.all_reduce().confusion_matrix
if rank == 0:
    alone = steps.ValStep(model, 19)
    for k in range(n_images):
        alone(*sample(k))
    want = alone.evaluator.confusion_matrix
    same = bool(np.array_equal(cm, want))
    print("validation over %d ranks: %d images, %d valid pixels; all-reduced matrix identical to the single-process one: %s; mIoU %.6f"
          % (world, n_images, int(cm.sum()), same, val.evaluator.Mean_Intersection_over_Union()[0]), flush=True)
    assert same
dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()
