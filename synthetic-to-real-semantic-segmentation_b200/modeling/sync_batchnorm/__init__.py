"""Synchronized BatchNorm classes with the reference's names
(modeling/sync_batchnorm/__init__.py:11-12, batchnorm.py:39-281, replicate.py:65-88).

The reference synchronises replicas of a single-process nn.DataParallel through Python
queues (comm.py).  Here there is one process per GPU: the classes are parameter containers with
the `_s2r_sync` marker; the engine all-reduces the per-channel sum / sum-of-squares (forward) and
sum(dy) / sum(dy*xhat) (backward) over NCCL when torch.distributed is initialised with more
than one rank, and then follows batchnorm.py:113-125 (clamp(var, eps)^-1/2, unbiased running var).
"""
import torch.nn as nn


class SynchronizedBatchNorm1d(nn.BatchNorm1d):
    _s2r_sync = True


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    _s2r_sync = True


class SynchronizedBatchNorm3d(nn.BatchNorm3d):
    _s2r_sync = True


_ONE_PROCESS_PER_GPU = ("this path runs ONE PROCESS PER GPU (torchrun --nproc-per-node N; torch.distributed/NCCL), not a "
                        "single-process nn.DataParallel over %d devices: replicas made by DataParallel.replicate would "
                        "share the per-call activation records of the C-ABI engine.  Wrap with device_ids=[local_rank] "
                        "(a pass-through, as the reference's scripts do on one GPU) or do not wrap at all")


def _check_single_device(data_parallel):
    ids = getattr(data_parallel, "device_ids", None) or []
    if len(ids) > 1:
        raise RuntimeError(_ONE_PROCESS_PER_GPU % len(ids))
    return data_parallel


def patch_replication_callback(data_parallel):
    """Kept for call-site compatibility (train_adapt.py:87-90: DataParallel(model, device_ids) followed by
    patch_replication_callback(model)).  With one device id nn.DataParallel calls the wrapped module directly and there
    is nothing to patch: ranks meet in the statistics exchange issued by the engine.  More than one device id in one
    process is refused loudly instead of silently training unsynchronised replicas."""
    return _check_single_device(data_parallel)


class DataParallelWithCallback(nn.DataParallel):
    """modeling/sync_batchnorm/replicate.py:47-62; same single-device rule as patch_replication_callback."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super().__init__(module, device_ids=device_ids, output_device=output_device, dim=dim)
        _check_single_device(self)


def convert_model(module):
    """The reference's helper turns SynchronizedBatchNorm modules back into nn.BatchNorm for single-GPU use; here the
    same classes serve both cases (the engine synchronises only when torch.distributed has more than one rank), so
    the module is returned as it is."""
    return module
