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


def patch_replication_callback(data_parallel):
    """No-op kept for call-site compatibility (train_adapt.py:89-90): there are no in-process
    replicas to tag, ranks meet in the NCCL all-reduce issued by the engine."""
    return data_parallel


class DataParallelWithCallback(nn.DataParallel):
    pass


def convert_model(module):
    return module
