"""Evaluator with the reference's interface (utils/metrics.py:4-46); the confusion matrix is
accumulated on the device by the fused histogram kernels and read back only when a metric is asked
for.  Counts are exact integers (the reference accumulates them in float64)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ..engine import _vp


class Evaluator(object):
    def __init__(self, num_class, device=None):
        self.num_class = num_class
        self.device = torch.device(device) if device is not None else None
        self._counts = None
        self._bad = None

    def _ensure(self, device):
        if self._counts is None:
            L.require_cuda()
            self.device = device if device is not None else (self.device or torch.device("cuda", torch.cuda.current_device()))
            self._counts = torch.zeros((self.num_class, self.num_class), dtype=torch.int64, device=self.device)
            self._bad = torch.zeros(1, dtype=torch.int64, device=self.device)

    @property
    def confusion_matrix(self):
        if self._counts is None:
            return np.zeros((self.num_class,) * 2)
        if int(self._bad.item()) != 0:
            raise ValueError("prediction outside [0, num_class) for a valid label")
        return self._counts.cpu().numpy().astype(np.float64)

    @confusion_matrix.setter
    def confusion_matrix(self, value):
        self._ensure(None)
        self._counts.copy_(torch.as_tensor(np.asarray(value)).to(torch.int64))

    def Pixel_Accuracy(self):
        cm = self.confusion_matrix
        return np.diag(cm).sum() / cm.sum()

    def Pixel_Accuracy_Class(self):
        cm = self.confusion_matrix
        Acc = np.diag(cm) / cm.sum(axis=1)
        return np.nanmean(Acc)

    def Mean_Intersection_over_Union(self):
        cm = self.confusion_matrix
        IoU = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
        return np.nanmean(IoU), IoU

    def Frequency_Weighted_Intersection_over_Union(self):
        cm = self.confusion_matrix
        freq = np.sum(cm, axis=1) / np.sum(cm)
        iu = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
        return (freq[freq > 0] * iu[freq > 0]).sum()

    def _as_cuda(self, a, want_int):
        if isinstance(a, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(a))
        else:
            t = a
        self._ensure(t.device if t.device.type == "cuda" else None)
        t = t.to(self.device, non_blocking=True)
        if want_int:
            t = t.to(torch.int64)
        elif t.dtype not in (torch.float32, torch.int64):
            t = t.to(torch.int64) if not t.dtype.is_floating_point else t.float()
        return t.contiguous()

    def add_batch(self, gt_image, pre_image):
        """gt_image: labels (any numeric dtype, numpy or torch); pre_image: integer predictions."""
        assert gt_image.shape == pre_image.shape
        gt = self._as_cuda(gt_image, False)
        pred = self._as_cuda(pre_image, True)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            L.call("s2r_confusion_matrix", _vp(gt), 1 if gt.dtype == torch.int64 else 0, _vp(pred), gt.numel(),
                   self.num_class, _vp(self._counts), _vp(self._bad), st)

    def add_batch_logits(self, gt_image, logits):
        """Fused argmax over the class axis of NCHW fp32 logits + histogram (replaces the logits
        D2H copy and np.argmax at val_adapt.py:131-135).  Returns nothing; counts stay on device."""
        gt = self._as_cuda(gt_image, False)
        if gt.dtype != torch.float32:
            gt = gt.float()
        logits = logits.contiguous().float()
        N, Cc = logits.shape[0], logits.shape[1]
        HW = logits.numel() // (N * Cc)
        assert gt.numel() == N * HW
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            L.call("s2r_argmax_confusion_nchw", _vp(logits), _vp(gt), N, Cc, HW, self.num_class,
                   _vp(self._counts), None, st)

    def add_batch_lowres(self, gt_image, act_ptr, pitch, N, Hi, Wi, Cc, stream):
        """The final bilinear up-sampling of deeplab.py:31 to the label map's size, the argmax and the histogram in
        one launch, from the decoder's NHWC bf16 logits (raw device pointer; called by DeepLab.forward_confusion on
        the engine's stream).  Same counts as forward() + add_batch_logits(), bit for bit."""
        gt = self._as_cuda(gt_image, False)
        if gt.dtype != torch.float32:
            gt = gt.float()
        if gt.dim() != 3 or gt.shape[0] != N:
            raise ValueError("add_batch_lowres: label map %s does not match a batch of %d" % (tuple(gt.shape), N))
        L.call("s2r_upsample_argmax_confusion_nhwc", C.c_void_p(act_ptr), pitch, N, Hi, Wi, Cc, _vp(gt),
               gt.shape[1], gt.shape[2], self.num_class, _vp(self._counts), stream)

    def all_reduce(self, group=None):
        """Data-parallel validation (one process per GPU, every rank evaluates its shard of the images): ONE
        all-reduce(sum) of the int64 [num_class, num_class] counts -- and of the bad-prediction flag -- turns every
        rank's matrix into the matrix of the whole set.  Integer sums: exact and independent of the order, so the
        metrics equal the reference's single-process Evaluator (utils/metrics.py:34-46) fed all images.  Call once,
        after the last add_batch; a no-op in a single process."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
            return self
        self._ensure(None)
        dist.all_reduce(self._counts, group=group)
        dist.all_reduce(self._bad, group=group)
        return self

    def reset(self):
        if self._counts is not None:
            self._counts.zero_()
            self._bad.zero_()
