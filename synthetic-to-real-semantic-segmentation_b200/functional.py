"""Loss-side autograd functions on NCHW fp32 tensors (the layout the reference's scripts hold
their logits in): batch-axis softmax, ignore-index cross entropy, BCE-with-logits.  Each is one
autograd node whose forward and backward are single C-ABI kernel launches."""
import ctypes as C

import os

import torch

from . import _lib as L
from .engine import _vp, dp_world, allreduce_small_f64


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check_cuda(t, who):
    if t.device.type != "cuda":
        raise L.S2RError("%s runs on CUDA tensors only (got %s); there is no CPU path" % (who, t.device))


class _SoftmaxDim0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _check_cuda(x, "softmax_dim0")
        x = x.contiguous().float()
        y = torch.empty_like(x)
        B = x.shape[0]
        M = x.numel() // B
        with torch.cuda.device(x.device):
            L.call("s2r_softmax_dim0_fwd", _vp(x), _vp(y), B, M, _stream(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        B = y.shape[0]
        with torch.cuda.device(y.device):
            L.call("s2r_softmax_dim0_bwd", _vp(y), _vp(dy), _vp(dx), B, y.numel() // B, _stream(y))
        return dx


def softmax_dim0(x):
    """F.softmax(x, dim=0) -- the batch-axis softmax of train_adapt.py:151,166,174."""
    return _SoftmaxDim0.apply(x)


# Deferred scaling of the cross-entropy gradient.  The kernel writes the UNSCALED gradient in its forward pass; the mean
# reduction's factor gout / sum(weights) is normally applied by one more pass over the tensor (the largest of the step:
# N x 19 x H x W fp32).  Inside the captured training steps, where the only consumer of that gradient is the DeepLab
# node's up-sampling backward, the factor is handed over as a device scalar instead: the backward registers it under the
# gradient's address and DeepLabRun.import_grad folds it into its kernel (s2r_upsample_bilinear_nchw_bwd_to_nhwc_scaled).
# Off by default: any other consumer (user code, hooks) sees the fully scaled gradient.
DEFER_CE_SCALE = [False]
PENDING_SCALE = {}
# The same, decided when the loss is EVALUATED: a cross entropy computed while DEFER_NEXT is set defers its scale in
# whatever backward pass it later takes part in (steps.FeatureStep sums the task loss with the domain losses and
# calls backward once: only the task loss feeds an up-sampling backward kernel that can take the factor).
DEFER_NEXT = [False]


def pop_pending_scale(t):
    return PENDING_SCALE.pop(t.data_ptr(), None) if PENDING_SCALE else None


# Data-parallel runs (one process per GPU, torch.distributed initialised with more than one rank): the reference
# evaluates its criterion on the batch GATHERED by nn.DataParallel (train_adapt.py:87-88,144-145), i.e. the mean runs
# over the valid pixels of the global batch.  With this switch on (default) the cross entropy all-reduces its three
# sums -- 24 bytes -- so that every rank returns that global mean, and scales its local gradient by world / sum_global(w):
# after the gradient all-reduce + 1/world of the fused optimizers this is exactly the gradient of the global-mean loss,
# also when the ranks hold different numbers of valid (non-ignored) pixels.  Constant-target calls without class
# weights (the domain loss, utils/loss.py:57-69) have equal counts on every rank and skip the exchange.
GLOBAL_BATCH_MEAN = [True]

# F.softmax(x, dim=0) in front of the discriminator (train_adapt.py:151,166,174) over the GLOBAL batch: the reference's
# single-process nn.DataParallel gathers the logits of all replicas on one device first (train_adapt.py:87-88), so its
# softmax runs over the images of every GPU.  Off by default -- the north star lists the BatchNorm sums and the gradients
# as the only exchanges, and this one costs two all-reduces of a [19,H,W] fp32 map per discriminator evaluation plus one
# per backward pass (5 x 40 MB per adaptation step at 512x1024); S2R_GLOBAL_SOFTMAX0=1 (or setting this flag) selects
# the exact semantics: FCDiscriminator.forward_softmax0* then exchange the batch maximum, the sum of exponentials and,
# in the backward pass, the sum of g*y over torch.distributed.
GLOBAL_SOFTMAX0 = [os.environ.get("S2R_GLOBAL_SOFTMAX0", "0") == "1"]


class _CrossEntropy(torch.autograd.Function):
    """mean_{valid}(w_t * (lse - x_t)) with ignore_index (nn.CrossEntropyLoss, reduction='mean')."""

    @staticmethod
    def forward(ctx, logit, target, const_target, weight, ignore_index, stats_out):
        _check_cuda(logit, "cross_entropy")
        logit = logit.contiguous().float()
        N, Cc = logit.shape[0], logit.shape[1]
        HW = logit.numel() // (N * Cc)
        if target is not None:
            target = target.contiguous()
            if target.dtype != torch.float32:
                target = target.float()
            if target.numel() != N * HW:
                raise ValueError("Expected target size %s, got %s" % ((N,) + tuple(logit.shape[2:]), tuple(target.shape)))
        need_grad = ctx.needs_input_grad[0]
        sums = torch.zeros(3, dtype=torch.float64, device=logit.device)
        grad = torch.empty_like(logit) if need_grad else None
        out = torch.empty((), dtype=torch.float32, device=logit.device)
        with torch.cuda.device(logit.device):
            st = _stream(logit)
            L.call("s2r_cross_entropy_nchw", _vp(logit), _vp(target), int(const_target), _vp(weight), N, Cc, HW,
                   int(ignore_index), _vp(sums), _vp(grad), st)
            world = dp_world() if (GLOBAL_BATCH_MEAN[0] and (target is not None or weight is not None)) else 1
            if world > 1:
                allreduce_small_f64(sums)
            L.call("s2r_ratio", _vp(sums), 0.0, _vp(out), st)
        ctx.grad = grad
        ctx.sums = sums
        ctx.world = world
        ctx.defer = bool(DEFER_NEXT[0])
        if stats_out is not None:
            stats_out.append(sums)
        return out

    @staticmethod
    def backward(ctx, gout):
        grad, sums = ctx.grad, ctx.sums
        ctx.grad = None
        gout = gout.contiguous().float()
        if ctx.world > 1:
            gout = gout * float(ctx.world)      # see GLOBAL_BATCH_MEAN
        if DEFER_CE_SCALE[0] or ctx.defer:
            PENDING_SCALE.clear()          # at most one pending gradient: a stale entry must never meet a recycled address
            PENDING_SCALE[grad.data_ptr()] = (gout.double() / sums[1]).float().reshape(1)
            return grad, None, None, None, None, None
        with torch.cuda.device(grad.device):
            L.call("s2r_scale_by_ratio", _vp(grad), grad.numel(), _vp(gout), _vp(sums), 0.0, _stream(grad))
        return grad, None, None, None, None, None


def cross_entropy(logit, target=None, const_target=0, weight=None, ignore_index=255, stats_out=None):
    if weight is not None:
        weight = weight.to(device=logit.device, dtype=torch.float32).contiguous()
    return _CrossEntropy.apply(logit, target, const_target, weight, ignore_index, stats_out)


class _BCEWithLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, const_target):
        _check_cuda(x, "bce_with_logits")
        x = x.contiguous().float()
        if target is not None:
            target = target.contiguous().float()
            if target.shape != x.shape:
                raise ValueError("Target size ({}) must be the same as input size ({})".format(target.shape, x.shape))
        n = x.numel()
        sums = torch.zeros(2, dtype=torch.float64, device=x.device)
        out = torch.empty((), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            st = _stream(x)
            L.call("s2r_bce_logits_fwd", _vp(x), _vp(target), float(const_target), n, _vp(sums), st)
            L.call("s2r_ratio", _vp(sums), float(n), _vp(out), st)
        ctx.save_for_backward(x, target)
        ctx.const_target = float(const_target)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, target = ctx.saved_tensors
        dx = torch.empty_like(x)
        gout = gout.contiguous().float()
        with torch.cuda.device(x.device):
            L.call("s2r_bce_logits_bwd", _vp(x), _vp(target), ctx.const_target, x.numel(), _vp(gout), _vp(dx),
                   _stream(x))
        return dx, None, None


def bce_with_logits(x, target):
    """torch.nn.BCEWithLogitsLoss()(x, target); target may be a python scalar (constant map)."""
    if isinstance(target, (int, float)):
        return _BCEWithLogits.apply(x, None, float(target))
    return _BCEWithLogits.apply(x, target, 0.0)
