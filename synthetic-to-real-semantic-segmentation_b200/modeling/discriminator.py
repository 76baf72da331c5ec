"""FCDiscriminator with the reference's names and state_dict keys (modeling/discriminator.py:6-35):
five biased 4x4 stride-2 convolutions, LeakyReLU(0.2) between them, as tap-GEMMs over the four
input-parity views with bias + LeakyReLU fused in the epilogue; the backward fuses the LeakyReLU
mask into the next layer's data-gradient epilogue.
"""
import ctypes as C

import torch.nn as nn

import torch

from .. import _lib as L
from ..engine import (conv_fwd, conv_dgrad, conv_wgrad, bias_grad, conv_out_hw, round_up, im2col, patch_weight,
                      PadAct, rowtap_ok, rowtap_fwd, rowtap_wgrad, rowtap_dgrad, BF16, _vp)
from ..runtime import RunBase, call_module


def _global_softmax0():
    """functional.GLOBAL_SOFTMAX0 and more than one rank: the batch-axis softmax runs over the gathered batch."""
    from .. import functional as _fn
    if not _fn.GLOBAL_SOFTMAX0[0]:
        return False
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class FCDiscriminatorRun(RunBase):
    """conv1 reads its NCHW fp32 argument directly (raw_inputs): the argument -- optionally through the batch-axis
    softmax of train_adapt.py:151,166,174 (softmax0=True, the fused form of `model_D(F.softmax(x, dim=0))`) -- is
    written once as bf16 into a zero-padded NHWC buffer and conv1 runs as a 4-tap row GEMM on it (engine.PadAct).
    Inputs the row-tap form does not cover (odd sizes, > 64 channels) take the patch-matrix path."""
    raw_inputs = True

    def __init__(self, mod, softmax0=False):
        self.convs = [mod.conv1, mod.conv2, mod.conv3, mod.conv4, mod.classifier]
        self.slope = mod.leaky_relu.negative_slope
        self.softmax0 = softmax0

    def forward(self, cx, x):
        self.in_shape = (x.N, x.H, x.W, x.C)
        self.rowtap = rowtap_ok(x.C, x.H, x.W) and not (self.softmax0 and x.N > 8)
        if self.rowtap:
            Cp = round_up(x.C, 8)
            xp = PadAct(torch.empty((x.N, x.H + 2, x.W + 2, Cp), dtype=BF16, device=cx.device), x.H, x.W, x.C)
            self.gstat = None
            if self.softmax0 and _global_softmax0():
                # the softmax runs over the images of ALL ranks (functional.GLOBAL_SOFTMAX0): batch maximum and sum of
                # exponentials per (class, pixel), all-reduced, then the normalised values into the padded buffer
                import torch.distributed as dist
                M = x.C * x.H * x.W
                gmax = torch.empty(M, dtype=torch.float32, device=cx.device)
                gsum = torch.empty(M, dtype=torch.float32, device=cx.device)
                L.call("s2r_softmax0_batch_stats", _vp(x.t), x.N, M, None, _vp(gmax), cx.stream)
                dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
                L.call("s2r_softmax0_batch_stats", _vp(x.t), x.N, M, _vp(gmax), _vp(gsum), cx.stream)
                dist.all_reduce(gsum, op=dist.ReduceOp.SUM)
                L.call("s2r_softmax0_nchw_to_nhwc_pad_global", _vp(x.t), x.N, x.C, x.H, x.W, _vp(gmax), _vp(gsum),
                       C.c_void_p(xp.ptr), Cp, cx.stream)
                self.gstat = (gmax, gsum)
            else:
                L.call("s2r_softmax0_nchw_to_nhwc_pad", _vp(x.t), x.N, x.C, x.H, x.W, 1 if self.softmax0 else 0,
                       C.c_void_p(xp.ptr), Cp, cx.stream)
            self.logits = x.t if self.softmax0 else None
            acts = [xp]
        else:
            self.gstat = None
            if self.softmax0:
                raise NotImplementedError("fused softmax input needs even sizes, <= 64 channels and batch <= 8")
            acts = [im2col(cx, x, 4, 4, 2, 1)]   # conv1 as a pointwise GEMM over 4x4 patches
        h = x
        for i, c in enumerate(self.convs):
            last = i == len(self.convs) - 1
            Cout = c.weight.shape[0]
            OH, OW = conv_out_hw(h.H, h.W, 4, 4, 2, 1, 1)
            y = cx.new(h.N, OH, OW, round_up(Cout, 8), zero=Cout % 8 != 0)
            y.C = Cout
            if i == 0 and self.rowtap:
                rowtap_fwd(cx, acts[0], c.weight, y, bias=c.bias, act=L.ACT_LEAKY, slope=self.slope)
            elif i == 0:
                conv_fwd(cx, acts[0], patch_weight(c.weight), y, bias=c.bias, act=L.ACT_LEAKY, slope=self.slope)
            else:
                conv_fwd(cx, h, c.weight, y, stride=2, pad=1, bias=c.bias,
                         act=L.ACT_NONE if last else L.ACT_LEAKY, slope=self.slope)
            acts.append(y)
            h = y
        self.acts = acts
        return h

    def backward(self, cx, douts, need=None, wgrad=True, keep=False):
        """wgrad=False: data gradients only (the discriminator frozen); keep=True: the saved activations stay for a
        second backward pass over the same forward (FCDiscriminator.forward_softmax0_shared)."""
        d = douts[0] if isinstance(douts, tuple) else douts
        acts = self.acts
        if not keep:
            self.acts = None
        need_dx = True if need is None else bool(need[0])
        dsums = None    # per-channel sums of d left by the data-gradient launch that produced it (bias gradient)
        for i in range(len(self.convs) - 1, -1, -1):
            c = self.convs[i]
            xin = acts[i]
            if wgrad and c.weight.requires_grad:
                if i == 0 and self.rowtap:
                    rowtap_wgrad(cx, xin, d, c.weight)
                elif i == 0:
                    conv_wgrad(cx, xin, d, patch_weight(c.weight), grad_param=c.weight)
                else:
                    conv_wgrad(cx, xin, d, c.weight, stride=2, pad=1)
            if wgrad and c.bias is not None and c.bias.requires_grad:
                bias_grad(cx, d, c.bias, presummed=dsums)
            if i == 0 and not need_dx:
                return None
            if i == 0 and self.rowtap:
                # gradient w.r.t. the padded bf16 input, then (softmax backward +) NHWC -> NCHW fp32 in one pass
                N, H, W, Cc = self.in_shape
                dxp = rowtap_dgrad(cx, d, c.weight, H, W)
                dx = torch.empty((N, Cc, H, W), dtype=torch.float32, device=cx.device)
                if self.softmax0 and self.gstat is not None:
                    import torch.distributed as dist
                    gmax, gsum = self.gstat
                    tsum = torch.empty(Cc * H * W, dtype=torch.float32, device=cx.device)
                    L.call("s2r_softmax0_nhwc_pad_bwd_global", _vp(self.logits), C.c_void_p(dxp.ptr), N, Cc, H, W, dxp.Cp,
                           _vp(gmax), _vp(gsum), _vp(tsum), None, None, cx.stream)
                    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
                    L.call("s2r_softmax0_nhwc_pad_bwd_global", _vp(self.logits), C.c_void_p(dxp.ptr), N, Cc, H, W, dxp.Cp,
                           _vp(gmax), _vp(gsum), None, _vp(tsum), _vp(dx), cx.stream)
                else:
                    L.call("s2r_softmax0_nhwc_pad_bwd", _vp(self.logits), C.c_void_p(dxp.ptr), N, Cc, H, W, dxp.Cp,
                           1 if self.softmax0 else 0, _vp(dx), cx.stream)
                if not keep:
                    self.logits = None
                    self.gstat = None
                return dx
            if i == 0:
                N, H, W, Cc = self.in_shape   # gradient w.r.t. the input pixels: the ordinary 4x4 data gradient
                dx = cx.new(N, H, W, round_up(Cc, 8))
                dx.C = Cc
            else:
                dx = cx.new(xin.N, xin.H, xin.W, xin.pitch)
                dx.C = xin.C
            # dz_{i-1} = dgrad * leaky'(y_{i-1}); the first layer's input has no activation.  When the layer below
            # needs its bias gradient, the epilogue of this launch leaves the per-channel sums of dz_{i-1}
            below = self.convs[i - 1] if i > 0 else None
            want = wgrad and below is not None and below.bias is not None and below.bias.requires_grad
            dsums = cx.f64(2 * dx.C) if (want and dx.C <= 1024 and dx.C % 8 == 0) else None
            conv_dgrad(cx, d, c.weight, dx, stride=2, pad=1, aux=xin if i > 0 else None,
                       aux_mode=L.AUX_LEAKY_MASK if i > 0 else L.AUX_NONE, slope=self.slope, stats=dsums)
            d = dx
        return d


class _SharedForward(object):
    """One evaluation of the discriminator whose saved activations serve two backward passes."""

    def __init__(self):
        self.run, self.out, self.pending = None, None, 2

    def done(self):
        self.pending -= 1
        if self.pending <= 0 and self.run is not None:
            self.run.acts = None
            self.run.logits = None
            self.run.gstat = None
            self.run = None


class _SharedInputFn(torch.autograd.Function):
    """out = D(softmax0(logits)) with the discriminator FROZEN: the gradient goes to the logits only."""

    @staticmethod
    def forward(ctx, module, holder, logits):
        from ..engine import Ctx, RawNCHW
        from ..runtime import to_nchw
        dev = logits.device
        if dev.type != "cuda":
            raise L.S2RError("s2r_b200 modules run on CUDA tensors only (got %s); there is no CPU path" % dev)
        with torch.cuda.device(dev):
            cx = Ctx(dev, module.training, None)
            run = FCDiscriminatorRun(module, softmax0=True)
            out = run.export(cx, 0, run.forward(cx, RawNCHW(logits)))
        # (a detached alias: `out` itself gets this node as its grad_fn, and holder -> out -> grad_fn -> ctx -> holder
        # would be a reference cycle that only the garbage collector frees)
        holder.run, holder.out = run, out.detach()
        ctx.holder, ctx.module, ctx.dev = holder, module, dev
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        from ..engine import Ctx
        h = ctx.holder
        res = None
        if dout is not None and ctx.needs_input_grad[2]:
            with torch.cuda.device(ctx.dev):
                cx = Ctx(ctx.dev, True, None, dropout=False, async_wgrad=True)
                res = h.run.backward(cx, (h.run.import_grad(cx, 0, dout),), (True,), wgrad=False, keep=True)
                cx.join()
        h.done()
        return None, None, res


class _SharedParamFn(torch.autograd.Function):
    """The same output values as a function of the discriminator's PARAMETERS only (the input detached): the backward
    pass accumulates the parameter gradients from the activations the shared forward saved."""

    @staticmethod
    def forward(ctx, module, holder, *params):
        ctx.holder, ctx.module, ctx.dev = holder, module, holder.out.device
        ctx.set_materialize_grads(False)
        return holder.out.detach().clone()

    @staticmethod
    def backward(ctx, dout):
        from ..engine import Ctx
        h = ctx.holder
        if dout is not None:
            with torch.cuda.device(ctx.dev):
                cx = Ctx(ctx.dev, True, None, dropout=False, async_wgrad=True)
                h.run.backward(cx, (h.run.import_grad(cx, 0, dout),), (False,), wgrad=True, keep=True)
                cx.join()
        h.done()
        return (None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


class FCDiscriminator(nn.Module):
    def __init__(self, num_classes, ndf=64):
        super().__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(ndf, ndf * 2, kernel_size=4, stride=2, padding=1)
        self.conv3 = nn.Conv2d(ndf * 2, ndf * 4, kernel_size=4, stride=2, padding=1)
        self.conv4 = nn.Conv2d(ndf * 4, ndf * 8, kernel_size=4, stride=2, padding=1)
        self.classifier = nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        return call_module(self, lambda: FCDiscriminatorRun(self), (x,))

    def forward_softmax0(self, logits):
        """`self(F.softmax(logits, dim=0))` -- the call train_adapt.py:151,166,174 makes -- with the batch-axis
        softmax (and its backward) fused into the input stage: the fp32 softmax tensor is never materialised."""
        return call_module(self, lambda: FCDiscriminatorRun(self, softmax0=True), (logits,))

    def forward_softmax0_shared(self, logits):
        """The adversarial pass and the discriminator's training pass on the target prediction evaluate
        `self(F.softmax(x, dim=0))` on the SAME tensor with the SAME weights (train_adapt.py:151 and :174; the
        discriminator is only updated at :181): one evaluation serves both.  Returns (frozen, attach):
          frozen  -- the output as a function of `logits` with the discriminator frozen (train_adapt.py:140-141,151-155:
                     its backward is the data-gradient chain down to the logits, no parameter gradients);
          attach  -- a callable returning the same values as a function of the discriminator's parameters with the
                     input detached (train_adapt.py:159,173-178: its backward accumulates the parameter gradients).
                     Call it on the stream the training pass should run on.
        Values and gradients are those of the two separate calls (the forward kernels are deterministic)."""
        holder = _SharedForward()
        frozen = _SharedInputFn.apply(self, holder, logits)
        params = [p for p in self.parameters()]
        return frozen, (lambda: _SharedParamFn.apply(self, holder, *params))
