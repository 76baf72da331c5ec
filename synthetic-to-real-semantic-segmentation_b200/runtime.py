"""The torch.autograd boundary of the drop-in modules.

A reference-style call `y = module(x)` creates ONE autograd node.  Its forward converts the
NCHW fp32 arguments to NHWC bf16 views, builds a fresh "run" object (the per-call record of saved
activations, so a module can be called several times before backward, as train.py:182-196 does)
and executes it on the C-ABI kernels; its backward executes the run's hand-written backward and
accumulates parameter gradients straight into `.grad`.
"""

import os

import torch

from . import _lib as L
from .engine import Act, Ctx, RawNCHW, round_up, _vp, bn_eval_all


def init_reference_weights(module):
    """The initialisation every constructor of the reference ends with (mobilenet.py:134-145, assp.py:22-32,80-91,
    decoder.py:45-54, domian.py:35-44): kaiming-normal Conv2d weights, BatchNorm weights one and biases zero, visited
    in `module.modules()` order -- the same order, hence the same draws from the random stream, as the reference;
    biases of Conv2d keep nn.Conv2d's own initialisation."""
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight)
        elif isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


def sync_group_for(module):
    """The process group BN statistics are reduced over, or None (single process / plain BN)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    if not getattr(module, "_s2r_has_sync_bn", False):
        return None
    return dist.group.WORLD


# Zero-copy hand-off between drop-in modules (train.py:182-196 chains four of them: backbone -> ASPP -> decoder, and the
# domain classifier on the backbone's features).  to_nchw() remembers on the fp32 tensor it returns which NHWC bf16
# buffer it was converted from; when that very tensor -- same object, same version counter -- comes back as the input
# (or, in the backward pass, as the incoming gradient) of another module, to_nhwc() hands out the bf16 buffer instead of
# converting the fp32 copy back.  bf16 -> fp32 -> bf16 is the identity, so the values are the same bit for bit.  An
# incoming GRADIENT is taken only once (consume=True): a module's backward may accumulate into it in place.
# S2R_ZERO_COPY=0 disables the hand-off.
ZERO_COPY = [os.environ.get("S2R_ZERO_COPY", "1") != "0"]


def to_nhwc(cx, x, consume=False):
    """NCHW fp32 -> NHWC bf16 (channels zero-padded to a multiple of 8)."""
    if x.dim() != 4:
        raise ValueError('expected 4D input (got {}D input)'.format(x.dim()))
    src = getattr(x, "_s2r_act", None)
    if src is not None and ZERO_COPY[0]:
        a, version = src
        if consume:
            try:
                del x._s2r_act
            except AttributeError:
                pass
        if (version == x._version and x.dtype == torch.float32 and a.t.device == x.device
                and (a.N, a.C, a.H, a.W) == tuple(x.shape)):
            return Act(a.t, a.C, a.off)
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    N, Cc, H, W = x.shape
    a = cx.new(N, H, W, round_up(Cc, 8))
    L.call("s2r_nchw_f32_to_nhwc_bf16", _vp(x), N, Cc, H * W, a.vp(), a.pitch, cx.stream)
    a.C = Cc
    return a


def to_nchw(cx, a, Cc=None):
    Cc = a.C if Cc is None else Cc
    y = torch.empty((a.N, Cc, a.H, a.W), dtype=torch.float32, device=cx.device)
    L.call("s2r_nhwc_bf16_to_nchw_f32", a.vp(), a.pitch, a.N, Cc, a.H * a.W, _vp(y), cx.stream)
    if Cc == a.C:
        y._s2r_act = (a, y._version)      # see ZERO_COPY
    return y


class ModuleFn(torch.autograd.Function):
    """forward(module, make_run, n_in, *inputs_and_params) -> tuple of NCHW fp32 outputs."""

    @staticmethod
    def forward(ctx, module, make_run, n_in, *args):
        inputs = args[:n_in]
        dev = inputs[0].device
        if dev.type != "cuda":
            raise L.S2RError("s2r_b200 modules run on CUDA tensors only (got %s); there is no CPU path" % dev)
        with torch.cuda.device(dev):
            cx = Ctx(dev, module.training, sync_group_for(module) if module.training else None,
                     dropout=not getattr(module, "_s2r_no_dropout", False))
            # (grad mode is always off inside Function.forward and needs_input_grad ignores it: call_module records
            # whether the CALLER had it on)
            cx.inference = (not module.training) and not _CALLER_GRAD[0]
            if cx.inference:
                bn_eval_all(cx, module)
            run = make_run()
            raw = getattr(run, "raw_inputs", False)
            acts = [RawNCHW(x) if raw else to_nhwc(cx, x) for x in inputs]
            outs = run.forward(cx, *acts)
            if not isinstance(outs, tuple):
                outs = (outs,)
            res = tuple(run.export(cx, i, o) for i, o in enumerate(outs))
        ctx.run = run
        ctx.n_in = n_in
        ctx.module = module
        ctx.dev = dev
        ctx.in_shapes = [tuple(x.shape) for x in inputs]
        ctx.set_materialize_grads(False)
        return res if len(res) > 1 else res[0]

    @staticmethod
    def backward(ctx, *douts):
        run, module, dev = ctx.run, ctx.module, ctx.dev
        with torch.cuda.device(dev):
            cx = Ctx(dev, True, sync_group_for(module), dropout=False, async_wgrad=True)
            dacts = tuple(run.import_grad(cx, i, d) for i, d in enumerate(douts))
            need = ctx.needs_input_grad[3:3 + ctx.n_in]
            dins = run.backward(cx, dacts, need)
            if not isinstance(dins, tuple):
                dins = (dins,)
            res = []
            for i, d in enumerate(dins):
                if d is None or not need[i]:
                    res.append(None)
                elif isinstance(d, torch.Tensor):   # the run already produced the NCHW fp32 gradient
                    res.append(d)
                else:
                    res.append(to_nchw(cx, d, ctx.in_shapes[i][1]))
            cx.join()        # side-stream weight gradients are ordered before whatever follows this backward
        ctx.run = None
        return (None, None, None) + tuple(res) + (None,) * (len(ctx.needs_input_grad) - 3 - ctx.n_in)


class RunBase:
    """Default export/import: plain layout conversion at the module boundary."""

    def export(self, cx, i, act):
        return to_nchw(cx, act)

    def import_grad(self, cx, i, d):
        if d is None:
            return None
        return to_nhwc(cx, d, consume=True)


_CALLER_GRAD = [True]


def call_module(module, make_run, inputs):
    params = [p for p in module.parameters()]
    _CALLER_GRAD[0] = torch.is_grad_enabled()
    try:
        return ModuleFn.apply(module, make_run, len(inputs), *inputs, *params)
    finally:
        _CALLER_GRAD[0] = True
