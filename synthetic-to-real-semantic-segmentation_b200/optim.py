"""Multi-tensor optimizers on the C-ABI kernels (s2r_sgd_step / s2r_adam_step) with the
torch.optim surface the reference's scripts use: param_groups with a mutable 'lr'
(utils/lr_scheduler.py:63-70 writes it), step(), zero_grad().

Gradients of all parameters live in ONE flat fp32 buffer (p.grad are views): zero_grad is a single
memset, the data-parallel gradient all-reduce is a single NCCL call, and the slot table the
kernels walk is built once.  Learning rate / bias corrections live in a small device buffer that
advance() rewrites before every step with a stream-ordered store whose values travel as kernel
arguments (s2r_store_f32): they are snapshotted when advance() is called, so a captured CUDA graph
replays with the schedule value of ITS step even when the host runs several steps ahead of the GPU.
"""
import ctypes as C
import math

import torch

from . import _lib as L
from .engine import _vp, WEIGHT_EPOCH


def _groups(params, defaults):
    params = list(params)
    if len(params) == 0:
        raise ValueError("optimizer got an empty parameter list")
    if not isinstance(params[0], dict):
        params = [{'params': params}]
    groups = []
    for g in params:
        g = dict(g)
        g['params'] = list(g['params'])
        for k, v in defaults.items():
            g.setdefault(k, v)
        groups.append(g)
    return groups


SLOT_ELEMS = 32768   # elements per optimizer slot (16 CTAs x 256 threads x 8)


class _FusedBase(object):
    n_state = 1

    def __init__(self, params, defaults):
        self.param_groups = _groups(params, defaults)
        self.defaults = defaults
        allp = [p for g in self.param_groups for p in g['params']]
        if len(set(id(p) for p in allp)) != len(allp):
            raise ValueError("some parameters appear in more than one parameter group")
        dev = allp[0].device
        if dev.type != "cuda":
            raise L.S2RError("fused optimizers need CUDA parameters; there is no CPU path")
        self.device = dev
        total = sum(p.numel() for p in allp)
        # flat gradient buffer; existing gradients are carried over
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.state_bufs = [torch.zeros(total, dtype=torch.float32, device=dev) for _ in range(self.n_state)]
        off = 0
        self._views = []
        for p in allp:
            n = p.numel()
            gv = self.flat_grad[off:off + n].view_as(p)
            if p.grad is not None:
                gv.copy_(p.grad)
            p.grad = gv
            self._views.append((p, off, n))
            off += n
        self._tables = None
        self._hyper_host = [[0.0, 1.0, 1.0, 0.0] for _ in self.param_groups]
        self._hyper_dev = [torch.zeros(4, dtype=torch.float32, device=dev) for _ in self.param_groups]
        self.steps = 0
        self.grad_scale = 1.0

    def _build_tables(self):
        tables = []
        lookup = {id(p): (off, n) for p, off, n in self._views}
        for g in self.param_groups:
            ps = [p for p in g['params'] if p.requires_grad or p.grad is not None]
            # large tensors are split into slots of <= SLOT_ELEMS elements: the kernel gives every slot the same
            # 16 CTAs, so one 2-M-element filter in a single slot (FCDiscriminator.conv4) was 0.34 ms on its own
            pieces = []
            for p in ps:
                off, n = lookup[id(p)]
                if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                    # the caller replaced .grad (e.g. zero_grad(set_to_none=True) elsewhere): re-attach
                    gv = self.flat_grad[off:off + n].view_as(p)
                    if p.grad is not None:
                        gv.copy_(p.grad)
                    else:
                        gv.zero_()
                    p.grad = gv
                for o in range(0, n, SLOT_ELEMS):
                    pieces.append((p.data_ptr() + 4 * o, p.grad.data_ptr() + 4 * o, off + o, min(SLOT_ELEMS, n - o)))
            arr = (L.ParamSlot * max(1, len(pieces)))()
            for i, (pp, gp, off, n) in enumerate(pieces):
                arr[i].p = pp
                arr[i].g = gp
                arr[i].s0 = self.state_bufs[0].data_ptr() + 4 * off
                arr[i].s1 = self.state_bufs[1].data_ptr() + 4 * off if self.n_state > 1 else None
                arr[i].n = n
                arr[i].lr_mult = 1.0
            raw = bytes(arr)[:C.sizeof(L.ParamSlot) * len(pieces)]
            dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device) if len(pieces) else None
            tables.append((dev, len(pieces)))
        self._tables = tables

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    def all_reduce_grads(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, group=group)
            self.grad_scale = 1.0 / dist.get_world_size(group)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _publish(self):
        """Stream-ordered store of the hyper-parameters computed by advance() (values passed by value: no host
        buffer is read after this returns).  Must not run inside a CUDA-graph capture: the captured step holds only
        launch(); advance() runs on the replay stream ahead of every replay."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("optimizer.advance() inside a CUDA-graph capture: capture launch() only")
        with torch.cuda.device(self.device):
            st = self._stream()
            for h, d in zip(self._hyper_host, self._hyper_dev):
                L.call("s2r_store_f32", _vp(d), 4, (C.c_float * 4)(*h), st)

    # state names in torch.optim's state_dict (torch.optim.SGD: 'momentum_buffer'; Adam: 'exp_avg', 'exp_avg_sq')
    state_names = ('momentum_buffer',)

    def state_dict(self):
        """torch.optim-format state dict -- what train_adapt.py:206 / train.py:248-251 store under 'optimizer' --
        so checkpoints move between the reference's optimizers and these: {'state': {index: {name: tensor, ...}},
        'param_groups': [{..hyper-parameters.., 'params': [indices]}]} with parameters numbered in group order."""
        state, groups, idx = {}, [], 0
        lookup = {id(p): (off, n) for p, off, n in self._views}
        for g in self.param_groups:
            ids = []
            for p in g['params']:
                off, n = lookup[id(p)]
                if self.steps > 0:
                    ent = {name: self.state_bufs[k][off:off + n].view_as(p).clone() for k, name in enumerate(self.state_names)}
                    if 'exp_avg' in ent:
                        ent['step'] = torch.tensor(float(self.steps))
                    state[idx] = ent
                ids.append(idx)
                idx += 1
            pg = {k: v for k, v in g.items() if k != 'params'}
            pg['params'] = ids
            groups.append(pg)
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, sd):
        """Accepts torch.optim-format dicts (reference checkpoints) and the flat format of earlier versions."""
        if isinstance(sd.get('state'), (list, tuple)):      # {'steps', 'state': [flat buffers], 'param_groups'}
            self.steps = sd['steps']
            for b, s_ in zip(self.state_bufs, sd['state']):
                b.copy_(s_)
            for g, s_ in zip(self.param_groups, sd['param_groups']):
                g.update({k: v for k, v in s_.items() if k != 'params'})
            return
        if len(sd['param_groups']) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        lookup = {id(p): (off, n) for p, off, n in self._views}
        idx, steps = 0, 0
        for g, sg in zip(self.param_groups, sd['param_groups']):
            if len(sg['params']) != len(g['params']):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            g.update({k: v for k, v in sg.items() if k != 'params'})
            for p, key in zip(g['params'], sg['params']):
                ent = sd['state'].get(key)
                off, n = lookup[id(p)]
                for k, name in enumerate(self.state_names):
                    if ent is not None and ent.get(name) is not None:
                        self.state_bufs[k][off:off + n].copy_(ent[name].reshape(-1))
                    else:
                        self.state_bufs[k][off:off + n].zero_()
                if ent is not None:
                    steps = max(steps, int(float(ent['step'])) if 'step' in ent else 1)
                idx += 1
        # SGD: the first step initialises the momentum buffer with the gradient (torch.optim.SGD); a loaded buffer
        # means that step is behind us.  Adam: the bias corrections continue from the stored step count.
        self.steps = steps


class FusedSGD(_FusedBase):
    """torch.optim.SGD(params, lr, momentum, weight_decay, nesterov) -- train_adapt.py:58-59."""
    n_state = 1

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))

    def advance(self):
        """Host half of a step: the current learning rates go to the device buffer the kernels read (stream-ordered
        ahead of the eager launch() or of the CUDA-graph replay that follows)."""
        for gi, g in enumerate(self.param_groups):
            self._hyper_host[gi][0] = float(g['lr'])
        self.steps += 1
        self._publish()

    @torch.no_grad()
    def launch(self):
        """Device half of a step (graph-capturable): one multi-tensor kernel per parameter group."""
        if self._tables is None:
            self._build_tables()
        with torch.cuda.device(self.device):
            for gi, g in enumerate(self.param_groups):
                tab, n = self._tables[gi]
                if n == 0:
                    continue
                L.call("s2r_sgd_step", _vp(tab), n, _vp(self._hyper_dev[gi]), float(g['momentum']),
                       float(g['dampening']), float(g['weight_decay']), 1 if g['nesterov'] else 0,
                       float(self.grad_scale), self._stream())
        WEIGHT_EPOCH[0] += 1     # the parameters change behind torch's version counters: bf16 filter copies are stale

    def step(self):
        self.advance()
        self.launch()


class FusedAdam(_FusedBase):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) -- train_adapt.py:60, train.py:77-80."""
    n_state = 2
    state_names = ('exp_avg', 'exp_avg_sq')

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    def advance(self):
        self.steps += 1
        for gi, g in enumerate(self.param_groups):
            b1, b2 = g['betas']
            h = self._hyper_host[gi]
            h[0] = float(g['lr'])
            h[1] = 1.0 - math.pow(b1, self.steps)
            h[2] = 1.0 - math.pow(b2, self.steps)
        self._publish()

    @torch.no_grad()
    def launch(self):
        if self._tables is None:
            self._build_tables()
        with torch.cuda.device(self.device):
            for gi, g in enumerate(self.param_groups):
                tab, n = self._tables[gi]
                if n == 0:
                    continue
                b1, b2 = g['betas']
                L.call("s2r_adam_step", _vp(tab), n, _vp(self._hyper_dev[gi]), float(b1), float(b2), float(g['eps']),
                       float(g['weight_decay']), float(self.grad_scale), self._stream())
        WEIGHT_EPOCH[0] += 1

    def step(self):
        self.advance()
        self.launch()
