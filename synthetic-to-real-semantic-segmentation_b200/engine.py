"""Host-side execution engine: NHWC bf16 activation views, tap tables for the tap-GEMM
convolutions, thin wrappers over the C ABI and the composite layers (conv+BN+act,
InvertedResidual) with hand-written backward passes.

Everything here only moves pointers and sizes; the arithmetic is in csrc/*.cu.  torch supplies
device memory (caching allocator), the current stream and, for data-parallel runs,
torch.distributed (NCCL) for the BN-statistics and gradient all-reduces.
"""
import ctypes as C
import os
import weakref

import torch

from . import _lib as L

BF16 = torch.bfloat16
TRACE = None
_SEED_DEV = {}


def seed_counter(device):
    """Per-device int64 counter added to every dropout seed on the device side; steps.py bumps it once
    per training step with a (graph-capturable) device op so replays of a captured step draw new masks."""
    key = (device.type, device.index)
    t = _SEED_DEV.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device=device)
        _SEED_DEV[key] = t
    return t


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def round_up(a, b):
    return (a + b - 1) // b * b


class Act:
    """View of an NHWC bf16 tensor: channels [off, off+C) of storage t[N,H,W,pitch]."""
    __slots__ = ("t", "N", "H", "W", "C", "pitch", "off")

    def __init__(self, t, C=None, off=0):
        assert t.dtype == BF16 and t.dim() == 4 and t.is_contiguous()
        self.t = t
        self.N, self.H, self.W, self.pitch = t.shape
        self.off = off
        self.C = (self.pitch - off) if C is None else C

    @property
    def ptr(self):
        return self.t.data_ptr() + 2 * self.off

    @property
    def P(self):
        return self.N * self.H * self.W

    def slice(self, off, C):
        return Act(self.t, C, self.off + off)

    def vp(self):
        return C.c_void_p(self.ptr)


POISON = bool(os.environ.get("S2R_POISON"))   # fill uninitialised activation buffers with NaN (tests/debug)
KEEP_ALL = [] if os.environ.get("S2R_KEEP_ALL") else None   # debug: never free engine buffers (lifetime-race bisection)


PEER = {"world": 0, "slot": 0}
# exchange channel of the BN-statistics all-reduce: every rank issues the same sequence of exchanges per channel, so
# the two streams of steps.AdaptStep use one channel each (set by the step around the passes it issues on stream B)
COMM_CHANNEL = [0]


def init_peer_exchange(group=None, slot=4096):
    """Set up the NVLink peer-memory exchange for the BN statistics (csrc/comm.cu): allocate this rank's inbox,
    swap CUDA IPC handles through torch.distributed (plumbing) and open the peers.  Idempotent; S2R_COMM=nccl
    keeps every exchange on NCCL."""
    import torch.distributed as dist
    if os.environ.get("S2R_COMM", "") == "nccl" or not (dist.is_available() and dist.is_initialized()):
        return False
    world = dist.get_world_size(group)
    if world <= 1 or PEER["world"] == world:
        return PEER["world"] == world
    rank = dist.get_rank(group)
    handle = C.create_string_buffer(64)
    L.call("s2r_comm_create", rank, world, slot, C.cast(handle, C.c_void_p))
    handles = [None] * world
    dist.all_gather_object(handles, handle.raw, group=group)
    blob = C.create_string_buffer(b"".join(handles), 64 * world)
    L.call("s2r_comm_open", C.cast(blob, C.c_void_p))
    torch.cuda.synchronize()
    dist.barrier(group=group)
    PEER["world"], PEER["slot"] = world, slot
    return True


def dp_world():
    """Number of data-parallel ranks (1 without an initialised torch.distributed)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def allreduce_small_f64(t, group=None):
    """Sum of a small contiguous fp64 vector over the ranks on the current stream: the NVLink peer-memory kernel
    (current exchange channel) once the exchange is set up, NCCL otherwise.  Every rank must call it in the same
    order per stream."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world <= 1:
        return
    if PEER["world"] == world and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= PEER["slot"]:
        L.call("s2r_allreduce_small_f64_ch", _vp(t), t.numel(), COMM_CHANNEL[0],
               C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream))
        return
    dist.all_reduce(t, group=group)


_SIDE_STREAMS = {}
FORK_ONLY = [k for k in os.environ.get("S2R_FORK_ONLY", "").split(",") if k]   # debug: fork only these wgrad kinds
FORK_JOIN = bool(os.environ.get("S2R_FORK_JOIN"))                             # debug: join right after every fork
WGRAD_STREAM = os.environ.get("S2R_WGRAD_STREAM", "1") != "0"


class _Fork(object):
    """Runs the enclosed launches on the context's side stream, after everything issued on the main stream so far."""

    def __init__(self, cx, keep):
        self.cx, self.keep = cx, keep

    def __enter__(self):
        cx = self.cx
        main = torch.cuda.current_stream(cx.device)
        ev = torch.cuda.Event()
        ev.record(main)
        cx.side.wait_event(ev)
        cx._forked = True
        cx._keep.extend(self.keep)       # operands stay allocated until join(): the main stream must not reuse them
        self._ctx = torch.cuda.stream(cx.side)
        self._ctx.__enter__()

    def __exit__(self, *a):
        self._ctx.__exit__(*a)
        if FORK_JOIN:
            self.cx.join()


class _NoFork(object):
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class Ctx:
    """Per-call execution context (stream, mode, cross-rank synchronisation of BN)."""

    def __init__(self, device, training, sync_group=None, dropout=True, async_wgrad=False):
        L.require_cuda()
        self.device = device
        self.training = training
        # inference: eval mode AND no backward pass will follow (set by runtime.ModuleFn under torch.no_grad()):
        # BatchNorm folds into the producing kernels' epilogues and nothing is saved for a backward
        self.inference = False
        self.eval_states = None   # inference: {id(bn): BNState} filled by bn_eval_all (one launch for all layers)
        # weight gradients are leaves of the backward pass (nothing downstream reads them before the optimizer): with
        # async_wgrad they are issued on a side stream and overlap the data-gradient chain, whose many small kernels
        # leave most SMs idle; join() orders them before whatever follows the module's backward
        self.side = None
        self._keep = []
        self._forked = False
        if async_wgrad and WGRAD_STREAM:
            # one side stream per ambient stream (two backward passes may run on two streams, steps.AdaptStep)
            key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
            self.side = _SIDE_STREAMS[key]
        self.group = sync_group
        self.world = 1
        if sync_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(sync_group)
            if self.world > 1 and PEER["world"] != self.world and not torch.cuda.is_current_stream_capturing():
                init_peer_exchange(sync_group)   # collective: every rank builds its first synchronised Ctx together
        self.dropout = dropout
        self.seed_dev = seed_counter(device)
        self.trace = TRACE  # when a list: receives (name, Act) of intermediate activations (tests/tools)
        # host-side seed stream for dropout masks (regenerated, never stored)
        self._seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (training and dropout) else 0

    @property
    def stream(self):
        # looked up per call: the current stream changes under torch.cuda.graph / torch.cuda.stream
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def fork(self, *keep, kind=""):
        """`with cx.fork(tensors...):` -- launches inside go to the side stream (no-op without async_wgrad)."""
        if self.side is None or (FORK_ONLY and kind not in FORK_ONLY):
            return _NoFork()
        return _Fork(self, keep)

    def join(self):
        if self.side is not None and self._forked:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
            self._forked = False
        self._keep = []

    def next_seed(self):
        self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        return self._seed

    def new(self, N, H, W, Cc, zero=False):
        f = torch.zeros if zero else torch.empty
        t = f((N, H, W, Cc), dtype=BF16, device=self.device)
        if POISON and not zero:
            t.fill_(float('nan'))   # debug: an element that is read before it is written poisons the result
        if KEEP_ALL is not None:
            KEEP_ALL.append(t)
        return Act(t)

    def f32(self, n, zero=True):
        t = (torch.zeros if zero else torch.empty)(n, dtype=torch.float32, device=self.device)
        if POISON and not zero:
            t.fill_(float('nan'))
        if KEEP_ALL is not None:
            KEEP_ALL.append(t)
        return t

    def f64(self, n):
        """Zeroed fp64 accumulator (BN sums): slices of an arena zeroed with ONE memset per 64 K doubles instead of
        one fill kernel per BatchNorm layer (~220 launches per step)."""
        n2 = (n + 1) & ~1   # keep 16-byte alignment
        a = getattr(self, "_f64_arena", None)
        if a is None or self._f64_used + n2 > a.numel():
            a = torch.zeros(max(65536, n2), dtype=torch.float64, device=self.device)
            self._f64_arena, self._f64_used = a, 0
            if KEEP_ALL is not None:
                KEEP_ALL.append(a)
        out = a[self._f64_used:self._f64_used + n]
        self._f64_used += n2
        return out

    def tr(self, name, act):
        if self.trace is not None:
            self.trace.append((name, act))
        return act

    def allreduce(self, t):
        """Sum of a small statistics vector over the ranks: the NVLink peer-memory kernel when the exchange has
        been set up (init_peer_exchange), NCCL otherwise."""
        if self.world > 1:
            allreduce_small_f64(t, self.group)


# --------------------------------------------------------------------------- weights
# Bumped by the fused optimizers, which update parameters through raw pointers and therefore do
# not touch torch's per-tensor version counter.
WEIGHT_EPOCH = [0]


def _pack_dims(w, mode):
    """(Cout, Cin, R*S, slices, A_pad, B_pad) of the packed copy of an OIHW filter; modes as s2r_pack_weight."""
    Cout, Cin, R, S = w.shape
    mode = int(mode)
    Cp = round_up(Cin, 8)
    if mode <= 1:
        A, B, ns = ((Cin, Cout) if mode else (Cout, Cin)) + (R * S,)
    elif mode == 2:
        A, B, ns = Cout, 4 * Cp, 4
    else:
        A, B, ns = 2 * Cp, Cout, 4
    return Cout, Cin, R * S, ns, round_up(A, 16), round_up(B, 64)


# every (filter, orientation) that has been packed: prepack_weights() refreshes all stale ones in one launch
_PACK_REGISTRY = {}
_PACK_JOB_ELEMS = 4096


def packed_weight(cx, w, mode):
    """bf16 [slices][A_pad][B_pad] copy of an OIHW fp32 parameter, cached per parameter version.
    mode: False/0 forward, True/1 data gradient, 2..4 the row-tap forms of a 4x4 stride-2 filter (s2r_pack_weight)."""
    mode = int(mode)
    cache = getattr(w, "_s2r_pack", None)
    if cache is None:
        cache = {}
        w._s2r_pack = cache
    key = (mode, w.data_ptr())
    ent = cache.get(key)
    stamp = (w._version, WEIGHT_EPOCH[0])
    if ent is not None and ent[0] == stamp:
        return ent[1], ent[2], ent[3]
    Cout, Cin, RS, ns, A_pad, B_pad = _pack_dims(w, mode)
    buf = ent[1] if ent is not None else torch.empty((ns, A_pad, B_pad), dtype=BF16, device=w.device)
    L.call("s2r_pack_weight", _vp(w.detach()), Cout, Cin, w.shape[2], w.shape[3], mode, _vp(buf),
           A_pad, B_pad, cx.stream)
    cache[key] = (stamp, buf, A_pad, B_pad)
    _PACK_REGISTRY[(id(w), mode)] = (weakref.ref(w), mode)
    return buf, A_pad, B_pad


def prepack_weights(stream, build_only=False):
    """Re-pack, in ONE launch, every filter copy whose parameter changed since it was packed (call after the
    optimizer kernels / at the start of a step).  Later packed_weight() calls then hit the cache.
    build_only: only upload the job table (a host->device copy, which must happen outside CUDA-graph capture) and
    return it: a CUDA graph that captures the launch holds raw pointers into the table, so whoever captures keeps the
    returned object alive with the graph."""
    stale = []
    for k, (ref, transpose) in list(_PACK_REGISTRY.items()):
        w = ref()
        if w is None:
            del _PACK_REGISTRY[k]
            continue
        ent = w._s2r_pack.get((transpose, w.data_ptr()))
        stamp = (w._version, WEIGHT_EPOCH[0])
        if ent is None or ent[0] == stamp:
            continue
        stale.append((w, transpose, ent, stamp))
    if not stale:
        return 0
    sig = tuple((w.data_ptr(), t, ent[1].data_ptr()) for w, t, ent, _ in stale)
    tab = _PACK_TABLES.get(sig)
    if tab is None:
        jobs = []
        for w, transpose, ent, _ in stale:
            Cout, Cin, RS, ns, A_pad, B_pad = _pack_dims(w, transpose)
            if int(transpose) <= 1 and RS > 1:
                # several taps: a job covers (a, b) pairs and writes all taps of each (contiguous in the OIHW source)
                total, mode, per = A_pad * B_pad, int(transpose) + 16, max(256, _PACK_JOB_ELEMS // RS)
            else:
                total, mode, per = ns * A_pad * B_pad, int(transpose), _PACK_JOB_ELEMS
            for b in range(0, total, per):
                jobs.append((w.data_ptr(), ent[1].data_ptr(), Cout, Cin, RS, mode, A_pad, B_pad, b, min(total, b + per)))
        chunks = []
        for s0 in range(0, len(jobs), 65535):
            part = jobs[s0:s0 + 65535]
            arr = (L.PackJob * len(part))()
            for i, j in enumerate(part):
                (arr[i].w, arr[i].packed, arr[i].Cout, arr[i].Cin, arr[i].RS, arr[i].transpose, arr[i].A_pad, arr[i].B_pad,
                 arr[i].begin, arr[i].end) = j
            dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(stale[0][0].device)
            chunks.append((dev, len(part)))
        tab = chunks
        while len(_PACK_TABLES) >= 8:            # a handful of live signatures (train / val / second model set)
            _PACK_TABLES.pop(next(iter(_PACK_TABLES)))
        _PACK_TABLES[sig] = tab
    if build_only:
        return tab
    for dev, n in tab:
        L.call("s2r_pack_weights_multi", _vp(dev), n, stream)
    for w, transpose, ent, stamp in stale:
        w._s2r_pack[(transpose, w.data_ptr())] = (stamp, ent[1], ent[2], ent[3])
    return len(stale)


_PACK_TABLES = {}


def invalidate_packs():
    """Mark every cached bf16 filter copy stale (call after parameters were rewritten behind torch's version counters,
    e.g. through raw pointers); the next packed_weight() / prepack_weights() refreshes them."""
    WEIGHT_EPOCH[0] += 1


def grad_of(p):
    """fp32 gradient buffer of a parameter, allocated (zero) on first use; kernels accumulate."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


class RawNCHW:
    """A module input kept as the caller's NCHW fp32 tensor (runs with `raw_inputs = True`): the few-channel
    first convolutions read it through the patch kernel instead of a NHWC copy."""
    __slots__ = ("t", "N", "C", "H", "W")

    def __init__(self, t):
        if t.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(t.dim()))
        t = t.contiguous()
        if t.dtype != torch.float32:
            t = t.float()
        self.t = t
        self.N, self.C, self.H, self.W = t.shape


def im2col(cx, raw, R, S, stride, pad):
    """Patch matrix [N,OH,OW,round_up(C*R*S,8)] (bf16) of an RxS conv on a RawNCHW input."""
    OH, OW = conv_out_hw(raw.H, raw.W, R, S, stride, pad, 1)
    K = raw.C * R * S
    P = cx.new(raw.N, OH, OW, round_up(K, 8))
    L.call("s2r_im2col_nchw_f32", _vp(raw.t), raw.N, raw.C, raw.H, raw.W, R, S, stride, pad, P.vp(), P.pitch,
           cx.stream)
    P.C = K
    return P


def patch_weight(w):
    """The OIHW filter viewed as the 1x1 filter [Cout, C*R*S, 1, 1] of the patch GEMM (same storage)."""
    v = getattr(w, "_s2r_w2", None)
    if v is None or v.data_ptr() != w.data_ptr():
        v = w.detach().view(w.shape[0], -1, 1, 1)
        w._s2r_w2 = v
    return v


# --------------------------------------------------------------------------- tap tables
def _fill_tap(tap, base, sn, sh, sw, H, W, dh, dw, wslice, wofs):
    tap.base = base
    tap.sn, tap.sh, tap.sw = sn, sh, sw
    tap.H, tap.W = H, W
    tap.dh, tap.dw = dh, dw
    tap.wslice = wslice
    tap.wofs = wofs


def fwd_taps(taps, x, R, S, stride, pad, dil):
    """Forward taps of an RxS conv on x: one strided view per input parity (stride 1: one view)."""
    n = 0
    for kh in range(R):
        eh = kh * dil - pad
        ph = eh % stride
        for kw in range(S):
            ew = kw * dil - pad
            pw = ew % stride
            base = x.ptr + 2 * (ph * x.W + pw) * x.pitch
            _fill_tap(taps[n], base, x.H * x.W * x.pitch, stride * x.W * x.pitch, stride * x.pitch,
                      (x.H - ph + stride - 1) // stride, (x.W - pw + stride - 1) // stride,
                      (eh - ph) // stride, (ew - pw) // stride, kh * S + kw, kh * S + kw)
            n += 1
    return n


def conv_out_hw(H, W, R, S, stride, pad, dil):
    return ((H + 2 * pad - dil * (R - 1) - 1) // stride + 1, (W + 2 * pad - dil * (S - 1) - 1) // stride + 1)


def conv_fwd(cx, x, w, out, stride=1, pad=0, dil=1, bias=None, act=L.ACT_NONE, slope=0.0, stats=None,
             aux=None, aux_mode=L.AUX_NONE, force_mma=False, oscale=None):
    """out = epi(conv(x, w)); x/out are Act views, w an OIHW fp32 parameter.
    oscale: per-output-channel fp32 scale applied to the accumulator before the bias (a folded eval-mode BatchNorm)."""
    Cout, Cin, R, S = w.shape
    assert x.C >= Cin and out.C >= Cout, (x.C, Cin, out.C, Cout)
    OH, OW = conv_out_hw(x.H, x.W, R, S, stride, pad, dil)
    assert (out.N, out.H, out.W) == (x.N, OH, OW), ((out.N, out.H, out.W), (x.N, OH, OW))
    wp, Cout_pad, Kpad = packed_weight(cx, w, False)
    a = L.ConvArgs()
    a.struct_size = C.sizeof(L.ConvArgs)
    a.ntaps = fwd_taps(a.taps, x, R, S, stride, pad, dil)
    a.N, a.OH, a.OW = x.N, OH, OW
    a.Cin, a.Cout = round_up(Cin, 8), Cout
    assert x.pitch - x.off >= a.Cin
    a.w, a.Cout_pad, a.Kpad = wp.data_ptr(), Cout_pad, Kpad
    a.out = out.ptr
    a.on, a.oh, a.ow = out.H * out.W * out.pitch, out.W * out.pitch, out.pitch
    a.bias = bias.data_ptr() if bias is not None else None
    a.act, a.slope = act, slope
    a.aux_mode = aux_mode
    if aux is not None:
        a.aux = aux.ptr
        a.an, a.ah, a.aw = aux.H * aux.W * aux.pitch, aux.W * aux.pitch, aux.pitch
    a.stats = stats.data_ptr() if stats is not None else None
    a.oscale = oscale.data_ptr() if oscale is not None else None
    L.call("s2r_conv_fwd_mma" if force_mma else "s2r_conv_fwd", C.byref(a), cx.stream)
    return out


def conv_dgrad(cx, dy, w, dx, stride=1, pad=0, dil=1, aux=None, aux_mode=L.AUX_NONE, slope=0.0,
               force_mma=False, stats=None):
    """dx = conv_transpose(dy, w) written per input-parity class; dx is an Act [N,H,W,>=Cin].
    stats (fp64 [2][Cin], zeroed): += per-channel sum / sum of squares of the stored dx (all parity classes add into
    it) -- the bias gradient of the layer below without a pass of its own."""
    Cout, Cin, R, S = w.shape
    assert dy.pitch - dy.off >= round_up(Cout, 8) and dx.C >= Cin
    wp, Cin_pad, Kpad = packed_weight(cx, w, True)
    H, W = dx.H, dx.W
    for ph in range(min(stride, H)):
        for pw in range(min(stride, W)):
            a = L.ConvArgs()
            a.struct_size = C.sizeof(L.ConvArgs)
            n = 0
            for kh in range(R):
                if (ph + pad - kh * dil) % stride:
                    continue
                for kw in range(S):
                    if (pw + pad - kw * dil) % stride:
                        continue
                    _fill_tap(a.taps[n], dy.ptr, dy.H * dy.W * dy.pitch, dy.W * dy.pitch, dy.pitch, dy.H, dy.W,
                              (ph + pad - kh * dil) // stride, (pw + pad - kw * dil) // stride, kh * S + kw, 0)
                    n += 1
            OHc, OWc = (H - ph + stride - 1) // stride, (W - pw + stride - 1) // stride
            if n == 0:
                raise NotImplementedError("conv_dgrad: parity class without taps")
            a.ntaps = n
            a.N, a.OH, a.OW = dx.N, OHc, OWc
            a.Cin, a.Cout = round_up(Cout, 8), Cin
            a.w, a.Cout_pad, a.Kpad = wp.data_ptr(), Cin_pad, Kpad
            a.out = dx.ptr + 2 * (ph * W + pw) * dx.pitch
            a.on, a.oh, a.ow = H * W * dx.pitch, stride * W * dx.pitch, stride * dx.pitch
            a.bias = None
            a.act, a.slope = L.ACT_NONE, slope
            a.aux_mode = aux_mode
            if aux is not None:
                a.aux = aux.ptr + 2 * (ph * W + pw) * aux.pitch
                a.an, a.ah, a.aw = H * W * aux.pitch, stride * W * aux.pitch, stride * aux.pitch
            a.stats = stats.data_ptr() if stats is not None else None
            L.call("s2r_conv_fwd_mma" if force_mma else "s2r_conv_fwd", C.byref(a), cx.stream)
    return dx


WGRAD_TAP_MAJOR = os.environ.get("S2R_WGRAD_TAP_MAJOR", "1") != "0"


def conv_wgrad(cx, x, dy, w, stride=1, pad=0, dil=1, grad_param=None):
    """w.grad += conv weight gradient; bias handled by the caller.  grad_param: the parameter whose .grad
    receives the result when w is a reshaped view of it (patch GEMM)."""
    Cout, Cin, R, S = w.shape
    g = grad_of(w if grad_param is None else grad_param)
    a = L.WgradArgs()
    a.struct_size = C.sizeof(L.WgradArgs)
    a.ntaps = fwd_taps(a.taps, x, R, S, stride, pad, dil)
    a.N, a.OH, a.OW = dy.N, dy.H, dy.W
    a.Cin, a.Cout = Cin, Cout
    a.dy = dy.ptr
    a.dn, a.dh, a.dw = dy.H * dy.W * dy.pitch, dy.W * dy.pitch, dy.pitch
    with cx.fork(x.t, dy.t, kind="kxk" if R * S > 1 else "1x1"):
        if R * S > 1 and WGRAD_TAP_MAJOR:
            # tap-major fp32 scratch [R*S][Cout][Cp]: the kernel's atomics hit 32 consecutive input channels (one
            # 128-byte line) per instruction instead of 32 elements R*S floats apart; one small pass adds it into the
            # OIHW gradient
            Cp = round_up(Cin, 32)
            G = cx.f32(R * S * Cout * Cp)
            for t in range(a.ntaps):
                a.taps[t].wofs = a.taps[t].wofs * Cout * Cp
            a.dweight = G.data_ptr()
            a.s_co, a.s_ci = Cp, 1
            L.call("s2r_conv_wgrad", C.byref(a), cx.stream)
            L.call("s2r_wgrad_scatter_taps", _vp(G), _vp(g), Cout, Cin, R * S, Cp, cx.stream)
            return
        a.dweight = g.data_ptr()
        a.s_co, a.s_ci = Cin * R * S, R * S
        L.call("s2r_conv_wgrad", C.byref(a), cx.stream)


# --------------------------------------------------------------------------- row-tap 4x4 stride-2 convolution
class PadAct:
    """Zero-padded NHWC bf16 image [N][H+2][W+2][Cp] of a logical [N,H,W,C] tensor (Cp = round8(C)).  In this buffer the
    four kw taps of a row of a 4x4 stride-2 pad-1 filter (modeling/discriminator.py:11) are ONE contiguous run of 4*Cp
    channels starting at padded pixel (2*oh + kh, 2*ow), so the convolution is a 4-tap GEMM over an overlapping strided
    view -- no patch matrix (16x the input) and no 16 taps of 24 channels padded to 64."""
    __slots__ = ("t", "N", "H", "W", "C", "Cp")

    def __init__(self, t, H, W, Cc):
        self.t, self.N, self.H, self.W, self.C = t, t.shape[0], H, W, Cc
        self.Cp = t.shape[3]
        assert tuple(t.shape) == (self.N, H + 2, W + 2, self.Cp) and t.dtype == BF16 and t.is_contiguous()

    @property
    def ptr(self):
        return self.t.data_ptr()


def rowtap_ok(Cc, H, W):
    return Cc <= 64 and H % 2 == 0 and W % 2 == 0 and H >= 2 and W >= 2


def _rowtap_views(taps, xp):
    Hp, Wp, Cp = xp.H + 2, xp.W + 2, xp.Cp
    OH, OW = xp.H // 2, xp.W // 2
    for kh in range(4):
        _fill_tap(taps[kh], xp.ptr + 2 * kh * Wp * Cp, Hp * Wp * Cp, 2 * Wp * Cp, 2 * Cp, OH, OW, 0, 0, kh, kh * 4 * Cp)
    return OH, OW


def rowtap_fwd(cx, xp, w, out, bias=None, act=L.ACT_NONE, slope=0.0):
    """out = act(conv4x4_s2_p1(x, w) + bias) with x given as a PadAct."""
    Cout, Cin, R, S = w.shape
    assert (R, S) == (4, 4) and Cin == xp.C
    wp, A_pad, B_pad = packed_weight(cx, w, 2)
    a = L.ConvArgs()
    a.struct_size = C.sizeof(L.ConvArgs)
    a.ntaps = 4
    OH, OW = _rowtap_views(a.taps, xp)
    assert (out.N, out.H, out.W) == (xp.N, OH, OW) and out.C >= Cout
    a.N, a.OH, a.OW = xp.N, OH, OW
    a.Cin, a.Cout = 4 * xp.Cp, Cout
    a.w, a.Cout_pad, a.Kpad = wp.data_ptr(), A_pad, B_pad
    a.out = out.ptr
    a.on, a.oh, a.ow = out.H * out.W * out.pitch, out.W * out.pitch, out.pitch
    a.bias = bias.data_ptr() if bias is not None else None
    a.act, a.slope = act, slope
    a.aux_mode = L.AUX_NONE
    a.stats = None
    L.call("s2r_conv_fwd", C.byref(a), cx.stream)
    return out


def rowtap_wgrad(cx, xp, dy, w):
    """w.grad += weight gradient of the same convolution (accumulated in the row-tap layout, then scattered to OIHW)."""
    Cout, Cin, R, S = w.shape
    Cp = xp.Cp
    g = grad_of(w)
    a = L.WgradArgs()
    a.struct_size = C.sizeof(L.WgradArgs)
    a.ntaps = 4
    _rowtap_views(a.taps, xp)
    a.N, a.OH, a.OW = dy.N, dy.H, dy.W
    a.Cin, a.Cout = 4 * Cp, Cout
    a.dy = dy.ptr
    a.dn, a.dh, a.dw = dy.H * dy.W * dy.pitch, dy.W * dy.pitch, dy.pitch
    with cx.fork(xp.t, dy.t, kind="rowtap"):
        G = cx.f32(Cout * 16 * Cp)
        a.dweight = G.data_ptr()
        a.s_co, a.s_ci = 16 * Cp, 1
        L.call("s2r_conv_wgrad", C.byref(a), cx.stream)
        L.call("s2r_rowtap_wgrad_scatter", _vp(G), _vp(g), Cout, Cin, cx.stream)


def rowtap_dgrad(cx, dy, w, H, W):
    """Gradient w.r.t. the padded input as a PadAct (border rows / columns hold the irrelevant gradient of the zero
    padding): one launch per padded-row parity, each output "pixel" = two adjacent padded pixels (2*Cp channels)."""
    Cout, Cin, R, S = w.shape
    Cp = round_up(Cin, 8)
    Hp, Wp = H + 2, W + 2
    dxp = PadAct(torch.empty((dy.N, Hp, Wp, Cp), dtype=BF16, device=cx.device), H, W, Cin)
    for ph in range(2):
        wp, A_pad, B_pad = packed_weight(cx, w, 3 + ph)
        a = L.ConvArgs()
        a.struct_size = C.sizeof(L.ConvArgs)
        a.ntaps = 4
        for t in range(4):
            _fill_tap(a.taps[t], dy.ptr, dy.H * dy.W * dy.pitch, dy.W * dy.pitch, dy.pitch, dy.H, dy.W,
                      -(t >> 1), -(t & 1), t, 0)
        a.N, a.OH, a.OW = dy.N, Hp // 2, Wp // 2
        a.Cin, a.Cout = round_up(Cout, 8), 2 * Cp
        assert dy.pitch - dy.off >= a.Cin
        a.w, a.Cout_pad, a.Kpad = wp.data_ptr(), A_pad, B_pad
        a.out = dxp.ptr + 2 * ph * Wp * Cp
        a.on, a.oh, a.ow = Hp * Wp * Cp, 2 * Wp * Cp, 2 * Cp
        a.bias = None
        a.act, a.slope = L.ACT_NONE, 0.0
        a.aux_mode = L.AUX_NONE
        a.stats = None
        L.call("s2r_conv_fwd", C.byref(a), cx.stream)
    return dxp


def bias_grad(cx, dy, b, presummed=None):
    """b.grad += per-channel sum of dy.  presummed: fp64 sums the kernel that produced dy has already left
    (conv_dgrad(stats=...)): no pass over dy."""
    if presummed is None:
        Cc = round_up(b.numel(), 8)
        presummed = cx.f64(2 * Cc)
        L.call("s2r_channel_sums_bf16", dy.vp(), dy.P, Cc, dy.pitch, 0, _vp(presummed), cx.stream)
    L.call("s2r_add_f64_to_f32", _vp(presummed), _vp(grad_of(b)), b.numel(), cx.stream)


# --------------------------------------------------------------------------- batch norm
class BNState:
    """Per-forward BN results: scale/shift and mean/invstd ([2][C] fp32 each) + element count.
    pending: the s2r_bn_tail of a training-mode BatchNorm whose sums have been produced but not yet turned into
    ss / mi -- the first kernel that consumes the normalised tensor (depthwise prologue, bn_apply) does that in its
    own prologue and publishes ss / mi (csrc/bn_tail.cuh); take() hands the descriptor to that launch, ready() forces
    it with one small launch for any other consumer."""
    __slots__ = ("ss", "mi", "count", "frozen", "pending", "C", "keep")

    def __init__(self, ss, mi, count, frozen, pending=None, C=0, keep=None):
        self.ss, self.mi, self.count, self.frozen, self.pending, self.C = ss, mi, count, frozen, pending, C
        self.keep = keep   # the fp64 sums the pending descriptor points at (raw pointer): alive as long as the state

    def take(self):
        """The pending descriptor (or None), for a launch that finalises it; afterwards ss / mi are valid in stream order."""
        p, self.pending = self.pending, None
        return p

    def ready(self, cx):
        p = self.take()
        if p is not None:
            L.call("s2r_bn_tail_run", C.byref(p), self.C, cx.stream)
        return self


def bn_finalize(cx, bn, sums, count_local):
    """Turn (cross-rank reduced) channel sums into scale/shift; updates the running statistics.
    Mirrors _SynchronizedBatchNorm._compute_mean_std (batchnorm.py:113-125) when synchronised
    across ranks and F.batch_norm (batchnorm.py:50-53) otherwise."""
    Cc = bn.num_features
    ss = cx.f32(2 * Cc, zero=False)
    mi = cx.f32(2 * Cc, zero=False)
    sync = cx.world > 1 and getattr(bn, "_s2r_sync", False)
    if sync:
        cx.allreduce(sums)
    count = float(count_local) * (cx.world if sync else 1)
    if count <= 1:
        raise ValueError("BatchNorm computes unbiased standard-deviation, which requires size > 1.")
    mom = bn.momentum if bn.momentum is not None else 0.1
    L.call("s2r_bn_finalize", _vp(sums), count, _vp(bn.weight), _vp(bn.bias), float(bn.eps), 1 if sync else 0,
           float(mom), _vp(bn.running_mean), _vp(bn.running_var), _vp(mi), _vp(ss), Cc, cx.stream)
    return BNState(ss, mi, count, False)


def bn_eval_all(cx, module):
    """Inference: scale / shift of EVERY BatchNorm layer of `module` from its running statistics in one launch
    (s2r_bn_eval_multi); bn_eval() then hands out the per-layer states.  The job table and the output buffers are
    built once per module (they hold raw pointers to the parameters / buffers: rebuilt when those move) and live on the
    module; the launch is repeated on every forward pass, so a captured graph always sees the current statistics.
    Several streams may run it at once on the same buffers (validation graph lanes): they write identical values."""
    bns = [m for m in module.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    if not bns:
        return
    sig = tuple(t.data_ptr() for m in bns for t in (m.weight, m.bias, m.running_mean, m.running_var))
    ent = getattr(module, "_s2r_eval_tab", None)
    if ent is None or ent[0] != sig:
        if torch.cuda.is_current_stream_capturing():
            return          # table upload is a host->device copy: per-layer bn_eval launches in this capture
        total = sum(2 * m.num_features for m in bns)
        buf = torch.empty(2 * total, dtype=torch.float32, device=cx.device)
        arr = (L.BnEvalJob * len(bns))()
        states, off = {}, 0
        for i, m in enumerate(bns):
            Cc = m.num_features
            mi, ss = buf[off:off + 2 * Cc], buf[total + off:total + off + 2 * Cc]
            off += 2 * Cc
            j = arr[i]
            j.gamma, j.beta = m.weight.data_ptr(), m.bias.data_ptr()
            j.running_mean, j.running_var = m.running_mean.data_ptr(), m.running_var.data_ptr()
            j.mean_invstd, j.scale_shift, j.C, j.eps = mi.data_ptr(), ss.data_ptr(), Cc, float(m.eps)
            states[id(m)] = BNState(ss, mi, 0.0, True)
        tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(cx.device)
        ent = (sig, tab, len(bns), states, buf)
        module._s2r_eval_tab = ent
    L.call("s2r_bn_eval_multi", _vp(ent[1]), ent[2], cx.stream)
    cx.eval_states = ent[3]


def bn_eval(cx, bn):
    pre = cx.eval_states.get(id(bn)) if cx.eval_states is not None else None
    if pre is not None:
        return pre
    Cc = bn.num_features
    ss = cx.f32(2 * Cc, zero=False)
    mi = cx.f32(2 * Cc, zero=False)
    L.call("s2r_bn_eval_scale_shift", _vp(bn.weight), _vp(bn.bias), _vp(bn.running_mean), _vp(bn.running_var),
           float(bn.eps), _vp(mi), _vp(ss), Cc, cx.stream)
    return BNState(ss, mi, 0.0, True)


def bn_plan(cx, bn, count_local):
    """(sums, finish) for the BatchNorm `bn` that normalises the output of the next statistics-producing launch:
    hand `sums` (fp64 [2][C], None in eval mode) to the producer, then call finish() -> BNState.
    Training mode on one rank: the state is PENDING (no finalize launch; BNState).  Synchronised across ranks: the sums
    are exchanged and finalised by one small launch (csrc/comm.cu), or all-reduced over NCCL (S2R_COMM=nccl) first.
    Mirrors _SynchronizedBatchNorm._compute_mean_std (batchnorm.py:113-125) when synchronised across ranks and
    F.batch_norm (batchnorm.py:50-53) otherwise."""
    if not (cx.training and bn.training):
        st = bn_eval(cx, bn)
        return None, (lambda: st)
    Cc = bn.num_features
    sync = cx.world > 1 and getattr(bn, "_s2r_sync", False)
    count = float(count_local) * (cx.world if sync else 1)
    if count <= 1:
        raise ValueError("BatchNorm computes unbiased standard-deviation, which requires size > 1.")
    sums = cx.f64(2 * Cc)
    ss = cx.f32(2 * Cc, zero=False)
    mi = cx.f32(2 * Cc, zero=False)
    t = L.BnTail()
    t.count = count
    t.sums = sums.data_ptr()
    t.gamma, t.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
    t.running_mean, t.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
    t.mean_invstd, t.scale_shift = mi.data_ptr(), ss.data_ptr()
    t.eps = float(bn.eps)
    t.momentum = float(bn.momentum if bn.momentum is not None else 0.1)
    t.clamp_mode = 1 if sync else 0
    t.channel = -1
    st = BNState(ss, mi, count, False, pending=t, C=Cc, keep=sums)

    def finish():
        if sync:
            if PEER["world"] == cx.world:
                t.channel = COMM_CHANNEL[0]   # exchange over NVLink peer memory + finalize: one small launch
            else:
                cx.allreduce(sums)            # NCCL
            st.ready(cx)
        return st
    return sums, finish


def bn_apply(cx, z, st, act, out, residual=None, drop_p=0.0, seed=0):
    p = st.take()
    L.call("s2r_bn_apply_act_bn", z.vp(), z.P, z.C, z.pitch, 0, C.byref(p) if p is not None else None, _vp(st.ss), act,
           residual.vp() if residual is not None else None, float(drop_p), seed, _vp(cx.seed_dev), out.vp(),
           out.pitch, 0,
           cx.stream)
    return out


def bn_backward(cx, bn, dy, z, st, act, dx, drop_p=0.0, seed=0, presummed=None, win=None):
    """dz of a training-mode BN (+act, +dropout) given dy; accumulates gamma/beta grads.
    presummed: fp64 [2][C] sums already produced by the kernel that wrote dy (dw dgrad)."""
    Cc = z.C
    st.ready(cx)
    if presummed is None:
        sums = cx.f64(2 * Cc)
        L.call("s2r_bn_bwd_reduce", dy.vp(), dy.pitch, 0, z.vp(), z.pitch, 0, _vp(st.mi), _vp(st.ss), act,
               float(drop_p), seed, _vp(cx.seed_dev), z.P, Cc, _vp(sums), cx.stream)
    else:
        sums = presummed
    sync = cx.world > 1 and getattr(bn, "_s2r_sync", False) and not st.frozen
    if sync:
        cx.allreduce(sums)
    dgamma = _vp(grad_of(bn.weight)) if bn.weight is not None and bn.weight.requires_grad else None
    dbeta = _vp(grad_of(bn.bias)) if bn.bias is not None and bn.bias.requires_grad else None
    wH, wW, wpad = win if win is not None else (0, 0, 0)
    L.call("s2r_bn_bwd_apply", dy.vp(), dy.pitch, 0, z.vp(), z.pitch, 0, _vp(st.mi), _vp(st.ss), act,
           float(drop_p), seed, _vp(cx.seed_dev), _vp(sums), 0.0 if st.frozen else st.count, z.P, Cc,
           dx.vp() if dx is not None else None, dx.pitch if dx is not None else 0, 0, dgamma, dbeta, wH, wW, wpad,
           cx.stream)
    return dx


# When a dict {id(bn module): list}: the forward pass appends the raw conv output (Act) every training-mode BatchNorm in
# the dict normalises -- bench.py's multi-rank parity record gathers them to check the synchronised statistics.
BN_PROBE = None


def _probe(bn, z):
    if BN_PROBE is not None and id(bn) in BN_PROBE:
        BN_PROBE[id(bn)].append(z)


# --------------------------------------------------------------------------- composite layers
class ConvBNAct:
    """conv -> BatchNorm -> activation [-> dropout], the BN output materialised in bf16.
    Used where the consumer is a TMA-fed GEMM (ASPP, decoder, domain classifier, pointwise)."""

    def __init__(self, conv, bn, act, drop_p=0.0):
        self.conv, self.bn, self.act, self.drop_p = conv, bn, act, drop_p
        self.stride, self.pad, self.dil = conv.stride[0], conv.padding[0], conv.dilation[0]

    def forward(self, cx, x, out=None, residual=None, count_pad=0):
        w = self.conv.weight
        Cout = w.shape[0]
        OH, OW = conv_out_hw(x.H, x.W, w.shape[2], w.shape[3], self.stride, self.pad, self.dil)
        if cx.inference:
            # inference: BatchNorm with its running statistics (batchnorm.py:50-53) + activation (+ residual) are the
            # convolution's epilogue, act(acc*scale + shift) + residual -- no pre-BN tensor, no bn_apply pass
            st = bn_eval(cx, self.bn)
            y = out if out is not None else cx.new(x.N, OH, OW, Cout)
            conv_fwd(cx, x, w, y, self.stride, self.pad, self.dil, bias=st.ss[Cout:], oscale=st.ss[:Cout], act=self.act,
                     aux=residual, aux_mode=L.AUX_ADD if residual is not None else L.AUX_NONE)
            self.saved = None
            return y
        z = cx.new(x.N, OH, OW, Cout)
        # count_pad: the reference ran this conv on an input padded by count_pad pixels per side
        # (mobilenet.py:62-67): the extra border outputs are exact zeros but count in the statistics
        sums, finish = bn_plan(cx, self.bn, x.N * (OH + 2 * count_pad) * (OW + 2 * count_pad))
        conv_fwd(cx, x, w, z, self.stride, self.pad, self.dil, stats=sums)
        _probe(self.bn, z)
        st = finish()
        p = self.drop_p if (cx.training and cx.dropout) else 0.0
        seed = cx.next_seed() if p > 0 else 0
        y = out if out is not None else cx.new(x.N, OH, OW, Cout)
        bn_apply(cx, z, st, self.act, y, residual, p, seed)
        self.saved = (x, z, st, p, seed)
        return y

    def forward_raw(self, cx, x, count_pad=0):
        """conv + statistics only; returns (z, BNState) for a consumer with a BN prologue.
        x may be a RawNCHW image: the conv then runs as a pointwise GEMM over its patch matrix."""
        w = self.conv.weight
        Cout = w.shape[0]
        OH, OW = conv_out_hw(x.H, x.W, w.shape[2], w.shape[3], self.stride, self.pad, self.dil)
        z = cx.new(x.N, OH, OW, Cout)
        if isinstance(x, RawNCHW):
            assert self.dil == 1
            x = im2col(cx, x, w.shape[2], w.shape[3], self.stride, self.pad)
            sums, finish = bn_plan(cx, self.bn, x.N * OH * OW)
            conv_fwd(cx, x, patch_weight(w), z, stats=sums)
            st = finish()
            self.saved = (x, z, st, 0.0, 0)
            self.patch = True
            return z, st
        self.patch = False
        sums, finish = bn_plan(cx, self.bn, x.N * (OH + 2 * count_pad) * (OW + 2 * count_pad))
        conv_fwd(cx, x, w, z, self.stride, self.pad, self.dil, stats=sums)
        st = finish()
        self.saved = (x, z, st, 0.0, 0)
        return z, st

    def backward(self, cx, dy, need_dx=True, dx=None, dx_accumulate=False):
        """dy: gradient w.r.t. the layer output (Act view, may be a concat slice)."""
        x, z, st, p, seed = self.saved
        dz = cx.new(z.N, z.H, z.W, z.C)
        bn_backward(cx, self.bn, dy, z, st, self.act, dz, p, seed)
        return self.backward_raw(cx, dz, need_dx, dx, dx_accumulate)

    def backward_raw(self, cx, dz, need_dx=True, dx=None, dx_accumulate=False):
        x = self.saved[0]
        w = self.conv.weight
        self.saved = None
        if getattr(self, "patch", False):
            assert not need_dx, "the patch GEMM is used for network inputs only"
            if w.requires_grad:
                conv_wgrad(cx, x, dz, patch_weight(w), grad_param=w)
            return None
        if w.requires_grad:
            conv_wgrad(cx, x, dz, w, self.stride, self.pad, self.dil)
        if not need_dx:
            return None
        if dx is None:
            dx = cx.new(x.N, x.H, x.W, round_up(w.shape[1], 8))
            dx_accumulate = False
        conv_dgrad(cx, dz, w, dx, self.stride, self.pad, self.dil,
                   aux=dx if dx_accumulate else None, aux_mode=L.AUX_ADD if dx_accumulate else L.AUX_NONE)
        return dx


def dw_fwd(cx, x, st_in, in_act, halo_const, w, stride, dil, pad, stats, out_ss=None):
    """Depthwise 3x3 on act(BN(x)); st_in may still be pending: the kernel then finalises it in its prologue.
    out_ss (inference): scale / shift of the BatchNorm that follows; the kernel stores relu6(conv*scale + shift)."""
    Ho, Wo = conv_out_hw(x.H, x.W, 3, 3, stride, pad, dil)
    y = cx.new(x.N, Ho, Wo, x.C)
    p = st_in.take() if st_in is not None else None
    L.call("s2r_dwconv3x3_fwd_bn", x.vp(), C.byref(p) if p is not None else None,
           _vp(st_in.ss) if st_in is not None else None, in_act, 1 if halo_const else 0, _vp(w), y.vp(),
           _vp(stats) if stats is not None else None, _vp(out_ss) if out_ss is not None else None,
           x.N, x.H, x.W, x.C, stride, dil, pad, cx.stream)
    return y


class InvertedResidual:
    """One MobileNetV2 block (modeling/backbone/mobilenet.py:26-68) on the fused kernels:
    expand GEMM (+stats) -> depthwise 3x3 with BN+ReLU6 prologue and halo constant (+stats)
    -> BN+ReLU6 apply -> project GEMM (+stats) -> BN apply (+residual)."""

    def __init__(self, mod):
        seq = mod.conv
        self.mod = mod
        self.expand = len(seq) == 8
        if self.expand:
            self.pw1 = ConvBNAct(seq[0], seq[1], L.ACT_RELU6)
            self.dw, self.bn2 = seq[3], seq[4]
            self.pw2 = ConvBNAct(seq[6], seq[7], L.ACT_NONE)
        else:
            self.pw1 = None
            self.dw, self.bn2 = seq[0], seq[1]
            self.pw2 = ConvBNAct(seq[3], seq[4], L.ACT_NONE)
        self.stride = self.dw.stride[0]
        self.dil = self.dw.dilation[0]
        self.res = mod.use_res_connect

    def forward(self, cx, x, lazy=None):
        """x: block input (Act).  lazy = BNState of a producer whose BN+ReLU6 has not been applied
        to x yet (the stem feeding the expand_ratio==1 block): applied in the dw prologue."""
        d = self.dil
        if self.expand:
            z1, st1 = self.pw1.forward_raw(cx, x, count_pad=d)
            dw_in, st_in, halo = z1, st1, True
        else:
            dw_in, st_in, halo = x, lazy, False
        if cx.inference:
            # inference: BN2 + ReLU6 in the depthwise kernel's store, BN3 (+ residual) in the project conv's epilogue
            st2 = bn_eval(cx, self.bn2)
            y2 = dw_fwd(cx, dw_in, st_in, L.ACT_RELU6, halo, self.dw.weight, self.stride, d, d, None, out_ss=st2.ss)
            self.saved = None
            return self.pw2.forward(cx, y2, residual=x if self.res else None)
        Ho, Wo = conv_out_hw(dw_in.H, dw_in.W, 3, 3, self.stride, d, d)
        sums2, finish2 = bn_plan(cx, self.bn2, dw_in.N * Ho * Wo)
        z2 = dw_fwd(cx, dw_in, st_in, L.ACT_RELU6, halo, self.dw.weight, self.stride, d, d, sums2)
        _probe(self.bn2, z2)
        st2 = finish2()
        y2 = cx.new(z2.N, z2.H, z2.W, z2.C)
        bn_apply(cx, z2, st2, L.ACT_RELU6, y2)
        out = self.pw2.forward(cx, y2, residual=x if self.res else None)
        self.saved = (x, dw_in, st_in, halo, z2, st2)
        return out

    def backward(self, cx, dout, need_dx=True):
        """Returns (dx, bwd_sums_for_lazy_producer).  For the expand_ratio==1 block dx is the
        act'-masked gradient w.r.t. the producer's BN output and the sums feed its BN backward."""
        x, dw_in, st_in, halo, z2, st2 = self.saved
        self.saved = None
        d = self.dil
        dy2 = self.pw2.backward(cx, dout)                      # BN3 bwd, wgrad, dgrad -> [N,Ho,Wo,Ch]
        dz2 = cx.new(z2.N, z2.H, z2.W, z2.C)
        bn_backward(cx, self.bn2, dy2, z2, st2, L.ACT_RELU6, dz2)
        # g in the unextended layout: the reference's padded border only enters the BN-backward sums
        g = cx.new(dw_in.N, dw_in.H, dw_in.W, dw_in.C)
        masked = st_in is not None
        bsums = cx.f64(2 * dw_in.C) if (masked and not st_in.frozen) else None
        L.call("s2r_dwconv3x3_bwd", dz2.vp(), _vp(self.dw.weight), dw_in.vp(), _vp(st_in.ss) if masked else None,
               _vp(st_in.mi) if masked else None, L.ACT_RELU6, 1 if halo else 0, 1, g.vp(),
               _vp(bsums) if bsums is not None else None,
               _vp(grad_of(self.dw.weight)) if self.dw.weight.requires_grad else None, dw_in.N, dw_in.H, dw_in.W,
               dw_in.C, self.stride, d, d, cx.stream)
        if not self.expand:
            return g, bsums
        # BN1 backward on the padded domain, interior written as dz1
        if st_in.frozen:
            bsums = cx.f64(2 * dw_in.C)
        dz1 = cx.new(dw_in.N, dw_in.H, dw_in.W, dw_in.C)
        bn_backward(cx, self.pw1.bn, g, dw_in, st_in, L.ACT_NONE, dz1, presummed=bsums)
        if self.res:
            dx = self.pw1.backward_raw(cx, dz1, True, dx=dout, dx_accumulate=True)
        else:
            dx = self.pw1.backward_raw(cx, dz1, need_dx)
        return dx, None
