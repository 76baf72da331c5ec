"""The three loop bodies of the reference as callables over the drop-in modules:
  AdaptStep    train_adapt.py:126-181  output-space adaptation (AdaptSegNet): G + FCDiscriminator
  FeatureStep  train.py:163-216        feature adaptation (FCN in the wild): domain classifier
  ValStep      val_adapt.py:122-135    eval forward + Evaluator confusion matrix
Order of forward/backward passes, requires_grad toggling, loss definitions (including the
batch-axis softmax and the un-weighted adversarial loss) and optimizer steps follow the
reference; losses stay on the device (no per-step .item()).

Data parallel (one process per GPU): each rank runs the step on its shard of the batch; BN
statistics are all-reduced inside the engine when the model was built with sync_bn=True, and the
flat gradient buffers are all-reduced (averaged) before the optimizer steps.  The cross-entropy mean is the
global-batch one (functional.GLOBAL_BATCH_MEAN).  Deviation from the single-process DataParallel of the
reference, stated in DESIGN.md: F.softmax(dim=0) is taken over the rank-local batch unless
functional.GLOBAL_SOFTMAX0 (S2R_GLOBAL_SOFTMAX0=1) asks for the gathered-batch semantics.
"""
import os

import torch

from . import _lib as _L
from .engine import seed_counter, prepack_weights, PEER, COMM_CHANNEL, WEIGHT_EPOCH
from . import functional as _fn
from .functional import softmax_dim0, bce_with_logits
from .optim import FusedSGD, FusedAdam
from .utils.loss import SegmentationLosses, DomainLosses
from .utils.lr_scheduler import LR_Scheduler
from .utils.metrics import Evaluator


def _backward_ce_deferred(loss, model):
    """loss.backward() for a loss whose cross-entropy gradient goes straight into `model`'s (DeepLab) backward: the mean
    reduction's factor is folded into the up-sampling backward kernel instead of one more pass over the N x 19 x H x W
    gradient (functional.DEFER_CE_SCALE).  Any other model takes the ordinary path."""
    if type(model).__name__ != 'DeepLab':
        loss.backward()
        return
    _fn.DEFER_CE_SCALE[0] = True
    try:
        loss.backward()
    finally:
        _fn.DEFER_CE_SCALE[0] = False
    if _fn.PENDING_SCALE:
        _fn.PENDING_SCALE.clear()
        raise RuntimeError("deferred cross-entropy scale was not consumed by the model's backward")


def check_peer_exchange():
    """Raise if the NVLink peer exchange of the BN statistics (csrc/comm.cu) hit its bounded wait -- a peer died, or
    the exchange kernels of the step's two streams were not co-resident on some rank and cross-waited.  The flag
    lives in mapped host memory, so this is a plain host read (no device synchronisation): the training steps poll
    it once per iteration, i.e. an error surfaces one step late at most instead of training on on partial sums."""
    if PEER["world"] > 1:
        e = _L.lib().s2r_comm_error()
        if e != 0:
            raise _L.S2RError("BN-statistics peer exchange timed out (code %d): a rank is gone, or the two exchange "
                              "channels of the two-stream step cross-waited -- rerun with S2R_OVERLAP=0 (one stream, "
                              "one channel) or S2R_COMM=nccl" % e)


def _shared_d_forward(model_D, logits):
    """model_D.forward_softmax0_shared when the fused input stage covers the shape (and S2R_SHARE_D_FWD != 0), else None."""
    fn = getattr(model_D, "forward_softmax0_shared", None)
    if (fn is None or os.environ.get("S2R_SHARE_D_FWD", "1") == "0" or logits.dim() != 4 or logits.shape[0] > 8
            or logits.shape[1] > 64 or logits.shape[2] % 2 or logits.shape[3] % 2):
        return None
    return fn


def _disc_on_softmax0(model_D, logits):
    """model_D(F.softmax(logits, dim=0)) (train_adapt.py:151,166,174); the discriminator's fused input stage
    when it covers the shape, the two separate calls otherwise."""
    fused = getattr(model_D, "forward_softmax0", None)
    if (fused is not None and logits.dim() == 4 and logits.shape[0] <= 8 and logits.shape[1] <= 64
            and logits.shape[2] % 2 == 0 and logits.shape[3] % 2 == 0):
        return fused(logits)
    return model_D(softmax_dim0(logits))


class _StagedInputs(object):
    """Input pipelining for a captured step (AdaptStep / FeatureStep): the next step's host tensors are copied into
    device staging buffers on a copy stream while the current graph replay runs; the replay then starts with a
    device-to-device copy into the graph's static inputs (what a prefetching data loader does around
    train_adapt.py:126-129 / train.py:163-172)."""

    def stage(self, src_image, src_label, tgt_image):
        """Start the asynchronous host->device copy of one step's inputs (pinned host tensors)."""
        dev = self._static[0].device
        if getattr(self, "_stage_bufs", None) is None:
            self._stage_bufs = tuple(torch.empty_like(t) for t in self._static)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged = None
        # the staging buffers may still be read by the previous replay's device-to-device copy (but the copy
        # must NOT wait for the replay itself, which is what it is meant to overlap)
        if getattr(self, "_d2d_done", None) is not None:
            self._copy_stream.wait_event(self._d2d_done)
        with torch.cuda.stream(self._copy_stream):
            for b, t in zip(self._stage_bufs, (src_image, src_label, tgt_image)):
                b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged = ev

    def replay_staged(self, i=0, epoch=0):
        """Replay the captured step on the inputs of the last stage() call."""
        assert getattr(self, "_staged", None) is not None, "stage() first"
        cur = torch.cuda.current_stream(self._static[0].device)
        cur.wait_event(self._staged)
        self._staged = None
        for st, b in zip(self._static, self._stage_bufs):
            st.copy_(b, non_blocking=True)
        self._d2d_done = torch.cuda.Event()
        self._d2d_done.record(cur)
        return self.replay(*self._static, i=i, epoch=epoch)


class AdaptStep(_StagedInputs):
    def __init__(self, model, model_D, lr=5e-4, momentum=0.9, weight_decay=5e-4, nesterov=False,
                 lr_scheduler='poly', epochs=200, iters_per_epoch=1000, class_weight=None, loss_type='ce'):
        self.model, self.model_D = model, model_D
        train_params = [{'params': list(model.get_1x_lr_params()), 'lr': lr},
                        {'params': list(model.get_10x_lr_params()), 'lr': lr * 10}]
        self.optimizer = FusedSGD(train_params, lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov)
        self.optimizer_D = FusedAdam(model_D.parameters(), lr=1e-4, betas=(0.9, 0.99))
        self.criterion = SegmentationLosses(weight=class_weight).build_loss(mode=loss_type)
        self.scheduler = LR_Scheduler(lr_scheduler, lr, epochs, iters_per_epoch)
        self.source_label, self.target_label = 0, 1

    def __call__(self, src_image, src_label, tgt_image, i=0, epoch=0):
        """One eager step.  With capture() done, use replay() instead."""
        check_peer_exchange()
        self.scheduler(self.optimizer, i, epoch)
        self.scheduler(self.optimizer_D, i, epoch)
        self.optimizer.advance()
        self.optimizer_D.advance()
        return self._device_step(src_image, src_label, tgt_image)

    def capture(self, src_image, src_label, tgt_image, warmup=2):
        """Capture the whole step (both generator passes, three discriminator passes, all backward
        passes, gradient all-reduce and both optimizer kernels) into one CUDA graph.  Inputs are copied
        into static buffers before each replay; learning rates and Adam bias corrections reach the
        device through optimizer.advance() ahead of every replay (values snapshotted at call time, optim.py),
        dropout masks through the device seed counter."""
        dev = src_image.device
        self._static = tuple(torch.empty_like(t) for t in (src_image, src_label, tgt_image))
        for st, t in zip(self._static, (src_image, src_label, tgt_image)):
            st.copy_(t)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for k in range(warmup):
                self(*self._static, i=0, epoch=0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the captured step holds launch() only: the host half (scheduler + advance(): step count, learning rates
        # and bias corrections into the device buffers) runs ahead of every replay, not here -- the capture itself
        # executes nothing and must not count as a step
        # the job table is uploaded here and stays alive with the graph; its launch is captured below
        self._pack_tables = prepack_weights(None, build_only=True)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_out = self._device_step(*self._static)
        return self

    def replay(self, src_image, src_label, tgt_image, i=0, epoch=0):
        check_peer_exchange()
        self.scheduler(self.optimizer, i, epoch)
        self.scheduler(self.optimizer_D, i, epoch)
        self.optimizer.advance()
        self.optimizer_D.advance()
        for st, t in zip(self._static, (src_image, src_label, tgt_image)):
            if st.data_ptr() != t.data_ptr():
                st.copy_(t, non_blocking=True)
        self._graph.replay()
        WEIGHT_EPOCH[0] += 1     # the captured optimizer kernels changed the parameters (engine.packed_weight stamps)
        return self._static_out

    def _device_step(self, src_image, src_label, tgt_image):
        model, model_D = self.model, self.model_D
        seed_counter(src_image.device).add_(1)
        self.optimizer.zero_grad()
        self.optimizer_D.zero_grad()
        # all bf16 filter copies invalidated by the previous optimizer step, in one launch
        prepack_weights(torch.cuda.current_stream(src_image.device).cuda_stream)
        if self._two_streams(src_image.device):
            loss_seg, loss_adv, loss_D_src, loss_D_tgt = self._passes_two_streams(src_image, src_label, tgt_image)
        else:
            # ---- train G; don't accumulate grads in D (train_adapt.py:140-155)
            for p in model_D.parameters():
                p.requires_grad = False
            src_output = model(src_image)
            loss_seg = self.criterion(src_output, src_label)
            _backward_ce_deferred(loss_seg, model)
            tgt_output = model(tgt_image)
            shared = _shared_d_forward(model_D, tgt_output)
            if shared is not None:
                # D(softmax(tgt_output)) is evaluated ONCE for the adversarial pass (:151) and the discriminator's
                # training pass (:174): same tensor, same weights (FCDiscriminator.forward_softmax0_shared)
                D_out, attach_tgt = shared(tgt_output)
            else:
                D_out, attach_tgt = _disc_on_softmax0(model_D, tgt_output), None
            loss_adv = bce_with_logits(D_out, self.source_label)
            loss_adv.backward()
            # ---- train D (train_adapt.py:160-178)
            for p in model_D.parameters():
                p.requires_grad = True
            src_output = src_output.detach()
            loss_D_src = bce_with_logits(_disc_on_softmax0(model_D, src_output), self.source_label)
            loss_D_src.backward()
            tgt_output = tgt_output.detach()
            D_tgt = attach_tgt() if attach_tgt is not None else _disc_on_softmax0(model_D, tgt_output)
            loss_D_tgt = bce_with_logits(D_tgt, self.target_label)
            loss_D_tgt.backward()
        if not getattr(self, "_g_reduced", False):
            self.optimizer.all_reduce_grads()
        self._g_reduced = False
        self.optimizer_D.all_reduce_grads()
        self.optimizer.launch()
        self.optimizer_D.launch()
        return {'loss_seg': loss_seg.detach(), 'loss_adv': loss_adv.detach(), 'loss_D_src': loss_D_src.detach(),
                'loss_D_tgt': loss_D_tgt.detach()}


def _adapt_two_streams(self, dev):
    """Two streams need two independent sequences of BN-statistics exchanges: single process, no synchronised BN, or the
    peer-memory exchange with its two channels."""
    if os.environ.get("S2R_OVERLAP", "1") == "0":
        return False
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and getattr(self.model, "_s2r_has_sync_bn", False):
        # with the NVLink peer exchange each stream has its own exchange channel (csrc/comm.cu); NCCL collectives of one
        # communicator cannot be issued from two streams at once
        return PEER["world"] == dist.get_world_size()
    return True


def _adapt_passes_two_streams(self, src_image, src_label, tgt_image):
    """train_adapt.py:140-178 with the same arithmetic on two streams.  The passes of the step form two chains that only
    meet in the gradient buffers:  A = G(src) forward -> CE -> G backward -> D training passes (src, tgt);
    B = G(tgt) forward -> D (frozen) -> BCE -> backward through D and G.  B's forward runs beside A's backward, B's
    backward (which accumulates into the same generator gradients) starts when A's backward has finished and runs
    beside the discriminator's training passes.  Each chain is a sequence of ~400 kernels most of which fill a fraction
    of the GPU, so the chains interleave well.  Autograd runs every backward node on the stream of its forward."""
    model, model_D = self.model, self.model_D
    dev = src_image.device
    A = torch.cuda.current_stream(dev)
    if getattr(self, "_stream_B", None) is None:
        self._stream_B = torch.cuda.Stream(device=dev)
    B = self._stream_B
    src_output = model(src_image)
    loss_seg = self.criterion(src_output, src_label)
    B.wait_stream(A)                       # BN running statistics: G(src) forward before G(tgt) forward
    for p in model_D.parameters():
        p.requires_grad = False            # train G: no gradients in D (train_adapt.py:140-141)
    COMM_CHANNEL[0] = 1                    # stream B's BN exchanges form their own sequence
    try:
        with torch.cuda.stream(B):
            tgt_output = model(tgt_image)
            shared = _shared_d_forward(model_D, tgt_output)
            if shared is not None:
                D_out, attach_tgt = shared(tgt_output)     # one evaluation for the adversarial AND the training pass
            else:
                D_out, attach_tgt = _disc_on_softmax0(model_D, tgt_output), None
            loss_adv = bce_with_logits(D_out, self.source_label)
            fwd_B = torch.cuda.Event()
            fwd_B.record(B)
    finally:
        COMM_CHANNEL[0] = 0
    three = os.environ.get("S2R_STREAMS", "2") == "3"   # measured: 17.59 ms with, 17.45 ms without (profiles/r2_notes.md)
    if three:
        # third chain C: the discriminator's training pass on the (detached) source prediction needs nothing but
        # G(src)'s forward and D's weights, so it runs beside G(src)'s backward and G(tgt)'s forward instead of after
        # them; likewise its pass on the target prediction runs beside the adversarial backward.  D's parameter
        # gradients come only from these two passes (requires_grad is off while the adversarial pass is issued,
        # train_adapt.py:140-141,158-159), both on C, in the reference's order.
        if getattr(self, "_stream_C", None) is None:
            self._stream_C = torch.cuda.Stream(device=dev)
        Cs = self._stream_C
        Cs.wait_stream(A)                  # G(src) forward (and the zeroed gradient buffers) are ready
        for p in model_D.parameters():
            p.requires_grad = True
        with torch.cuda.stream(Cs):
            src_det = src_output.detach()
            loss_D_src = bce_with_logits(_disc_on_softmax0(model_D, src_det), self.source_label)
            loss_D_src.backward()
        for p in model_D.parameters():
            p.requires_grad = False
    _backward_ce_deferred(loss_seg, model)  # on A, beside B's forward
    # the two generator backward passes run one after the other: overlapping them as well (all gradient accumulation
    # is atomic, so it would be legal) measured no gain -- 17.8 ms either way, the GPU is full by then
    B.wait_stream(A)
    COMM_CHANNEL[0] = 1
    try:
        with torch.cuda.stream(B):
            loss_adv.backward()
            # data parallel: the generator's gradient is complete here (A's backward was waited for above), the
            # discriminator's training passes on A do not touch it -- its all-reduce runs on B beside them instead of
            # after them.  The discriminator's all-reduce follows on A after A has waited for B (end of this function):
            # the two collectives of the one communicator stay in a fixed order on every rank.
            self.optimizer.all_reduce_grads()
            self._g_reduced = True
    finally:
        COMM_CHANNEL[0] = 0
    for p in model_D.parameters():
        p.requires_grad = True             # train D (train_adapt.py:158-159)
    if three:
        Cs.wait_event(fwd_B)
        with torch.cuda.stream(Cs):
            tgt_det = tgt_output.detach()
            D_tgt = attach_tgt() if attach_tgt is not None else _disc_on_softmax0(model_D, tgt_det)
            loss_D_tgt = bce_with_logits(D_tgt, self.target_label)
            loss_D_tgt.backward()
        A.wait_stream(Cs)
        A.wait_stream(B)
        return loss_seg, loss_adv, loss_D_src, loss_D_tgt
    src_output = src_output.detach()
    loss_D_src = bce_with_logits(_disc_on_softmax0(model_D, src_output), self.source_label)
    loss_D_src.backward()
    A.wait_event(fwd_B)
    tgt_output = tgt_output.detach()
    D_tgt = attach_tgt() if attach_tgt is not None else _disc_on_softmax0(model_D, tgt_output)
    loss_D_tgt = bce_with_logits(D_tgt, self.target_label)
    loss_D_tgt.backward()
    A.wait_stream(B)
    return loss_seg, loss_adv, loss_D_src, loss_D_tgt


AdaptStep._two_streams = _adapt_two_streams
AdaptStep._passes_two_streams = _adapt_passes_two_streams


class FeatureStep(_StagedInputs):
    def __init__(self, backbone_model, assp_model, y_model, d_model, lr=5e-4, optimizer='Adam', momentum=0.9,
                 weight_decay=5e-4, nesterov=False, lr_scheduler='poly', epochs=200, iters_per_epoch=1000):
        self.f, self.a, self.y, self.d = backbone_model, assp_model, y_model, d_model
        f_params = list(backbone_model.parameters()) + list(assp_model.parameters())
        y_params = list(y_model.parameters())
        d_params = list(d_model.parameters())
        # train.py:63-82: task (f+y), d, d_inv (f again); c_optimizer exists there but never steps
        if optimizer == 'SGD':
            mk = lambda ps: FusedSGD(ps, lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov)  # noqa: E731
        elif optimizer == 'Adam':
            mk = lambda ps: FusedAdam(ps, lr=lr)  # noqa: E731
        else:
            raise NotImplementedError
        self.task_optimizer = mk(f_params + y_params)
        self.d_optimizer = mk(d_params)
        self.d_inv_optimizer = _SharedGradOptimizer(mk, f_params, self.task_optimizer)
        self.task_loss = SegmentationLosses().build_loss('ce')
        self._domain_losses = DomainLosses()
        self.domain_loss = self._domain_losses.build_loss()
        self.scheduler = LR_Scheduler(lr_scheduler, lr, epochs, iters_per_epoch)

    def _forward(self, image):
        high0, low = self.f(image)
        high = self.a(high0)
        # F.interpolate(self.y_model(high, low), image.size()[2:], mode='bilinear', align_corners=True)
        # (train.py:184,194): the up-sampling is fused into the decoder's output conversion (resize.cu)
        out = self.y(high, low, size=image.size()[2:])
        return out, self.d(high)

    def _optimizers(self):
        return (self.task_optimizer, self.d_optimizer, self.d_inv_optimizer)

    def __call__(self, src_image, src_label, tgt_image=None, i=0, epoch=0):
        """One eager step.  tgt_image=None: the single-domain branch of the loop (args.dataset == 'gtav',
        train.py:164-165,205-210) -- task loss only, only task_optimizer steps."""
        check_peer_exchange()
        for o in self._optimizers():
            self.scheduler(o, i, epoch)
        return self._device_step(src_image, src_label, tgt_image)

    def _advance(self, i, epoch):
        """Host half of a captured step: learning rates / Adam bias corrections into the pinned buffers."""
        for o in self._optimizers():
            self.scheduler(o, i, epoch)
        self.task_optimizer.advance()
        self.d_optimizer.advance()
        self.d_inv_optimizer.advance()

    def capture(self, src_image, src_label, tgt_image, warmup=2):
        """Capture the whole feature-adaptation step (two forwards through backbone + ASPP + decoder + domain
        classifier, one backward, gradient all-reduce, the three optimizer kernels) into one CUDA graph -- the
        counterpart of AdaptStep.capture.  The domain accuracy stays on the device (no .item() inside the step)."""
        dev = src_image.device
        self._static = tuple(torch.empty_like(t) for t in (src_image, src_label, tgt_image))
        for st, t in zip(self._static, (src_image, src_label, tgt_image)):
            st.copy_(t)
        self._domain_losses.device_acc = True
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for k in range(warmup):
                self(*self._static, i=0, epoch=0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._pack_tables = prepack_weights(None, build_only=True)    # see AdaptStep.capture
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_out = self._device_step(*self._static, launch_only=True)
        return self

    def replay(self, src_image, src_label, tgt_image, i=0, epoch=0):
        check_peer_exchange()
        self._advance(i, epoch)
        for st, t in zip(self._static, (src_image, src_label, tgt_image)):
            if st.data_ptr() != t.data_ptr():
                st.copy_(t, non_blocking=True)
        self._graph.replay()
        WEIGHT_EPOCH[0] += 1
        return self._static_out

    def _device_step(self, src_image, src_label, tgt_image, launch_only=False):
        seed_counter(src_image.device).add_(1)
        self.task_optimizer.zero_grad()
        self.d_optimizer.zero_grad()
        prepack_weights(torch.cuda.current_stream(src_image.device).cuda_stream)
        src_output, src_d_pred = self._forward(src_image)
        # the task loss's gradient goes straight into the decoder's fused up-sampling backward, which takes the mean
        # reduction's factor along (functional.DEFER_NEXT): no scaling pass over the N x 19 x H x W gradient
        _fn.DEFER_NEXT[0] = type(self.y).__name__ == 'Decoder'
        try:
            task_loss = self.task_loss(src_output, src_label)
        finally:
            _fn.DEFER_NEXT[0] = False
        if tgt_image is None:
            # train.py:205-210: the domain classifier ran on the source features (:187, its BatchNorm statistics
            # moved) but only the task loss is back-propagated and only task_optimizer steps
            del src_d_pred
            task_loss.backward()
            self.task_optimizer.all_reduce_grads()
            if launch_only:
                self.task_optimizer.launch()
            else:
                self.task_optimizer.step()
            zero = torch.zeros((), device=src_image.device)
            return {'task_loss': task_loss.detach(), 'd_loss': zero, 'd_inv_loss': zero, 'd_acc': 0}
        _, tgt_d_pred = self._forward(tgt_image)
        d_loss, d_acc = self.domain_loss(src_d_pred, tgt_d_pred)
        d_inv_loss, _ = self.domain_loss(tgt_d_pred, src_d_pred)
        loss = task_loss + d_loss + d_inv_loss
        loss.backward()
        if _fn.PENDING_SCALE:
            _fn.PENDING_SCALE.clear()
            raise RuntimeError("deferred cross-entropy scale was not consumed by the decoder's backward")
        self.task_optimizer.all_reduce_grads()
        self.d_optimizer.all_reduce_grads()
        for o in self._optimizers():
            if launch_only:
                o.launch()     # the host half (advance) ran before the capture / runs before every replay
            else:
                o.step()
        return {'task_loss': task_loss.detach(), 'd_loss': d_loss.detach(), 'd_inv_loss': d_inv_loss.detach(),
                'd_acc': d_acc}


class _SharedGradOptimizer(object):
    """A second optimizer over parameters whose gradients already live in another optimizer's
    flat buffer (train.py:71-73,204: d_inv_optimizer steps backbone+ASPP again with the same grads)."""

    def __init__(self, mk, params, owner):
        grads = [p.grad for p in params]
        self.inner = mk(params)          # re-points p.grad at its own flat buffer
        self.params, self.owner = params, owner
        for p, g in zip(params, grads):  # restore the owner's views: both optimizers read the same grads
            p.grad = g
        self.param_groups = self.inner.param_groups
        self.inner._tables = None

    def step(self):
        if self.inner._tables is None:
            # build tables against the owner's gradient views
            inner = self.inner
            inner.flat_grad = self.owner.flat_grad
            lookup = {id(p): off for p, off, n in self.owner._views}
            inner._views = [(p, lookup[id(p)], p.numel()) for p in self.params]
            # state buffers stay private but are indexed with the owner's offsets: size them alike
            inner.state_bufs = [torch.zeros_like(self.owner.flat_grad) for _ in range(inner.n_state)]
            inner._build_tables()
        self.inner.grad_scale = self.owner.grad_scale
        self.inner.step()

    def _prepare(self):
        if self.inner._tables is None:
            inner = self.inner
            inner.flat_grad = self.owner.flat_grad
            lookup = {id(p): off for p, off, n in self.owner._views}
            inner._views = [(p, lookup[id(p)], p.numel()) for p in self.params]
            inner.state_bufs = [torch.zeros_like(self.owner.flat_grad) for _ in range(inner.n_state)]
            inner._build_tables()
        self.inner.grad_scale = self.owner.grad_scale

    def advance(self):
        self.inner.advance()

    def state_dict(self):
        self._prepare()
        return self.inner.state_dict()

    def load_state_dict(self, sd):
        self._prepare()
        self.inner.load_state_dict(sd)

    def launch(self):
        self._prepare()
        self.inner.launch()

    def zero_grad(self):
        pass


FUSED_VAL = os.environ.get("S2R_FUSED_VAL", "1") != "0"   # ValStep: fused up-sampling + argmax + confusion matrix


class ValStep(object):
    """val_adapt.py:122-135 with argmax + confusion matrix fused on the device."""

    def __init__(self, model, num_class=19):
        self.model = model
        self._evaluator = Evaluator(num_class)
        self.criterion = SegmentationLosses().build_loss('ce')

    @property
    def evaluator(self):
        """The Evaluator, with every graph lane ordered before the current stream (so that reading a metric sees all
        replayed images)."""
        self.finish()
        return self._evaluator

    @torch.no_grad()
    def __call__(self, image, target, with_loss=False):
        net = getattr(self.model, "module", self.model)     # nn.DataParallel wrapper of the reference's scripts
        if not with_loss and FUSED_VAL and hasattr(net, "forward_confusion") and not net.training:
            # up-sampling + argmax + histogram in one launch on the low-resolution logits (no fp32 logits tensor)
            net.forward_confusion(image, target, self._evaluator)
            return None
        output = self.model(image)
        loss = self.criterion(output, target) if with_loss else None
        self._evaluator.add_batch_logits(target, output)
        return loss

    @torch.no_grad()
    def capture(self, image, target, lanes=1):
        """Capture forward + fused argmax/confusion matrix into one CUDA graph (a batch-1 eval forward is ~200
        short launches: issued from Python they starve the GPU).  replay() copies the inputs into the static
        buffers; the confusion matrix keeps accumulating on the device.
        lanes > 1: that many copies of the graph, each with its own static buffers and stream; successive images go
        to the lanes round-robin and overlap on the GPU (a batch-1 forward is a chain of kernels that fill a fraction
        of the SMs; the counts are integer atomics, so the result does not depend on the interleaving).  Call
        finish() before reading the evaluator."""
        dev = image.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self(image, target)          # warm-up: weight packing, lazy buffers
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self._lanes, self._next = [], 0
        for lane in range(max(1, lanes)):
            static = (torch.empty_like(image), torch.empty_like(target))
            for st, t in zip(static, (image, target)):
                st.copy_(t)
            stream = torch.cuda.Stream(device=dev) if lanes > 1 else None
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(dev)
            if stream is None:
                with torch.cuda.graph(graph):
                    self(*static)
            else:
                with torch.cuda.graph(graph, stream=stream):
                    self(*static)
            self._lanes.append((static, graph, stream))
        self._static, self._graph = self._lanes[0][0], self._lanes[0][1]
        self._evaluator.reset()           # drop the warm-up counts
        return self

    @torch.no_grad()
    def replay(self, image, target):
        # the graph holds no pack kernels (capture() packed the filters in its eager warm-up): bf16 filter copies
        # made stale since then -- by an optimizer step of a training graph, a checkpoint load, any in-place update
        # -- are refreshed here, in one launch on the current stream, which every lane waits for below
        prepack_weights(torch.cuda.current_stream(image.device).cuda_stream)
        static, graph, stream = self._lanes[self._next]
        self._next = (self._next + 1) % len(self._lanes)
        consumed = torch.cuda.Event()
        if stream is None:
            for st, t in zip(static, (image, target)):
                if st.data_ptr() != t.data_ptr():
                    st.copy_(t, non_blocking=True)
            consumed.record(torch.cuda.current_stream(image.device))
            graph.replay()
            return consumed
        stream.wait_stream(torch.cuda.current_stream(image.device))     # the caller's tensors are ready
        with torch.cuda.stream(stream):
            for st, t in zip(static, (image, target)):
                if st.data_ptr() != t.data_ptr():
                    st.copy_(t, non_blocking=True)
            consumed.record(stream)
            graph.replay()
        return consumed     # the caller may overwrite image / target once this event has completed (input pipelining)

    def all_reduce(self, group=None):
        """Data-parallel validation: every rank has run its shard of the images (rank r takes images r, r + world,
        ...); sums the confusion matrices over the ranks (Evaluator.all_reduce) and returns the evaluator."""
        return self.evaluator.all_reduce(group)

    def finish(self):
        """Order every lane before the current stream (call before reading the evaluator)."""
        for static, graph, stream in getattr(self, "_lanes", []):
            if stream is not None:
                torch.cuda.current_stream(static[0].device).wait_stream(stream)
