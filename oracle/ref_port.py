"""CPU oracle: a functional fp32 restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs as the checker and the reported CPU baseline; never by the
product path (synthetic-to-real-semantic-segmentation_b200/), which has no CPU fallback.

The reference (haofengsiji/synthetic-to-real-semantic-segmentation) is Python on top of an
un-vendored, un-pinned PyTorch, so the arithmetic is restated with
torch.nn.functional ops in fp32 on the CPU, driven by a plain {name: tensor} state dict with the
reference's key names.  Each function cites the reference lines it follows.

Pinning: the reference ships no golden vectors or tests for this path.  tests/golden/make_golden.py
imports the real reference modules from /root/reference, checks this port against them (forward,
loss, parameter gradients; the reference's own Trainer.training of train_adapt.py and
Trainer.validation of val_adapt.py compiled unmodified out of the scripts and run on the CPU; the
feature-adaptation step of train.py on the reference's modules; the synchronised-BatchNorm protocol
executed through the reference's SyncMaster) and writes the fixtures under tests/golden/ that
tests/test_oracle.py re-checks on every run -- parity is pinned on reference outputs generated in
the build container, not on reference-owned vectors.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# Optional hooks used by tests/tools/layer_trace.py and the bf16-emulation parity tests:
#   Q      rounds a tensor where the B200 path stores it in bf16 (identity = the fp32 reference);
#   TRACE  when a list, receives (name, tensor) for the intermediate activations.
Q = None
TRACE = None


def _q(t):
    return t if Q is None else Q(t)


# bf16-storage emulation (tests/emul.py) of the B200 INFERENCE path: in eval mode under torch.no_grad() the BatchNorm
# (+ activation, + residual) is the epilogue of the producing convolution, so the pre-BatchNorm output of the
# depthwise / project / ASPP / decoder / domain-classifier convolutions is never stored (hence never rounded).
Q_FOLD_EVAL = False


def _qz(t, cfg):
    """Rounding point of a pre-BatchNorm convolution output."""
    return t if (Q_FOLD_EVAL and not cfg.training) else _q(t)


def _tr(name, t):
    if TRACE is not None:
        TRACE.append((name, t.detach()))
    return t


MNV2_TABLE = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
              (6, 320, 1, 1))


def mnv2_plan(output_stride=16):
    """(inp, oup, stride, dilation, expand) per block -- modeling/backbone/mobilenet.py:78-109."""
    plan, inp, cur, rate = [], 32, 2, 1
    for t, c, n, s in MNV2_TABLE:
        if cur == output_stride:
            stride, dil = 1, rate
            rate *= s
        else:
            stride, dil = s, 1
            cur *= s
        for i in range(n):
            plan.append((inp, c, stride if i == 0 else 1, dil, t))
            inp = c
    return plan


class BNCfg:
    """How BatchNorm is evaluated: training (batch statistics, running-stat update) or eval."""

    def __init__(self, training=True, momentum=0.1, eps=1e-5, sync_clamp=False):
        self.training, self.momentum, self.eps, self.sync_clamp = training, momentum, eps, sync_clamp


def batch_norm(sd, pre, x, cfg):
    """modeling/sync_batchnorm/batchnorm.py:48-53 (F.batch_norm branch) or, with sync_clamp, the
    data-parallel branch :55-78 + _compute_mean_std :113-125 evaluated over the whole batch."""
    w, b = sd[pre + '.weight'], sd[pre + '.bias']
    rm, rv = sd[pre + '.running_mean'], sd[pre + '.running_var']
    if not (cfg.training and cfg.sync_clamp):
        return F.batch_norm(x, rm, rv, w, b, cfg.training, cfg.momentum, cfg.eps)
    C = x.shape[1]
    xv = x.transpose(0, 1).reshape(C, -1)
    size = xv.shape[1]
    assert size > 1, 'BatchNorm computes unbiased standard-deviation, which requires size > 1.'
    s, ss = xv.sum(1), (xv ** 2).sum(1)
    mean = s / size
    sumvar = ss - s * mean
    with torch.no_grad():
        rm.mul_(1 - cfg.momentum).add_(cfg.momentum * mean)
        rv.mul_(1 - cfg.momentum).add_(cfg.momentum * sumvar / (size - 1))
    inv_std = (sumvar / size).clamp(cfg.eps) ** -0.5
    shp = (1, C, 1, 1)
    return (x - mean.view(shp)) * (inv_std * w).view(shp) + b.view(shp)


def fixed_padding(x, dilation):
    """modeling/backbone/mobilenet.py:17-23 for kernel_size 3: pad by `dilation` on every side."""
    return F.pad(x, (dilation,) * 4)


def inverted_residual(sd, pre, x, inp, oup, stride, dil, expand, cfg):
    """modeling/backbone/mobilenet.py:26-68.  The block input is padded BEFORE the 1x1 expand."""
    h = fixed_padding(x, dil)
    i = 0
    if expand != 1:
        h = _q(F.conv2d(h, _q(sd[pre + '.conv.0.weight'])))
        h = F.relu6(batch_norm(sd, pre + '.conv.1', h, cfg))       # consumed by the dw prologue: not stored
        i = 3
    hidden = h.shape[1]
    h = _qz(F.conv2d(h, sd['%s.conv.%d.weight' % (pre, i)], None, stride, 0, dil, hidden), cfg)
    h = _q(F.relu6(batch_norm(sd, '%s.conv.%d' % (pre, i + 1), h, cfg)))
    h = _qz(F.conv2d(h, _q(sd['%s.conv.%d.weight' % (pre, i + 3)])), cfg)
    h = batch_norm(sd, '%s.conv.%d' % (pre, i + 4), h, cfg)
    return _q(x + h) if (stride == 1 and inp == oup) else _q(h)


def mobilenet_forward(sd, x, cfg, output_stride=16, pre='features'):
    """modeling/backbone/mobilenet.py:119-122 -> (high, low_level_feat)."""
    h = _q(F.conv2d(_q(x), _q(sd[pre + '.0.0.weight']), None, 2, 1))
    h = F.relu6(batch_norm(sd, pre + '.0.1', h, cfg))
    low = None
    for k, (inp, oup, stride, dil, t) in enumerate(mnv2_plan(output_stride), start=1):
        h = _tr('block%d' % k, inverted_residual(sd, '%s.%d' % (pre, k), h, inp, oup, stride, dil, t, cfg))
        if k == 3:
            low = h
    return h, low


def aspp_forward(sd, x, cfg, output_stride=16, pre='', drop=None):
    """modeling/assp.py:65-78."""
    dils = {16: (1, 6, 12, 18), 8: (1, 12, 24, 36)}[output_stride]
    outs = []
    for k, d in enumerate(dils, start=1):
        w = sd['%saspp%d.atrous_conv.weight' % (pre, k)]
        h = _qz(F.conv2d(x, _q(w), None, 1, 0 if k == 1 else d, d), cfg)
        outs.append(_q(F.relu(batch_norm(sd, '%saspp%d.bn' % (pre, k), h, cfg))))
    g = _q(F.adaptive_avg_pool2d(x, 1))
    g = _qz(F.conv2d(g, _q(sd[pre + 'global_avg_pool.1.weight'])), cfg)
    g = _q(F.relu(batch_norm(sd, pre + 'global_avg_pool.2', g, cfg)))
    outs.append(F.interpolate(g, size=x.shape[2:], mode='bilinear', align_corners=True))
    h = _qz(F.conv2d(_tr('aspp_cat', torch.cat(outs, 1)), _q(sd[pre + 'conv1.weight'])), cfg)
    h = F.relu(batch_norm(sd, pre + 'bn1', h, cfg))
    return _tr('aspp_out', _q(_dropout(h, 0.5, cfg, drop)))


def _dropout(x, p, cfg, drop):
    if drop is False or not cfg.training:
        return x
    return F.dropout(x, p, True)


def decoder_forward(sd, x, low, cfg, pre='', drop=None):
    """modeling/decoder.py:34-43."""
    l = _qz(F.conv2d(low, _q(sd[pre + 'conv1.weight'])), cfg)
    l = _q(F.relu(batch_norm(sd, pre + 'bn1', l, cfg)))
    x = _q(F.interpolate(x, size=l.shape[2:], mode='bilinear', align_corners=True))
    h = _tr('dec_cat', torch.cat((x, l), 1))
    h = _qz(F.conv2d(h, _q(sd[pre + 'last_conv.0.weight']), None, 1, 1), cfg)
    h = _tr('dec_y1', _q(_dropout(F.relu(batch_norm(sd, pre + 'last_conv.1', h, cfg)), 0.5, cfg, drop)))
    h = _qz(F.conv2d(h, _q(sd[pre + 'last_conv.4.weight']), None, 1, 1), cfg)
    h = _tr('dec_y2', _q(_dropout(F.relu(batch_norm(sd, pre + 'last_conv.5', h, cfg)), 0.1, cfg, drop)))
    return _tr('dec_logits', _q(F.conv2d(h, _q(sd[pre + 'last_conv.8.weight']), sd[pre + 'last_conv.8.bias'])))


def deeplab_forward(sd, x, cfg, output_stride=16, drop=None):
    """modeling/deeplab.py:27-33."""
    high, low = mobilenet_forward(sd, x, cfg, output_stride, 'backbone.features')
    h = aspp_forward(sd, high, cfg, output_stride, 'aspp.', drop)
    h = decoder_forward(sd, h, low, cfg, 'decoder.', drop)
    return F.interpolate(h, size=x.shape[2:], mode='bilinear', align_corners=True)


def discriminator_forward(sd, x):
    """modeling/discriminator.py:22-35."""
    x = _q(x)
    for name in ('conv1', 'conv2', 'conv3', 'conv4'):
        x = _q(F.leaky_relu(F.conv2d(x, _q(sd[name + '.weight']), sd[name + '.bias'], 2, 1), 0.2))
    return _q(F.conv2d(x, _q(sd['classifier.weight']), sd['classifier.bias'], 2, 1))


def domain_classifier_forward(sd, x, cfg, drop=None):
    """modeling/domian.py:27-32."""
    h = _qz(F.conv2d(_q(x), _q(sd['DC_adnn1.0.weight'])), cfg)
    h = _q(_dropout(F.relu(batch_norm(sd, 'DC_adnn1.1', h, cfg)), 0.5, cfg, drop))
    h = _qz(F.conv2d(h, _q(sd['DC_adnn2.0.weight']), None, 1, 1), cfg)
    h = _q(_dropout(F.relu(batch_norm(sd, 'DC_adnn2.1', h, cfg)), 0.5, cfg, drop))
    return _q(F.conv2d(h, _q(sd['DC_adnn3.weight']), sd['DC_adnn3.bias'], 1, 1))


def seg_cross_entropy(logit, target, weight=None, ignore_index=255):
    """utils/loss.py:21-30."""
    return F.cross_entropy(logit, target.long(), weight=weight, ignore_index=ignore_index, reduction='mean')


def focal_loss(logit, target, weight=None, ignore_index=255, gamma=2, alpha=0.5):
    """utils/loss.py:32-46."""
    logpt = -seg_cross_entropy(logit, target, weight, ignore_index)
    pt = torch.exp(logpt)
    if alpha is not None:
        logpt = logpt * alpha
    return -((1 - pt) ** gamma) * logpt


def domain_loss(src_logit, tgt_logit):
    """utils/loss.py:57-69 -> (loss tensor, accuracy float)."""
    assert src_logit.size() == tgt_logit.size()
    n, _, h, w = src_logit.shape
    zeros = torch.zeros((n, h, w), dtype=torch.long)
    loss = F.cross_entropy(src_logit, zeros) + F.cross_entropy(tgt_logit, zeros + 1)
    acc = (torch.sum(1 - torch.argmax(src_logit, 1)) + torch.sum(torch.argmax(tgt_logit, 1))).float() / 2 / n / h / w
    return loss, acc.item()


def confusion_matrix(gt, pred, num_class):
    """utils/metrics.py:34-39 (numpy)."""
    mask = (gt >= 0) & (gt < num_class)
    label = num_class * gt[mask].astype('int') + pred[mask]
    return np.bincount(label, minlength=num_class ** 2).reshape(num_class, num_class)


def evaluator_metrics(cm):
    """utils/metrics.py:9-32 -> dict(PA, mPA, mIoU, IoU, fwIoU) from a float64 confusion matrix."""
    cm = cm.astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        pa = np.diag(cm).sum() / cm.sum()
        mpa = np.nanmean(np.diag(cm) / cm.sum(axis=1))
        iou = np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))
        freq = cm.sum(axis=1) / cm.sum()
        fw = (freq[freq > 0] * iou[freq > 0]).sum()
    return dict(PA=pa, mPA=mpa, mIoU=np.nanmean(iou), IoU=iou, fwIoU=fw)


def poly_lr(base_lr, T, N):
    """utils/lr_scheduler.py:47-48."""
    return base_lr * pow((1 - 1.0 * T / N), 0.9)


def split_lr_groups(sd_names):
    """modeling/deeplab.py:42-72: backbone parameters (1x lr) vs ASPP+decoder (10x lr)."""
    one = [k for k in sd_names if k.startswith('backbone.')]
    ten = [k for k in sd_names if k.startswith('aspp.') or k.startswith('decoder.')]
    return one, ten


PARAM_SUFFIXES = ('.weight', '.bias')


def leaf_params(sd):
    """Parameter entries of a state dict (everything but BN running statistics / counters).  The
    backbone registers features[0:4] / features[4:] a second time as low_level_features /
    high_level_features (mobilenet.py:116-117); those aliases are not separate parameters."""
    return {k: v for k, v in sd.items() if k.endswith(PARAM_SUFFIXES) and v.dtype.is_floating_point
            and '.low_level_features.' not in k and '.high_level_features.' not in k
            and not k.startswith(('low_level_features.', 'high_level_features.'))}


def _conv_w(sd, name, cout, cin, k, gen, bias=False, default_init=False):
    """Weights of one nn.Conv2d: kaiming_normal_ where the reference's _init_weight / _initialize_weights loops touch
    the layer (mobilenet.py:134-145, assp.py:22-32,80-91, decoder.py:45-54), nn.Conv2d's own default
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias) otherwise (discriminator.py has no init loop; biases are
    never re-initialised)."""
    fan_in = cin * k * k
    bound = 1.0 / math.sqrt(fan_in)
    if default_init:
        sd[name + '.weight'] = (torch.rand(cout, cin, k, k, generator=gen) * 2 - 1) * bound
    else:
        sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=gen) * math.sqrt(2.0 / fan_in)
    if bias:
        sd[name + '.bias'] = (torch.rand(cout, generator=gen) * 2 - 1) * bound


def _bn(sd, name, c):
    sd[name + '.weight'], sd[name + '.bias'] = torch.ones(c), torch.zeros(c)
    sd[name + '.running_mean'], sd[name + '.running_var'] = torch.zeros(c), torch.ones(c)
    sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)


def init_deeplab(seed=1, output_stride=16, num_classes=19):
    """A randomly initialised state dict of DeepLab(backbone='mobilenet') with the reference's key names, shapes and
    initial distributions (deeplab.py:9-25 -> mobilenet.py:71-117, assp.py:34-63, decoder.py:7-32) built from plain
    tensors -- so that the CPU baseline needs nothing but this module and torch.  Same distributions as the
    reference's constructors, not the same random stream (tests that need identical weights on both sides copy them
    from one model)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    f = 'backbone.features.'
    _conv_w(sd, f + '0.0', 32, 3, 3, gen)
    _bn(sd, f + '0.1', 32)
    for k, (inp, oup, stride, dil, t) in enumerate(mnv2_plan(output_stride), start=1):
        hidden, i = inp * t, 0
        if t != 1:
            _conv_w(sd, '%s%d.conv.0' % (f, k), hidden, inp, 1, gen)
            _bn(sd, '%s%d.conv.1' % (f, k), hidden)
            i = 3
        sd['%s%d.conv.%d.weight' % (f, k, i)] = torch.randn(hidden, 1, 3, 3, generator=gen) * math.sqrt(2.0 / 9)
        _bn(sd, '%s%d.conv.%d' % (f, k, i + 1), hidden)
        _conv_w(sd, '%s%d.conv.%d' % (f, k, i + 3), oup, hidden, 1, gen)
        _bn(sd, '%s%d.conv.%d' % (f, k, i + 4), oup)
    for k in (1, 2, 3, 4):
        _conv_w(sd, 'aspp.aspp%d.atrous_conv' % k, 256, 320, 1 if k == 1 else 3, gen)
        _bn(sd, 'aspp.aspp%d.bn' % k, 256)
    _conv_w(sd, 'aspp.global_avg_pool.1', 256, 320, 1, gen)
    _bn(sd, 'aspp.global_avg_pool.2', 256)
    _conv_w(sd, 'aspp.conv1', 256, 1280, 1, gen)
    _bn(sd, 'aspp.bn1', 256)
    _conv_w(sd, 'decoder.conv1', 48, 24, 1, gen)
    _bn(sd, 'decoder.bn1', 48)
    _conv_w(sd, 'decoder.last_conv.0', 256, 304, 3, gen)
    _bn(sd, 'decoder.last_conv.1', 256)
    _conv_w(sd, 'decoder.last_conv.4', 256, 256, 3, gen)
    _bn(sd, 'decoder.last_conv.5', 256)
    _conv_w(sd, 'decoder.last_conv.8', num_classes, 256, 1, gen, bias=True)
    return sd


def init_discriminator(seed=2, num_classes=19, ndf=64):
    """State dict of FCDiscriminator (discriminator.py:8-20) with nn.Conv2d's default initialisation."""
    gen = torch.Generator().manual_seed(seed)
    sd, cin = {}, num_classes
    for name, cout in (('conv1', ndf), ('conv2', ndf * 2), ('conv3', ndf * 4), ('conv4', ndf * 8), ('classifier', 1)):
        _conv_w(sd, name, cout, cin, 4, gen, bias=True, default_init=True)
        cin = cout
    return sd


def adapt_step(g_sd, d_sd, opt_g, opt_d, src_image, src_label, tgt_image, cfg, drop=None, output_stride=16):
    """One iteration of Trainer.training in train_adapt.py:137-181 (device-agnostic): returns
    (loss_seg, loss_adv, loss_D_src, loss_D_tgt) as python floats.  g_sd/d_sd hold leaf tensors
    with requires_grad=True for parameters; opt_g/opt_d are torch optimizers over them."""
    opt_g.zero_grad()
    opt_d.zero_grad()
    d_params = [v for v in leaf_params(d_sd).values()]
    for p in d_params:
        p.requires_grad_(False)
    src_out = deeplab_forward(g_sd, src_image, cfg, output_stride, drop)
    loss_seg = seg_cross_entropy(src_out, src_label)
    loss_seg.backward()
    tgt_out = deeplab_forward(g_sd, tgt_image, cfg, output_stride, drop)
    d_out = discriminator_forward(d_sd, F.softmax(tgt_out, dim=0))
    loss_adv = F.binary_cross_entropy_with_logits(d_out, torch.zeros_like(d_out))
    loss_adv.backward()
    for p in d_params:
        p.requires_grad_(True)
    d_out = discriminator_forward(d_sd, F.softmax(src_out.detach(), dim=0))
    loss_d_src = F.binary_cross_entropy_with_logits(d_out, torch.zeros_like(d_out))
    loss_d_src.backward()
    d_out = discriminator_forward(d_sd, F.softmax(tgt_out.detach(), dim=0))
    loss_d_tgt = F.binary_cross_entropy_with_logits(d_out, torch.ones_like(d_out))
    loss_d_tgt.backward()
    opt_g.step()
    opt_d.step()
    return loss_seg.item(), loss_adv.item(), loss_d_src.item(), loss_d_tgt.item()


def feature_step(f_sd, a_sd, y_sd, dc_sd, opts, src_image, src_label, tgt_image, cfg, drop=None, output_stride=16):
    """One iteration of Trainer.training in train.py:173-204: the summed loss reaches every
    parameter; task, d and d_inv optimizers step (c_optimizer never does).  opts = (task, d, d_inv).
    tgt_image=None is the single-domain branch (args.dataset == 'gtav', train.py:205-210): the domain classifier
    still runs on the source features (:187, its BatchNorm running statistics move) but only the task loss is
    back-propagated and only task_optimizer steps."""
    for o in opts:
        o.zero_grad()

    def fwd(img):
        high0, low = mobilenet_forward(f_sd, img, cfg, output_stride)
        high = aspp_forward(a_sd, high0, cfg, output_stride, '', drop)
        out = F.interpolate(decoder_forward(y_sd, high, low, cfg, '', drop), img.shape[2:], mode='bilinear',
                            align_corners=True)
        return out, domain_classifier_forward(dc_sd, high, cfg, drop)

    src_out, src_d = fwd(src_image)
    task = seg_cross_entropy(src_out, src_label)
    if tgt_image is None:
        task.backward()
        opts[0].step()
        return task.item(), 0.0, 0.0, 0
    _, tgt_d = fwd(tgt_image)
    d_loss, d_acc = domain_loss(src_d, tgt_d)
    d_inv_loss, _ = domain_loss(tgt_d, src_d)
    (task + d_loss + d_inv_loss).backward()
    for o in opts:
        o.step()
    return task.item(), d_loss.item(), d_inv_loss.item(), d_acc
