"""Generates tests/golden/*.npz from the REAL reference (/root/reference, build container only)
and checks oracle/ref_port.py against it.  Run:  python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs on fixed seeds are committed here as
small fixtures; tests/test_oracle.py re-checks the oracle against them everywhere.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("S2R_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from modeling.backbone import mobilenet as ref_mobilenet  # noqa: E402
ref_mobilenet.MobileNetV2._load_pretrained_model = lambda self: None  # the checkpoint blob is not in the tree
from modeling.deeplab import DeepLab as RefDeepLab  # noqa: E402
from modeling.discriminator import FCDiscriminator as RefD  # noqa: E402
from modeling.domian import DomainClassifer as RefDC  # noqa: E402
from modeling.assp import ASPP as RefASPP  # noqa: E402
from modeling.decoder import Decoder as RefDecoder  # noqa: E402
from utils.loss import SegmentationLosses as RefSegLoss, DomainLosses as RefDomLoss  # noqa: E402
from utils.metrics import Evaluator as RefEvaluator  # noqa: E402

from oracle import ref_port as O  # noqa: E402

torch.set_num_threads(8)


def no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def clone_sd(model, grad=True):
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if grad:
        for k, v in O.leaf_params(sd).items():
            v.requires_grad_(True)
    return sd


def checksums(model):
    names, vals = [], []
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            names.append(k)
            vals.append([float(v.double().sum()), float(v.double().abs().sum())])
    return np.array(names), np.array(vals, dtype=np.float64)


def head(t, n=4096):
    """First n elements of a tensor (fixtures stay small; norms of the full tensors are stored too)."""
    return t.detach().reshape(-1)[:n].numpy().copy()


def relerr(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_inputs(seed, n, h, w, ncls=19):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g)
    lab = torch.randint(0, ncls + 1, (n, h, w), generator=g).float()
    lab[lab == ncls] = 255
    return x, lab


def deeplab_case(tag, n, h, w, train):
    torch.manual_seed(1)
    ref = no_dropout(RefDeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False))
    ref.train(train)
    names, sums = checksums(ref)
    sd = clone_sd(ref)
    x, lab = make_inputs(0, n, h, w)
    crit = RefSegLoss().build_loss('ce')
    out = ref(x)
    loss = crit(out, lab)
    fix = dict(x=x.numpy(), label=lab.numpy(), logits=out.detach().numpy(), loss=np.float64(loss.item()),
               param_names=names, param_sums=sums)
    cfg = O.BNCfg(training=train)
    o_out = O.deeplab_forward(sd, x, cfg, 16, drop=False)
    o_loss = O.seg_cross_entropy(o_out, lab)
    print(tag, 'logits rel err oracle vs reference', relerr(o_out.detach(), out.detach()), 'loss', loss.item(), o_loss.item())
    assert relerr(o_out.detach(), out.detach()) < 1e-5
    if train:
        loss.backward()
        o_loss.backward()
        picks = ['backbone.features.0.0.weight', 'backbone.features.1.conv.0.weight', 'backbone.features.2.conv.0.weight',
                 'backbone.features.2.conv.1.weight', 'backbone.features.2.conv.1.bias',
                 'backbone.features.2.conv.3.weight', 'backbone.features.7.conv.3.weight',
                 'backbone.features.17.conv.6.weight', 'aspp.aspp1.atrous_conv.weight', 'aspp.aspp3.atrous_conv.weight',
                 'aspp.global_avg_pool.1.weight', 'aspp.conv1.weight', 'aspp.bn1.weight', 'decoder.conv1.weight',
                 'decoder.last_conv.0.weight', 'decoder.last_conv.4.weight', 'decoder.last_conv.8.weight',
                 'decoder.last_conv.8.bias']
        refp = dict(ref.named_parameters())
        worst = 0.0
        norms = {}
        for k, p in refp.items():
            e = relerr(sd[k].grad, p.grad)
            worst = max(worst, e)
            norms[k] = float(p.grad.double().norm())
        print(tag, 'worst param-grad rel err oracle vs reference', worst)
        assert worst < 2e-3, worst
        for k in picks:
            fix['grad:' + k] = head(refp[k].grad)
        fix['grad_norm_names'] = np.array(list(norms.keys()))
        fix['grad_norms'] = np.array(list(norms.values()), dtype=np.float64)
        # running statistics after one training forward
        for k in ['backbone.features.0.1.running_mean', 'backbone.features.0.1.running_var',
                  'backbone.features.2.conv.1.running_mean', 'backbone.features.2.conv.1.running_var',
                  'aspp.global_avg_pool.2.running_var', 'decoder.last_conv.5.running_var']:
            fix['buf:' + k] = ref.state_dict()[k].numpy()
            assert relerr(sd[k], ref.state_dict()[k]) < 1e-5, k
    np.savez_compressed(os.path.join(HERE, tag + '.npz'), **fix)


def discriminator_case():
    torch.manual_seed(2)
    ref = RefD(num_classes=19)
    names, sums = checksums(ref)
    sd = clone_sd(ref)
    g = torch.Generator().manual_seed(3)
    x = torch.softmax(torch.randn(2, 19, 64, 96, generator=g), dim=0).requires_grad_(True)
    out = ref(x)
    loss = torch.nn.BCEWithLogitsLoss()(out, torch.zeros_like(out))
    loss.backward()
    xo = x.detach().clone().requires_grad_(True)
    o_out = O.discriminator_forward(sd, xo)
    o_loss = torch.nn.functional.binary_cross_entropy_with_logits(o_out, torch.zeros_like(o_out))
    o_loss.backward()
    assert relerr(o_out.detach(), out.detach()) < 1e-6
    assert relerr(xo.grad, x.grad) < 1e-5
    fix = dict(x=x.detach().numpy(), out=out.detach().numpy(), loss=np.float64(loss.item()), dx=x.grad.numpy(),
               param_names=names, param_sums=sums)
    for k, p in ref.named_parameters():
        assert relerr(sd[k].grad, p.grad) < 1e-4, k
        if k in ('conv1.weight', 'conv1.bias', 'conv3.weight', 'classifier.weight', 'classifier.bias'):
            fix['grad:' + k] = head(p.grad)
            fix['gradnorm:' + k] = np.float64(p.grad.double().norm())
    np.savez_compressed(os.path.join(HERE, 'discriminator.npz'), **fix)
    print('discriminator ok')


def domain_case():
    torch.manual_seed(4)
    ref = no_dropout(RefDC('mobilenet', torch.nn.BatchNorm2d))
    ref.train()
    names, sums = checksums(ref)
    sd = clone_sd(ref)
    g = torch.Generator().manual_seed(5)
    xs = torch.randn(2, 256, 9, 12, generator=g)
    xt = torch.randn(2, 256, 9, 12, generator=g)
    crit = RefDomLoss().build_loss()
    cfg = O.BNCfg(True)
    ps, pt = ref(xs), ref(xt)
    loss, acc = crit(ps, pt)
    loss.backward()
    os_, ot_ = O.domain_classifier_forward(sd, xs, cfg, False), O.domain_classifier_forward(sd, xt, cfg, False)
    ol, oacc = O.domain_loss(os_, ot_)
    ol.backward()
    assert relerr(os_.detach(), ps.detach()) < 1e-5 and abs(oacc - acc) < 1e-7
    refp = dict(ref.named_parameters())
    for k, p in refp.items():
        assert relerr(sd[k].grad, p.grad) < 1e-3, (k, relerr(sd[k].grad, p.grad))
    fix = dict(xs=xs.numpy(), xt=xt.numpy(), ps=ps.detach().numpy(), pt=pt.detach().numpy(),
               loss=np.float64(loss.item()), acc=np.float64(acc), param_names=names, param_sums=sums)
    for k in ('DC_adnn3.weight', 'DC_adnn1.0.weight', 'DC_adnn2.0.weight', 'DC_adnn2.1.weight'):
        fix['grad:' + k] = head(refp[k].grad)
        fix['gradnorm:' + k] = np.float64(refp[k].grad.double().norm())
    np.savez_compressed(os.path.join(HERE, 'domain_classifier.npz'), **fix)
    # the analytic known answer of utils/loss.py:80-87
    a, b = torch.ones(1, 1, 7, 7), torch.zeros(1, 1, 7, 7)
    l, ac = crit(torch.cat([a, b], 1), torch.cat([b, a], 1))
    assert abs(l.item() - 0.626523) < 1e-5 and ac == 1.0
    print('domain classifier ok', l.item(), ac)


def evaluator_case():
    rng = np.random.RandomState(7)
    gt = rng.randint(0, 20, size=(2, 37, 53)).astype(np.float32)
    gt[gt == 19] = 255
    pred = rng.randint(0, 19, size=(2, 37, 53)).astype(np.int64)
    ev = RefEvaluator(19)
    ev.add_batch(gt, pred)
    ev.add_batch(gt[:, ::-1].copy(), pred)
    cm = ev.confusion_matrix
    miou, iou = ev.Mean_Intersection_over_Union()
    fix = dict(gt=gt, pred=pred, cm=cm, PA=ev.Pixel_Accuracy(), mPA=ev.Pixel_Accuracy_Class(), mIoU=miou, IoU=iou,
               fwIoU=ev.Frequency_Weighted_Intersection_over_Union())
    ocm = O.confusion_matrix(gt, pred, 19) + O.confusion_matrix(gt[:, ::-1].copy(), pred, 19)
    assert (ocm == cm).all()
    m = O.evaluator_metrics(ocm)
    assert m['mIoU'] == miou and m['fwIoU'] == fix['fwIoU']
    np.savez_compressed(os.path.join(HERE, 'evaluator.npz'), **fix)
    print('evaluator ok', miou)


def adapt_step_case():
    """Two iterations of train_adapt.py:137-181 on the reference modules (device-agnostic restatement
    of the script body, which hard-codes .cuda()) vs the oracle's adapt_step."""
    import torch.nn.functional as F
    torch.manual_seed(1)
    G = no_dropout(RefDeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False))
    D = RefD(num_classes=19)
    G.train()
    D.train()
    g_sd, d_sd = clone_sd(G), clone_sd(D)
    lr = 5e-4
    opt = torch.optim.SGD([{'params': G.get_1x_lr_params(), 'lr': lr}, {'params': G.get_10x_lr_params(), 'lr': lr * 10}],
                          momentum=0.9, weight_decay=5e-4, nesterov=False)
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.9, 0.99))
    one, ten = O.split_lr_groups(list(O.leaf_params(g_sd).keys()))
    o_opt = torch.optim.SGD([{'params': [g_sd[k] for k in one], 'lr': lr}, {'params': [g_sd[k] for k in ten], 'lr': lr * 10}],
                            momentum=0.9, weight_decay=5e-4, nesterov=False)
    o_opt_d = torch.optim.Adam(list(O.leaf_params(d_sd).values()), lr=1e-4, betas=(0.9, 0.99))
    crit = RefSegLoss().build_loss('ce')
    bce = torch.nn.BCEWithLogitsLoss()
    cfg = O.BNCfg(True)
    hist = []
    for it in range(2):
        src, lab = make_inputs(100 + it, 2, 65, 97)
        tgt, _ = make_inputs(200 + it, 2, 65, 97)
        for o in (opt, opt_d, o_opt, o_opt_d):
            for gi, grp in enumerate(o.param_groups):
                grp['lr'] = O.poly_lr(lr, it, 10) * (10 if gi > 0 else 1)   # lr_scheduler.py:63-70
        opt.zero_grad()
        opt_d.zero_grad()
        for p in D.parameters():
            p.requires_grad = False
        so = G(src)
        ls = crit(so, lab)
        ls.backward()
        to = G(tgt)
        la = bce(D(F.softmax(to, dim=0)), torch.zeros(2, 1, 2, 3))
        la.backward()
        for p in D.parameters():
            p.requires_grad = True
        l1 = bce(D(F.softmax(so.detach(), dim=0)), torch.zeros(2, 1, 2, 3))
        l1.backward()
        l2 = bce(D(F.softmax(to.detach(), dim=0)), torch.ones(2, 1, 2, 3))
        l2.backward()
        opt.step()
        opt_d.step()
        ref_losses = (ls.item(), la.item(), l1.item(), l2.item())
        o_losses = O.adapt_step(g_sd, d_sd, o_opt, o_opt_d, src, lab, tgt, cfg, drop=False)
        print('adapt it', it, ref_losses, o_losses)
        assert np.allclose(ref_losses, o_losses, rtol=2e-4, atol=1e-6)
        hist.append(ref_losses)
    fix = dict(losses=np.array(hist, dtype=np.float64))
    for k in ['backbone.features.0.0.weight', 'decoder.last_conv.8.weight', 'aspp.conv1.weight']:
        w = dict(G.named_parameters())[k].detach()
        assert relerr(g_sd[k].detach(), w) < 1e-4, k
        fix['w:' + k] = head(w)
        fix['wnorm:' + k] = np.float64(w.double().norm())
    fix['wd:conv1.weight'] = head(D.conv1.weight)
    fix['wdnorm:conv1.weight'] = np.float64(D.conv1.weight.detach().double().norm())
    assert relerr(d_sd['conv1.weight'].detach(), D.conv1.weight.detach()) < 1e-4
    np.savez_compressed(os.path.join(HERE, 'adapt_step.npz'), **fix)


def feature_step_case():
    """Two iterations of train.py:173-204 (BASELINE config 4: FCN-in-the-wild feature adaptation) on the reference's
    own modules -- MobileNetV2, ASPP, Decoder, DomainClassifer built as train.py:46-56 builds them, the three stepping
    optimizers of :63-82 (both flavours: Adam, the script default, and SGD), DomainLosses -- vs the oracle's
    feature_step.  The script body is restated device-agnostically (it calls .cuda() on its inputs)."""
    import torch.nn.functional as F
    nn = torch.nn
    fix = {}
    for flavour in ('Adam', 'SGD'):
        torch.manual_seed(7)
        bb = no_dropout(ref_mobilenet.MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d))
        aspp = no_dropout(RefASPP(backbone='mobilenet', output_stride=16, BatchNorm=nn.BatchNorm2d))
        dec = no_dropout(RefDecoder(num_classes=19, backbone='mobilenet', BatchNorm=nn.BatchNorm2d))
        dc = no_dropout(RefDC(backbone='mobilenet', BatchNorm=nn.BatchNorm2d))
        mods = (bb, aspp, dec, dc)
        for m in mods:
            m.train()
        sds = [clone_sd(m) for m in mods]
        lr = 5e-4
        if flavour == 'Adam':
            mk = lambda ps: torch.optim.Adam(ps, lr=lr)  # noqa: E731
        else:
            mk = lambda ps: torch.optim.SGD(ps, lr=lr, momentum=0.9, weight_decay=5e-4, nesterov=False)  # noqa: E731
        f_params = list(bb.parameters()) + list(aspp.parameters())
        opts = (mk(f_params + list(dec.parameters())), mk(list(dc.parameters())), mk(f_params))
        c_opt = mk(f_params + list(dec.parameters()))              # train.py:73-75: scheduled and zeroed, never stepped
        o_fp = list(O.leaf_params(sds[0]).values()) + list(O.leaf_params(sds[1]).values())
        o_opts = (mk(o_fp + list(O.leaf_params(sds[2]).values())), mk(list(O.leaf_params(sds[3]).values())), mk(o_fp))
        task_loss_fn = RefSegLoss().build_loss('ce')
        domain_loss_fn = RefDomLoss().build_loss()
        g = torch.Generator().manual_seed(11)
        hist = []
        for it in range(2):
            src = torch.randn(2, 3, 64, 96, generator=g)
            tgt = torch.randn(2, 3, 64, 96, generator=g)
            lab = torch.randint(0, 19, (2, 64, 96), generator=g).float()
            for o in opts + (c_opt,) + o_opts:
                o.param_groups[0]['lr'] = O.poly_lr(lr, it, 10)          # lr_scheduler.py:63-70 (one group each)
            for o in opts + (c_opt,):
                o.zero_grad()
            sh0, sl = bb(src)
            sh = aspp(sh0)
            so = F.interpolate(dec(sh, sl), src.size()[2:], mode='bilinear', align_corners=True)
            sd_pred = dc(sh)
            task = task_loss_fn(so, lab)
            th0, tl = bb(tgt)
            th = aspp(th0)
            F.interpolate(dec(th, tl), tgt.size()[2:], mode='bilinear', align_corners=True)   # tgt_output: computed, unused
            td_pred = dc(th)
            d_loss, d_acc = domain_loss_fn(sd_pred, td_pred)
            d_inv_loss, _ = domain_loss_fn(td_pred, sd_pred)
            (task + d_loss + d_inv_loss).backward()
            for o in opts:
                o.step()
            ref = (task.item(), d_loss.item(), d_inv_loss.item(), float(d_acc))
            got = O.feature_step(sds[0], sds[1], sds[2], sds[3], o_opts, src, lab, tgt, O.BNCfg(True), drop=False)
            print('feature', flavour, 'it', it, ref, got)
            assert np.allclose(ref, got, rtol=2e-4, atol=1e-6), (ref, got)
            hist.append(ref)
        fix['losses_' + flavour] = np.array(hist, dtype=np.float64)
        for sd, m, k in ((sds[0], bb, 'features.0.0.weight'), (sds[1], aspp, 'conv1.weight'),
                         (sds[2], dec, 'last_conv.8.weight'), (sds[3], dc, 'DC_adnn3.weight')):
            w = dict(m.named_parameters())[k].detach()
            assert relerr(sd[k].detach(), w) < 2e-4, (flavour, k, relerr(sd[k].detach(), w))
            fix['w_%s:%s' % (flavour, k)] = head(w)
            fix['wnorm_%s:%s' % (flavour, k)] = np.float64(w.double().norm())
        # running statistics saw both domains, source first
        rm = dict(bb.named_buffers())['features.0.1.running_mean']
        assert relerr(sds[0]['features.0.1.running_mean'], rm) < 1e-5
        fix['rm_%s:features.0.1.running_mean' % flavour] = head(rm)
    np.savez_compressed(os.path.join(HERE, 'feature_step.npz'), **fix)


def sync_bn_case():
    """The reference's synchronised BatchNorm protocol (row a11: batchnorm.py:48-125, comm.py:18-129, replicate.py:27-44)
    EXECUTED on the CPU: two replicas of one SynchronizedBatchNorm2d registered through the reference's own
    execute_replication_callbacks, their forwards run in two threads so that the replica statistics really travel
    through SlavePipe / SyncMaster.run_master / _data_parallel_master / _compute_mean_std.  Only the two torch-internal
    CUDA collectives that function calls (torch.nn.parallel._functions.ReduceAddCoalesced / Broadcast) are replaced
    by their CPU meaning (sum of the replica tensors / the same tensors for every replica).  Compared with the
    oracle's sync_clamp branch on the concatenated batch: outputs, running statistics, all gradients.  (Regenerating
    the fixture can change `dx` in the last bit: the replicas' contributions to the shared statistics' gradients are
    accumulated in thread-arrival order.  The tests compare with 1e-5.)"""
    import copy
    import threading
    from modeling.sync_batchnorm import batchnorm as ref_bn
    from modeling.sync_batchnorm import replicate as ref_rep

    class _Reduce:
        @staticmethod
        def apply(dev, n, *ts):
            return tuple(sum(ts[i::n][1:], ts[i::n][0]) for i in range(n))

    class _Bcast:
        @staticmethod
        def apply(devs, *ts):
            return tuple(ts) * len(devs)

    ref_bn.ReduceAddCoalesced, ref_bn.Broadcast = _Reduce, _Bcast
    Cc = 6
    g = torch.Generator().manual_seed(21)
    x_all = torch.randn(5, Cc, 7, 9, generator=g) * 2 + 0.5
    x_all[:, 2] = 0.75                                     # a constant channel: variance 0 -> clamp(eps), not var + eps
    dy_all = torch.randn(5, Cc, 7, 9, generator=g)
    shards = (slice(0, 2), slice(2, 5))                    # unequal replica batches
    bn = ref_bn.SynchronizedBatchNorm2d(Cc)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(Cc, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(Cc, generator=g))
    bn.train()
    sd = {'bn.weight': bn.weight.detach().clone().requires_grad_(True), 'bn.bias': bn.bias.detach().clone().requires_grad_(True),
          'bn.running_mean': bn.running_mean.clone(), 'bn.running_var': bn.running_var.clone()}
    replicas = []
    for _ in shards:
        r = copy.copy(bn)
        r._parameters, r._buffers = bn._parameters.copy(), bn._buffers.copy()
        replicas.append(r)
    ref_rep.execute_replication_callbacks(replicas)
    xs = [x_all[s].clone().requires_grad_(True) for s in shards]
    outs = [None, None]

    def work(i):
        outs[i] = replicas[i](xs[i])

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert all(o is not None for o in outs)
    sum((o * dy_all[s]).sum() for o, s in zip(outs, shards)).backward()
    y_ref = torch.cat([o.detach() for o in outs])
    xr = x_all.clone().requires_grad_(True)
    y = O.batch_norm(sd, 'bn', xr, O.BNCfg(True, 0.1, 1e-5, sync_clamp=True))
    (y * dy_all).sum().backward()
    # the updated running statistics land on the ORIGINAL module: the replicas share its SyncMaster, whose callback is
    # the original's bound _data_parallel_master (batchnorm.py:44, :122-123 assign them on that `self`)
    rm, rv = bn.running_mean, bn.running_var
    assert relerr(y.detach(), y_ref) < 1e-6
    assert relerr(sd['bn.running_mean'], rm) < 1e-6 and relerr(sd['bn.running_var'], rv) < 1e-6
    dx_ref = torch.cat([x.grad for x in xs])
    assert relerr(xr.grad, dx_ref) < 1e-5, relerr(xr.grad, dx_ref)
    assert relerr(sd['bn.weight'].grad, bn.weight.grad) < 1e-5 and relerr(sd['bn.bias'].grad, bn.bias.grad) < 1e-5
    # F.batch_norm (the non-parallel branch) differs on the constant channel: 1/sqrt(var + eps) vs clamp(var, eps)^-1/2
    print('sync-bn: y', relerr(y.detach(), y_ref), 'dx', relerr(xr.grad, dx_ref), 'running_var', rv.tolist())
    np.savez_compressed(os.path.join(HERE, 'sync_bn.npz'), x=x_all.numpy(), dy=dy_all.numpy(),
                        weight=bn.weight.detach().numpy(), bias=bn.bias.detach().numpy(), y=y_ref.numpy(),
                        dx=dx_ref.numpy(), dweight=bn.weight.grad.numpy(), dbias=bn.bias.grad.numpy(),
                        running_mean=rm.numpy(), running_var=rv.numpy())


def reference_method(script, name, ns):
    """A method of one of the reference's top-level scripts, compiled UNMODIFIED from its source segment (the scripts
    themselves cannot be imported: tensorboardX is missing and they parse argv / build CUDA loaders at import)."""
    import ast
    tree = ast.parse(open(os.path.join(REF, script)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            exec(compile(ast.Module(body=[node], type_ignores=[]), script, 'exec'), ns)
            return ns[name]
    raise RuntimeError('%s not found in %s' % (name, script))


def checkpoint_layout(ck):
    """JSON description of a checkpoint dict the reference wrote: every tensor replaced by [dtype, shape], everything
    else (keys, order, parameter-group index lists, hyper-parameters, scalars) kept."""
    import json

    def conv(v):
        if torch.is_tensor(v):
            return {'__tensor__': [str(v.dtype).replace('torch.', ''), list(v.shape)]}
        if isinstance(v, dict):
            return {'__items__': [[k if isinstance(k, str) else {'__int__': int(k)}, conv(x)] for k, x in v.items()]}
        if isinstance(v, (list, tuple)):
            return [conv(x) for x in v]
        if isinstance(v, (bool, int, float, str)) or v is None:
            return v
        return float(v)
    return json.dumps(conv(ck))


def adapt_loop_case():
    """BASELINE configs 2/3: the reference's own `Trainer.training` (train_adapt.py:115-196, unmodified source
    segment) run for one epoch of ten iterations on the CPU -- reference DeepLab / FCDiscriminator / criterion /
    LR_Scheduler / torch optimizers; the only concession is `Tensor.cuda()` made the identity while it runs (the body
    calls it unconditionally on its BCE targets).  Per-iteration losses are recorded by wrapping the two loss
    callables.  Compared with ten oracle adapt_step iterations under the oracle's poly schedule."""
    import types
    import torch.nn.functional as F
    from torch.autograd import Variable
    from utils.lr_scheduler import LR_Scheduler as RefSched

    class Bar(list):
        def set_description(self, text):
            pass

    training = reference_method('train_adapt.py', 'training', {'np': np, 'torch': torch, 'F': F, 'Variable': Variable,
                                                               'tqdm': lambda it: Bar(it)})
    torch.manual_seed(1)
    G = no_dropout(RefDeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False))
    D = RefD(num_classes=19)
    D.train()
    g_sd, d_sd = clone_sd(G), clone_sd(D)
    lr, n_it = 5e-4, 10
    opt = torch.optim.SGD([{'params': G.get_1x_lr_params(), 'lr': lr}, {'params': G.get_10x_lr_params(), 'lr': lr * 10}],
                          momentum=0.9, weight_decay=5e-4, nesterov=False)                      # train_adapt.py:54-58
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.9, 0.99))                         # :59-60
    one, ten = O.split_lr_groups(list(O.leaf_params(g_sd).keys()))
    o_opt = torch.optim.SGD([{'params': [g_sd[k] for k in one], 'lr': lr}, {'params': [g_sd[k] for k in ten], 'lr': lr * 10}],
                            momentum=0.9, weight_decay=5e-4, nesterov=False)
    o_opt_d = torch.optim.Adam(list(O.leaf_params(d_sd).values()), lr=1e-4, betas=(0.9, 0.99))
    loader = []
    for it in range(n_it):
        src, lab = make_inputs(400 + it, 2, 49, 65)
        tgt, _ = make_inputs(500 + it, 2, 49, 65)
        loader.append({'src_image': src, 'src_label': lab, 'tgt_image': tgt})
    seen = []

    def recorded(fn):
        def wrapper(*a):
            out = fn(*a)
            seen.append(out.item())
            return out
        return wrapper

    quiet = types.SimpleNamespace(add_scalar=lambda *a: None, visualize_image=lambda *a: None)
    saved = []
    # train_adapt.py:87-88 wraps the model in nn.DataParallel, which on a host without GPUs calls the module directly;
    # with no_val the epoch ends by handing the checkpoint dict of :202-209 to Saver.save_checkpoint
    trainer = types.SimpleNamespace(model=torch.nn.DataParallel(G), saver=types.SimpleNamespace(save_checkpoint=lambda st, best: saved.append(st)),
                                    model_D=D, optimizer=opt, optimizer_D=opt_d, train_loader=loader,
                                    scheduler=RefSched('poly', lr, 1, n_it), best_pred=0.0,
                                    criterion=recorded(RefSegLoss().build_loss('ce')),
                                    bce_loss=recorded(torch.nn.BCEWithLogitsLoss()), writer=quiet, summary=quiet,
                                    args=types.SimpleNamespace(cuda=False, batch_size=2, dataset='gtav2cityscapes', no_val=True))
    old_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        training(trainer, 0)
    finally:
        torch.Tensor.cuda = old_cuda
    ref_hist = np.array(seen, dtype=np.float64).reshape(n_it, 4)     # loss_seg, loss_adv, loss_D(src), loss_D(tgt)
    cfg = O.BNCfg(True)
    o_hist = []
    for it, b in enumerate(loader):
        for o in (o_opt, o_opt_d):
            for gi, grp in enumerate(o.param_groups):
                grp['lr'] = O.poly_lr(lr, it, n_it) * (10 if gi > 0 else 1)
        o_hist.append(O.adapt_step(g_sd, d_sd, o_opt, o_opt_d, b['src_image'], b['src_label'], b['tgt_image'], cfg, drop=False))
    o_hist = np.array(o_hist, dtype=np.float64)
    print('adapt loop: reference', ref_hist[[0, 1, n_it - 1]].tolist(), 'oracle', o_hist[[0, 1, n_it - 1]].tolist())
    assert np.allclose(ref_hist, o_hist, rtol=1e-3, atol=1e-5), np.abs(ref_hist / o_hist - 1).max()
    fix = dict(losses=ref_hist)
    for k in ['backbone.features.0.0.weight', 'decoder.last_conv.8.weight']:
        w = dict(G.named_parameters())[k].detach()
        assert relerr(g_sd[k].detach(), w) < 1e-3, (k, relerr(g_sd[k].detach(), w))
        fix['w:' + k] = head(w)
    assert relerr(d_sd['conv1.weight'].detach(), D.conv1.weight.detach()) < 1e-3
    fix['wd:conv1.weight'] = head(D.conv1.weight)
    assert len(saved) == 1
    fix['checkpoint_layout'] = np.array(checkpoint_layout(saved[0]))
    np.savez_compressed(os.path.join(HERE, 'adapt_loop.npz'), **fix)


def feature_loop_case():
    """BASELINE config 4 on the reference's own code: `Trainer.training` of train.py (:152-232, unmodified source
    segment) for one epoch of ten iterations on the CPU with the reference's modules, DomainLosses, LR_Scheduler and
    the script's default Adam optimizers (train.py:76-80), against ten oracle feature_step iterations."""
    import types
    import torch.nn.functional as F
    from utils.lr_scheduler import LR_Scheduler as RefSched
    nn = torch.nn

    class Bar(list):
        def set_description(self, text):
            pass

    training = reference_method('train.py', 'training', {'np': np, 'torch': torch, 'F': F, 'tqdm': lambda it: Bar(it)})
    torch.manual_seed(7)
    bb = no_dropout(ref_mobilenet.MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d))
    aspp = no_dropout(RefASPP(backbone='mobilenet', output_stride=16, BatchNorm=nn.BatchNorm2d))
    dec = no_dropout(RefDecoder(num_classes=19, backbone='mobilenet', BatchNorm=nn.BatchNorm2d))
    dc = no_dropout(RefDC(backbone='mobilenet', BatchNorm=nn.BatchNorm2d))
    sds = [clone_sd(m) for m in (bb, aspp, dec, dc)]
    lr, n_it = 5e-4, 10
    f_params = list(bb.parameters()) + list(aspp.parameters())
    mk = lambda ps: torch.optim.Adam(ps, lr=lr)  # noqa: E731
    o_fp = list(O.leaf_params(sds[0]).values()) + list(O.leaf_params(sds[1]).values())
    o_opts = (mk(o_fp + list(O.leaf_params(sds[2]).values())), mk(list(O.leaf_params(sds[3]).values())), mk(o_fp))
    loader = []
    g = torch.Generator().manual_seed(13)
    for it in range(n_it):
        loader.append({'src_image': torch.randn(2, 3, 48, 64, generator=g), 'tgt_image': torch.randn(2, 3, 48, 64, generator=g),
                       'src_label': torch.randint(0, 19, (2, 48, 64), generator=g).float()})
    seen = []

    def recorded(fn):
        def wrapper(*a):
            out = fn(*a)
            seen.append(out.item() if torch.is_tensor(out) else (out[0].item(), out[1]))
            return out
        return wrapper

    quiet = types.SimpleNamespace(add_scalar=lambda *a: None, visualize_image=lambda *a: None)
    trainer = types.SimpleNamespace(
        backbone_model=bb, assp_model=aspp, y_model=dec, d_model=dc,
        task_optimizer=mk(f_params + list(dec.parameters())), d_optimizer=mk(list(dc.parameters())),
        d_inv_optimizer=mk(f_params), c_optimizer=mk(f_params + list(dec.parameters())),
        scheduler=RefSched('poly', lr, 1, n_it), best_pred=0.0, train_loader=loader,
        task_loss=recorded(RefSegLoss().build_loss('ce')), domain_loss=recorded(RefDomLoss().build_loss()),
        writer=quiet, summary=quiet, args=types.SimpleNamespace(cuda=False, batch_size=2, dataset='gtav2cityscapes', no_val=False))
    training(trainer, 0)
    # per iteration the body calls task_loss, domain_loss(src, tgt), domain_loss(tgt, src)
    ref_hist = np.array([[seen[3 * i], seen[3 * i + 1][0], seen[3 * i + 2][0], seen[3 * i + 1][1]] for i in range(n_it)])
    o_hist = []
    for it, b in enumerate(loader):
        for o in o_opts:
            o.param_groups[0]['lr'] = O.poly_lr(lr, it, n_it)
        o_hist.append(O.feature_step(sds[0], sds[1], sds[2], sds[3], o_opts, b['src_image'], b['src_label'], b['tgt_image'],
                                     O.BNCfg(True), drop=False))
    o_hist = np.array(o_hist, dtype=np.float64)
    print('feature loop: reference', ref_hist[[0, n_it - 1]].tolist(), 'oracle', o_hist[[0, n_it - 1]].tolist())
    assert np.allclose(ref_hist, o_hist, rtol=2e-3, atol=1e-5), np.abs(ref_hist / o_hist - 1).max()
    for sd, m, k in ((sds[0], bb, 'features.0.0.weight'), (sds[2], dec, 'last_conv.8.weight'), (sds[3], dc, 'DC_adnn3.weight')):
        w = dict(m.named_parameters())[k].detach()
        assert relerr(sd[k].detach(), w) < 2e-3, (k, relerr(sd[k].detach(), w))
    # train.py:254-314, unmodified: validation over two batches; the first epoch beats best_pred = 0, so it ends by
    # handing the four-model checkpoint dict of :300-313 to Saver.save_checkpoint (models wrapped in nn.DataParallel
    # as train.py:107-110 does -- a pass-through on a host without GPUs)
    class VBar(list):
        def set_description(self, text):
            pass

    validation = reference_method('train.py', 'validation', {'np': np, 'torch': torch, 'F': F, 'tqdm': lambda it, desc='': VBar(it)})
    saved = []
    for name in ('backbone_model', 'assp_model', 'y_model', 'd_model'):
        setattr(trainer, name, torch.nn.DataParallel(getattr(trainer, name)))
    trainer.saver = types.SimpleNamespace(save_checkpoint=lambda st, best: saved.append((st, best)))
    trainer.evaluator = RefEvaluator(19)
    trainer.val_loader = [{'image': x, 'label': lab} for x, lab in (make_inputs(600, 2, 48, 64), make_inputs(601, 1, 48, 64))]
    trainer.task_loss = RefSegLoss().build_loss('ce')
    validation(trainer, 0)
    assert len(saved) == 1 and saved[0][1] is True
    cm = np.zeros((19, 19), np.int64)
    with torch.no_grad():
        for b in trainer.val_loader:
            hi, lo = O.mobilenet_forward(sds[0], b['image'], O.BNCfg(False))
            out = torch.nn.functional.interpolate(O.decoder_forward(sds[2], O.aspp_forward(sds[1], hi, O.BNCfg(False)), lo, O.BNCfg(False)),
                                                  b['image'].shape[2:], mode='bilinear', align_corners=True)
            cm += O.confusion_matrix(b['label'].numpy(), np.argmax(out.numpy(), axis=1), 19)
    assert np.array_equal(cm, trainer.evaluator.confusion_matrix)
    assert saved[0][0]['best_pred'] == O.evaluator_metrics(cm)['mIoU']
    np.savez_compressed(os.path.join(HERE, 'feature_loop.npz'), losses=ref_hist, val_confusion_matrix=cm,
                        checkpoint_layout=np.array(checkpoint_layout(saved[0][0])))


def feature_single_loop_case():
    """train.py's single-domain branch (args.dataset == 'gtav', train.py:164-165,205-210) on the reference's own code:
    the unmodified `Trainer.training` for ten iterations on {'image', 'label'} samples -- task loss only, only
    task_optimizer steps, the domain classifier still runs forward on the source features (:187) -- against ten oracle
    feature_step(..., tgt_image=None) iterations: losses, updated weights, the (moved) domain-classifier BN statistics
    and its (unmoved) weights."""
    import types
    import torch.nn.functional as F
    from utils.lr_scheduler import LR_Scheduler as RefSched
    nn = torch.nn

    class Bar(list):
        def set_description(self, text):
            pass

    training = reference_method('train.py', 'training', {'np': np, 'torch': torch, 'F': F, 'tqdm': lambda it: Bar(it)})
    torch.manual_seed(17)
    bb = no_dropout(ref_mobilenet.MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d))
    aspp = no_dropout(RefASPP(backbone='mobilenet', output_stride=16, BatchNorm=nn.BatchNorm2d))
    dec = no_dropout(RefDecoder(num_classes=19, backbone='mobilenet', BatchNorm=nn.BatchNorm2d))
    dc = no_dropout(RefDC(backbone='mobilenet', BatchNorm=nn.BatchNorm2d))
    sds = [clone_sd(m) for m in (bb, aspp, dec, dc)]
    dc_w0 = dc.DC_adnn3.weight.detach().clone()
    lr, n_it = 5e-4, 10
    f_params = list(bb.parameters()) + list(aspp.parameters())
    mk = lambda ps: torch.optim.Adam(ps, lr=lr)  # noqa: E731
    o_fp = list(O.leaf_params(sds[0]).values()) + list(O.leaf_params(sds[1]).values())
    o_opts = (mk(o_fp + list(O.leaf_params(sds[2]).values())), mk(list(O.leaf_params(sds[3]).values())), mk(o_fp))
    loader = []
    g = torch.Generator().manual_seed(19)
    for it in range(n_it):
        loader.append({'image': torch.randn(2, 3, 48, 64, generator=g),
                       'label': torch.randint(0, 19, (2, 48, 64), generator=g).float()})
    seen = []

    def recorded(fn):
        def wrapper(*a):
            out = fn(*a)
            seen.append(out.item())
            return out
        return wrapper

    def no_domain_loss(*a):
        raise AssertionError('the single-domain branch must not evaluate the domain loss')

    quiet = types.SimpleNamespace(add_scalar=lambda *a: None, visualize_image=lambda *a: None)
    trainer = types.SimpleNamespace(
        backbone_model=bb, assp_model=aspp, y_model=dec, d_model=dc,
        task_optimizer=mk(f_params + list(dec.parameters())), d_optimizer=mk(list(dc.parameters())),
        d_inv_optimizer=mk(f_params), c_optimizer=mk(f_params + list(dec.parameters())),
        scheduler=RefSched('poly', lr, 1, n_it), best_pred=0.0, train_loader=loader,
        task_loss=recorded(RefSegLoss().build_loss('ce')), domain_loss=no_domain_loss,
        writer=quiet, summary=quiet, args=types.SimpleNamespace(cuda=False, batch_size=2, dataset='gtav', no_val=False))
    training(trainer, 0)
    ref_hist = np.array(seen)
    o_hist = []
    for it, b in enumerate(loader):
        for o in o_opts:
            o.param_groups[0]['lr'] = O.poly_lr(lr, it, n_it)
        o_hist.append(O.feature_step(sds[0], sds[1], sds[2], sds[3], o_opts, b['image'], b['label'], None, O.BNCfg(True),
                                     drop=False)[0])
    o_hist = np.array(o_hist, dtype=np.float64)
    print('single-domain feature loop: reference', ref_hist.tolist(), 'oracle', o_hist.tolist())
    assert np.allclose(ref_hist, o_hist, rtol=2e-3, atol=1e-5), np.abs(ref_hist / o_hist - 1).max()
    fix = {'losses': ref_hist}
    for sd, m, k in ((sds[0], bb, 'features.0.0.weight'), (sds[1], aspp, 'conv1.weight'), (sds[2], dec, 'last_conv.8.weight'),
                     (sds[3], dc, 'DC_adnn1.1.running_mean'), (sds[3], dc, 'DC_adnn3.weight')):
        w = m.state_dict()[k].detach()
        assert relerr(sd[k].detach(), w) < 2e-3, (k, relerr(sd[k].detach(), w))
        fix['w:' + k] = head(w)
    assert torch.equal(dc.DC_adnn3.weight.detach(), dc_w0)                  # d_optimizer never stepped
    assert float(dc.DC_adnn1[1].running_mean.abs().sum()) > 0               # ... but its forward ran (train.py:187)
    np.savez_compressed(os.path.join(HERE, 'feature_single_loop.npz'), **fix)


def validation_case():
    """BASELINE config 5: the reference's own `Trainer.validation` (val_adapt.py:117-175, unmodified) run on the CPU
    over three batches (2 + 2 + 1 images) with the reference's DeepLab, criterion and Evaluator; it appends its report
    to val_info.txt.  Checked against it: the oracle's eval forward + argmax + confusion matrix + metrics and summed
    loss, and the product's report text (utils.report.validation_report) character for character."""
    import importlib
    import tempfile
    import types

    class Bar(list):
        def set_description(self, text):
            pass

    validation = reference_method('val_adapt.py', 'validation', {'np': np, 'torch': torch, 'tqdm': lambda it, desc='': Bar(it)})
    torch.manual_seed(1)
    ref = RefDeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    # give the BatchNorms non-trivial running statistics, as after training
    g = torch.Generator().manual_seed(5)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    sd = clone_sd(ref, grad=False)
    batches = []
    for k, n in enumerate((2, 2, 1)):
        x, lab = make_inputs(300 + k, n, 65, 97)
        batches.append({'image': x, 'label': lab})
    trainer = types.SimpleNamespace(model=ref, evaluator=RefEvaluator(19), val_loader=batches,
                                    criterion=RefSegLoss().build_loss('ce'), args=types.SimpleNamespace(cuda=False, batch_size=2))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            validation(trainer, 3)
            text = open('val_info.txt').read()
        finally:
            os.chdir(cwd)
    # oracle: val_adapt.py:122-135 restated
    cm = np.zeros((19, 19), np.int64)
    test_loss = 0.0
    with torch.no_grad():
        for b in batches:
            out = O.deeplab_forward(sd, b['image'], O.BNCfg(False), 16)
            test_loss += O.seg_cross_entropy(out, b['label']).item()
            cm += O.confusion_matrix(b['label'].numpy(), np.argmax(out.numpy(), axis=1), 19)
    assert np.array_equal(cm, trainer.evaluator.confusion_matrix)
    m = O.evaluator_metrics(cm)
    assert m['mIoU'] == trainer.evaluator.Mean_Intersection_over_Union()[0]
    rep = importlib.import_module('synthetic-to-real-semantic-segmentation_b200.utils.report')
    mine = rep.validation_report(trainer.evaluator, 3, 5, test_loss)
    assert mine == text, (mine, text)
    print(text)
    np.savez_compressed(os.path.join(HERE, 'validation.npz'), text=np.array(text), confusion_matrix=cm,
                        test_loss=np.float64(test_loss), epoch=np.int64(3), num_images=np.int64(5))


def shapes_case():
    """The output shapes of the reference's __main__ smoke blocks."""
    torch.manual_seed(0)
    with torch.no_grad():
        m = ref_mobilenet.MobileNetV2(output_stride=16, BatchNorm=torch.nn.BatchNorm2d)
        hi, lo = m(torch.rand(1, 3, 512, 512))
        assert tuple(hi.shape) == (1, 320, 32, 32) and tuple(lo.shape) == (1, 24, 128, 128)
        a = RefASPP('mobilenet', 16, torch.nn.BatchNorm2d).eval()
        assert tuple(a(torch.rand(2, 320, 32, 32)).shape) == (2, 256, 32, 32)
        d = RefDecoder(19, 'mobilenet', torch.nn.BatchNorm2d).eval()
        assert tuple(d(torch.rand(1, 256, 32, 32), torch.rand(1, 24, 128, 128)).shape) == (1, 19, 128, 128)
        assert tuple(RefD(19)(torch.rand(1, 19, 512, 512)).shape) == (1, 1, 16, 16)
    print('shapes ok')


def policy_case():
    """utils/lr_scheduler.py (poly / cos / step, warm-up, 1 and 2+ parameter groups -- the second call also overwrites
    the discriminator's Adam lr, train_adapt.py:131-133) and SegmentationLosses.FocalLoss (utils/loss.py:32-46) of the
    reference, on fixed inputs."""
    import contextlib
    import io
    from utils.lr_scheduler import LR_Scheduler as RefSched

    class Opt(object):
        def __init__(self, n):
            self.param_groups = [{'lr': -1.0} for _ in range(n)]

    rows = []
    cfgs = [('poly', 5e-4, 4, 25, 0, 0), ('cos', 1e-3, 3, 10, 0, 0), ('step', 7e-3, 6, 5, 2, 0), ('poly', 2.5e-4, 5, 8, 0, 2)]
    for ci, (mode, lr, epochs, ipe, lr_step, warm) in enumerate(cfgs):
        with contextlib.redirect_stdout(io.StringIO()):
            sch = RefSched(mode, lr, epochs, ipe, lr_step=lr_step, warmup_epochs=warm)
            for ngroups in (1, 2, 3):
                for epoch in range(epochs):
                    for i in range(0, ipe, 3):
                        o = Opt(ngroups)
                        sch(o, i, epoch, 0.0)
                        rows.append([ci, ngroups, epoch, i] + [g['lr'] for g in o.param_groups] + [0.0] * (3 - ngroups))
    fix = dict(sched_cfgs=np.array([[0 if m == 'poly' else 1 if m == 'cos' else 2, lr, e, ipe, st, w] for m, lr, e, ipe, st, w in cfgs],
                                   dtype=np.float64), sched_rows=np.array(rows, dtype=np.float64))
    g = torch.Generator().manual_seed(41)
    logit = torch.randn(2, 19, 9, 11, generator=g, requires_grad=True)
    lab = torch.randint(0, 20, (2, 9, 11), generator=g).float()
    lab[lab == 19] = 255
    for tag, wgt in (('', None), ('_w', torch.rand(19, generator=g) + 0.5)):
        logit.grad = None
        loss = RefSegLoss(weight=wgt).build_loss('focal')(logit, lab)
        loss.backward()
        fix['focal_loss' + tag] = np.float64(loss.item())
        fix['focal_grad' + tag] = logit.grad.numpy().copy()
        if wgt is not None:
            fix['focal_weight'] = wgt.numpy()
    fix['focal_logit'], fix['focal_label'] = logit.detach().numpy(), lab.numpy()
    np.savez_compressed(os.path.join(HERE, 'policy.npz'), **fix)
    print('policy fixtures:', len(rows), 'schedule points; focal', fix['focal_loss'], fix['focal_loss_w'])


def config1_case():
    """BASELINE.json configs[0] / SURVEY.md section 8(d) config 1: DeepLab('mobilenet', 16, 19).train() under
    torch.manual_seed(1), batch 2x3x513x513 from Generator(0), forward + CE + backward on the CPU (dropout off so the
    run is reproducible).  Only summaries are stored (the logits alone would be 40 MB)."""
    torch.manual_seed(1)
    ref = no_dropout(RefDeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)).train()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 513, 513, generator=g)
    lab = torch.randint(0, 20, (2, 513, 513), generator=g).float()
    lab[lab == 19] = 255
    out = ref(x)
    loss = RefSegLoss().build_loss('ce')(out, lab)
    loss.backward()
    names = [k for k, p in ref.named_parameters()]
    fix = dict(loss=np.float64(loss.item()), logits_head=head(out, 8192),
               class_mean=out.detach().double().mean((0, 2, 3)).numpy(), class_std=out.detach().double().std((0, 2, 3)).numpy(),
               grad_norm_names=np.array(names),
               grad_norms=np.array([float(p.grad.double().norm()) for k, p in ref.named_parameters()], dtype=np.float64))
    print('config1 loss', loss.item(), 'logits std', float(out.std()))
    np.savez_compressed(os.path.join(HERE, 'config1_2x513x513.npz'), **fix)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'config1':
        config1_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'policy':
        policy_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'featureloop':
        feature_loop_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'featuresingle':
        feature_single_loop_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'adaptloop':
        adapt_loop_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'validation':
        validation_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'syncbn':
        sync_bn_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'feature':
        feature_step_case()
        sys.exit(0)
    shapes_case()
    evaluator_case()
    discriminator_case()
    domain_case()
    deeplab_case('deeplab_train_2x65x97', 2, 65, 97, True)
    deeplab_case('deeplab_eval_1x97x65', 1, 97, 65, False)
    adapt_step_case()
    adapt_loop_case()
    feature_loop_case()
    feature_single_loop_case()
    sync_bn_case()
    validation_case()
    feature_step_case()
    config1_case()
    policy_case()
    print('all golden fixtures written to', HERE)
