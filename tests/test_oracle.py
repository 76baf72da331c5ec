"""The CPU oracle (oracle/ref_port.py, oracle/confusion_matrix.c) against the fixtures that
tests/golden/make_golden.py generated from the real reference.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, sub
from oracle import ref_port as O


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def seeded_state(kind):
    """Reference-initialised weights come from OUR module constructors under the same seed; that
    they coincide with the reference's is itself checked against the stored checksums."""
    if kind == 'deeplab':
        torch.manual_seed(1)
        m = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    elif kind == 'disc':
        torch.manual_seed(2)
        m = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    else:
        torch.manual_seed(4)
        m = sub("modeling.domian").DomainClassifer('mobilenet', torch.nn.BatchNorm2d)
    return m


def check_sums(m, fix):
    sd = m.state_dict()
    names = [str(n) for n in fix['param_names']]
    assert names == [k for k, v in sd.items() if v.dtype.is_floating_point]
    for n, (s, a) in zip(names, fix['param_sums']):
        v = sd[n].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), n
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), n


def grad_sd(m):
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for v in O.leaf_params(sd).values():
        v.requires_grad_(True)
    return sd


def test_module_init_matches_reference_checksums():
    check_sums(seeded_state('deeplab'), golden('deeplab_train_2x65x97'))
    check_sums(seeded_state('disc'), golden('discriminator'))
    check_sums(seeded_state('dc'), golden('domain_classifier'))


def test_state_dict_layout():
    m = seeded_state('deeplab')
    sd = m.state_dict()
    assert len(sd) == 668          # 668 entries for DeepLab
    for k in ('backbone.features.0.0.weight', 'backbone.features.1.conv.0.weight', 'aspp.aspp4.atrous_conv.weight',
              'aspp.global_avg_pool.1.weight', 'decoder.last_conv.8.bias', 'backbone.features.17.conv.7.running_var'):
        assert k in sd
    assert sum(p.numel() for p in m.parameters()) == 5815539
    one = sum(p.numel() for p in m.get_1x_lr_params())
    ten = sum(p.numel() for p in m.get_10x_lr_params())
    assert one + ten == 5815539 and one == sum(p.numel() for p in m.backbone.parameters())
    d = seeded_state('disc')
    assert sum(p.numel() for p in d.parameters()) == 2781121
    dc = seeded_state('dc')
    assert sum(p.numel() for p in dc.parameters()) == 9721858


def test_oracle_init_matches_module_layout():
    """oracle.init_deeplab / init_discriminator (what bench.py's CPU arm runs on, with nothing of the product package
    on that path): same keys in the same order, shapes, dtypes and initial distributions as the module tree."""
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    for want, got in ((G.state_dict(), O.init_deeplab()), (D.state_dict(), O.init_discriminator())):
        want = {k: v for k, v in want.items() if '.low_level_features.' not in k and '.high_level_features.' not in k}
        assert list(want.keys()) == list(got.keys())
        for k, v in want.items():
            assert v.shape == got[k].shape and v.dtype == got[k].dtype, k
            if v.dim() == 4 and v.numel() >= 4096:              # conv weights: same distribution
                assert abs(float(got[k].std()) / float(v.std()) - 1) < 0.1 and abs(float(got[k].mean())) < 0.1 * float(v.std()), k
            elif v.dim() <= 1 and (bool((v == 0).all()) or bool((v == 1).all())):
                assert torch.equal(v, got[k]), k                # BN parameters and buffers: ones / zeros
    # and it runs: one adaptation step of the oracle on those weights
    g_sd, d_sd = O.init_deeplab(), O.init_discriminator()
    for sd in (g_sd, d_sd):
        for v in O.leaf_params(sd).values():
            v.requires_grad_(True)
    one, ten = O.split_lr_groups(list(O.leaf_params(g_sd).keys()))
    opt = torch.optim.SGD([{'params': [g_sd[k] for k in one], 'lr': 5e-4}, {'params': [g_sd[k] for k in ten], 'lr': 5e-3}],
                          momentum=0.9, weight_decay=5e-4)
    opt_d = torch.optim.Adam(list(O.leaf_params(d_sd).values()), lr=1e-4, betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(0)
    src, tgt = torch.randn(2, 3, 64, 96, generator=g), torch.randn(2, 3, 64, 96, generator=g)
    lab = torch.randint(0, 19, (2, 64, 96), generator=g).float()
    losses = O.adapt_step(g_sd, d_sd, opt, opt_d, src, lab, tgt, O.BNCfg(True))
    assert all(np.isfinite(v) for v in losses) and 2.0 < losses[0] < 4.5 and 0.5 < losses[1] < 0.9


def test_deeplab_train_forward_backward():
    fix = golden('deeplab_train_2x65x97')
    sd = grad_sd(seeded_state('deeplab'))
    x, lab = torch.from_numpy(fix['x']), torch.from_numpy(fix['label'])
    out = O.deeplab_forward(sd, x, O.BNCfg(True), 16, drop=False)
    assert rel(out.detach(), fix['logits']) < 1e-5
    loss = O.seg_cross_entropy(out, lab)
    assert abs(loss.item() - float(fix['loss'])) < 1e-5
    loss.backward()
    for k in fix.files:
        if k.startswith('grad:'):
            g = sd[k[5:]].grad.reshape(-1)[:4096]
            assert rel(g, fix[k]) < 2e-3, k
        if k.startswith('buf:'):
            assert rel(sd[k[4:]], fix[k]) < 1e-5, k
    norms = dict(zip([str(n) for n in fix['grad_norm_names']], fix['grad_norms']))
    for k, v in O.leaf_params(sd).items():
        assert abs(float(v.grad.double().norm()) - norms[k]) <= 2e-3 * norms[k] + 1e-9, k


def test_deeplab_eval_forward():
    fix = golden('deeplab_eval_1x97x65')
    sd = grad_sd(seeded_state('deeplab'))
    out = O.deeplab_forward(sd, torch.from_numpy(fix['x']), O.BNCfg(False), 16)
    assert rel(out.detach(), fix['logits']) < 1e-5


def test_discriminator():
    fix = golden('discriminator')
    sd = grad_sd(seeded_state('disc'))
    x = torch.from_numpy(fix['x']).requires_grad_(True)
    out = O.discriminator_forward(sd, x)
    assert rel(out.detach(), fix['out']) < 1e-6
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.zeros_like(out))
    assert abs(loss.item() - float(fix['loss'])) < 1e-6
    loss.backward()
    assert rel(x.grad, fix['dx']) < 1e-5
    for k in fix.files:
        if k.startswith('grad:'):
            assert rel(sd[k[5:]].grad.reshape(-1)[:4096], fix[k]) < 1e-4, k


def test_domain_classifier_and_loss_known_answer():
    fix = golden('domain_classifier')
    sd = grad_sd(seeded_state('dc'))
    cfg = O.BNCfg(True)
    ps = O.domain_classifier_forward(sd, torch.from_numpy(fix['xs']), cfg, False)
    pt = O.domain_classifier_forward(sd, torch.from_numpy(fix['xt']), cfg, False)
    assert rel(ps.detach(), fix['ps']) < 1e-5 and rel(pt.detach(), fix['pt']) < 1e-5
    loss, acc = O.domain_loss(ps, pt)
    assert abs(loss.item() - float(fix['loss'])) < 1e-5 and abs(acc - float(fix['acc'])) < 1e-7
    # utils/loss.py:80-87: loss = 2 ln(1 + e^-1), accuracy 1
    a, b = torch.ones(1, 1, 7, 7), torch.zeros(1, 1, 7, 7)
    l, ac = O.domain_loss(torch.cat([a, b], 1), torch.cat([b, a], 1))
    assert abs(l.item() - 0.626523) < 1e-5 and ac == 1.0


def test_evaluator_numpy_and_c():
    fix = golden('evaluator')
    gt, pred = fix['gt'], fix['pred']
    cm = O.confusion_matrix(gt, pred, 19) + O.confusion_matrix(gt[:, ::-1].copy(), pred, 19)
    assert (cm == fix['cm']).all()
    m = O.evaluator_metrics(cm)
    assert m['mIoU'] == float(fix['mIoU']) and m['PA'] == float(fix['PA'])
    assert m['mPA'] == float(fix['mPA']) and m['fwIoU'] == float(fix['fwIoU'])
    assert np.array_equal(m['IoU'], fix['IoU'])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "liboracle_cm.so"))
    lib.oracle_confusion_matrix_f32.restype = ctypes.c_int64
    counts = np.zeros((19, 19), dtype=np.int64)
    for g in (gt, gt[:, ::-1].copy()):
        g = np.ascontiguousarray(g)
        bad = lib.oracle_confusion_matrix_f32(g.ctypes.data_as(ctypes.c_void_p), pred.ctypes.data_as(ctypes.c_void_p),
                                              ctypes.c_int64(g.size), 19, counts.ctypes.data_as(ctypes.c_void_p))
        assert bad == 0
    assert (counts == fix['cm']).all()
    # empty and all-ignored inputs
    assert O.confusion_matrix(np.zeros((0,), np.float32), np.zeros((0,), np.int64), 19).sum() == 0
    assert O.confusion_matrix(np.full((5,), 255, np.float32), np.zeros((5,), np.int64), 19).sum() == 0


def test_adapt_step_two_iterations():
    fix = golden('adapt_step')
    import importlib
    mk = importlib.import_module("tests.golden.make_golden") if False else None  # inputs are re-derived below
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    g_sd, d_sd = grad_sd(G), grad_sd(D)
    lr = 5e-4
    one, ten = O.split_lr_groups(list(O.leaf_params(g_sd).keys()))
    opt = torch.optim.SGD([{'params': [g_sd[k] for k in one], 'lr': lr}, {'params': [g_sd[k] for k in ten], 'lr': lr * 10}],
                          momentum=0.9, weight_decay=5e-4)
    opt_d = torch.optim.Adam(list(O.leaf_params(d_sd).values()), lr=1e-4, betas=(0.9, 0.99))

    def inputs(seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(2, 3, 65, 97, generator=g)
        lab = torch.randint(0, 20, (2, 65, 97), generator=g).float()
        lab[lab == 19] = 255
        return x, lab

    for it in range(2):
        src, lab = inputs(100 + it)
        tgt, _ = inputs(200 + it)
        for o in (opt, opt_d):
            for gi, grp in enumerate(o.param_groups):
                grp['lr'] = O.poly_lr(lr, it, 10) * (10 if gi > 0 else 1)
        losses = O.adapt_step(g_sd, d_sd, opt, opt_d, src, lab, tgt, O.BNCfg(True), drop=False)
        assert np.allclose(losses, fix['losses'][it], rtol=2e-4, atol=1e-6), (losses, fix['losses'][it])
    for k in fix.files:
        if k.startswith('w:'):
            assert rel(g_sd[k[2:]].detach().reshape(-1)[:4096], fix[k]) < 1e-4, k
    assert rel(d_sd['conv1.weight'].detach().reshape(-1)[:4096], fix['wd:conv1.weight']) < 1e-4


def test_adapt_loop_against_reference_training_run():
    """BASELINE configs 2/3: the loss history of the reference's own `Trainer.training` (train_adapt.py:115-196, run
    unmodified for ten iterations on the CPU by tests/golden/make_golden.py adapt_loop_case, where all ten are checked)
    against the oracle's adapt_step from the same seed -- the first three iterations here, to keep the suite short."""
    fix = golden('adapt_loop')
    assert fix['losses'].shape == (10, 4)
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    g_sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    d_sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    for sd in (g_sd, d_sd):
        for v in O.leaf_params(sd).values():
            v.requires_grad_(True)
    one, ten = O.split_lr_groups(list(O.leaf_params(g_sd).keys()))
    opt = torch.optim.SGD([{'params': [g_sd[k] for k in one], 'lr': 5e-4}, {'params': [g_sd[k] for k in ten], 'lr': 5e-3}],
                          momentum=0.9, weight_decay=5e-4)
    opt_d = torch.optim.Adam(list(O.leaf_params(d_sd).values()), lr=1e-4, betas=(0.9, 0.99))

    def inputs(seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(2, 3, 49, 65, generator=g)
        lab = torch.randint(0, 20, (2, 49, 65), generator=g).float()
        lab[lab == 19] = 255
        return x, lab

    for it in range(3):
        src, lab = inputs(400 + it)
        tgt, _ = inputs(500 + it)
        for o in (opt, opt_d):
            for gi, grp in enumerate(o.param_groups):
                grp['lr'] = O.poly_lr(5e-4, it, 10) * (10 if gi > 0 else 1)
        got = O.adapt_step(g_sd, d_sd, opt, opt_d, src, lab, tgt, O.BNCfg(True), drop=False)
        assert np.allclose(got, fix['losses'][it], rtol=1e-3, atol=1e-5), (it, got, fix['losses'][it])


def test_feature_loop_against_reference_training_run():
    """BASELINE config 4: the loss / domain-accuracy history of the reference's own `Trainer.training` of train.py (run
    unmodified for ten iterations on the CPU by tests/golden/make_golden.py feature_loop_case, where all ten are
    checked) against the oracle's feature_step -- the first two iterations here."""
    fix = golden('feature_loop')
    assert fix['losses'].shape == (10, 4)
    nn = torch.nn
    torch.manual_seed(7)
    mods = (sub("modeling.backbone.mobilenet").MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d),
            sub("modeling.assp").ASPP('mobilenet', 16, nn.BatchNorm2d),
            sub("modeling.decoder").Decoder(19, 'mobilenet', nn.BatchNorm2d),
            sub("modeling.domian").DomainClassifer('mobilenet', nn.BatchNorm2d))
    sds = []
    for mod in mods:
        sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
        for v in O.leaf_params(sd).values():
            v.requires_grad_(True)
        sds.append(sd)
    fp = list(O.leaf_params(sds[0]).values()) + list(O.leaf_params(sds[1]).values())
    opts = (torch.optim.Adam(fp + list(O.leaf_params(sds[2]).values()), lr=5e-4),
            torch.optim.Adam(list(O.leaf_params(sds[3]).values()), lr=5e-4), torch.optim.Adam(fp, lr=5e-4))
    g = torch.Generator().manual_seed(13)
    for it in range(2):
        src, tgt = torch.randn(2, 3, 48, 64, generator=g), torch.randn(2, 3, 48, 64, generator=g)
        lab = torch.randint(0, 19, (2, 48, 64), generator=g).float()
        for o in opts:
            o.param_groups[0]['lr'] = O.poly_lr(5e-4, it, 10)
        got = O.feature_step(sds[0], sds[1], sds[2], sds[3], opts, src, lab, tgt, O.BNCfg(True), drop=False)
        assert np.allclose(got, fix['losses'][it], rtol=2e-3, atol=1e-5), (it, got, fix['losses'][it])


def test_feature_single_domain_loop_against_reference_training_run():
    """train.py's single-domain branch (args.dataset == 'gtav', train.py:164-165,205-210): the task-loss history of the
    reference's own unmodified `Trainer.training` (ten iterations, tests/golden/make_golden.py
    feature_single_loop_case, which checks all ten and the final weights) against the oracle's
    feature_step(..., tgt_image=None) -- the first three iterations here, plus the branch's defining properties: the
    domain classifier's weights do not move, its BatchNorm statistics do (its forward still runs, train.py:187)."""
    fix = golden('feature_single_loop')
    assert fix['losses'].shape == (10,)
    nn = torch.nn
    torch.manual_seed(17)
    mods = (sub("modeling.backbone.mobilenet").MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d),
            sub("modeling.assp").ASPP('mobilenet', 16, nn.BatchNorm2d),
            sub("modeling.decoder").Decoder(19, 'mobilenet', nn.BatchNorm2d),
            sub("modeling.domian").DomainClassifer('mobilenet', nn.BatchNorm2d))
    sds = [grad_sd(mod) for mod in mods]
    dc_w0 = sds[3]['DC_adnn3.weight'].detach().clone()
    fp = list(O.leaf_params(sds[0]).values()) + list(O.leaf_params(sds[1]).values())
    opts = (torch.optim.Adam(fp + list(O.leaf_params(sds[2]).values()), lr=5e-4),
            torch.optim.Adam(list(O.leaf_params(sds[3]).values()), lr=5e-4), torch.optim.Adam(fp, lr=5e-4))
    g = torch.Generator().manual_seed(19)
    for it in range(3):
        img = torch.randn(2, 3, 48, 64, generator=g)
        lab = torch.randint(0, 19, (2, 48, 64), generator=g).float()
        for o in opts:
            o.param_groups[0]['lr'] = O.poly_lr(5e-4, it, 10)
        got = O.feature_step(sds[0], sds[1], sds[2], sds[3], opts, img, lab, None, O.BNCfg(True), drop=False)
        assert got[1:] == (0.0, 0.0, 0)
        assert abs(got[0] - fix['losses'][it]) <= 2e-3 * fix['losses'][it], (it, got, fix['losses'][it])
    assert torch.equal(sds[3]['DC_adnn3.weight'].detach(), dc_w0)
    assert float(sds[3]['DC_adnn1.1.running_mean'].abs().sum()) > 0


def _layout_to_state(node):
    """Inverse of make_golden.checkpoint_layout with zero tensors of the recorded dtype / shape."""
    if isinstance(node, dict) and '__tensor__' in node:
        dtype, shape = node['__tensor__']
        return torch.zeros(shape, dtype=getattr(torch, dtype))
    if isinstance(node, dict) and '__items__' in node:
        return {(k['__int__'] if isinstance(k, dict) else k): _layout_to_state(v) for k, v in node['__items__']}
    if isinstance(node, list):
        return [_layout_to_state(v) for v in node]
    return node


def _skeleton(v):
    """Structure of a checkpoint dict: keys in order, tensor dtypes / shapes, integers, booleans and strings; floats
    (learning rates, best_pred, Adam's step counters) reduced to their type."""
    if torch.is_tensor(v):
        return ('tensor', str(v.dtype), tuple(v.shape)) if v.dim() else ('scalar-tensor',)
    if isinstance(v, dict):
        return [(k, _skeleton(x)) for k, x in v.items()]
    if isinstance(v, (list, tuple)):
        return [_skeleton(x) for x in v]
    return 'float' if isinstance(v, float) else v


def _fake_step(opt):
    for g in opt.param_groups:
        for p in g['params']:
            p.grad = torch.zeros_like(p)
    opt.step()


def test_checkpoint_layouts_against_the_reference_scripts():
    """SURVEY.md 8(f) row 2 on the reference's own code: the checkpoint dicts its unmodified `Trainer.training`
    (train_adapt.py:202-209) and `Trainer.validation` (train.py:300-313) handed to Saver.save_checkpoint, recorded as
    layouts (keys, order, tensor dtypes / shapes, parameter-group index lists) by tests/golden/make_golden.py.
    utils.checkpoint.adapt_state / feature_state on the product's modules produce the same structure, and load_adapt /
    load_feature accept the reference's (DataParallel-unwrapped) dicts.  CPU: torch optimizers stand in for the fused
    ones, whose torch.optim-format state is a GPU test."""
    import json
    ck_mod = sub("utils.checkpoint")
    nn = torch.nn
    # ---- train_adapt.py layout
    ref = _layout_to_state(json.loads(str(golden('adapt_loop')['checkpoint_layout'])))
    assert list(ref.keys()) == ['epoch', 'state_dict', 'optimizer', 'best_pred']
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    opt = torch.optim.SGD([{'params': list(G.get_1x_lr_params()), 'lr': 5e-4}, {'params': list(G.get_10x_lr_params()), 'lr': 5e-3}],
                          momentum=0.9, weight_decay=5e-4, nesterov=False)
    _fake_step(opt)
    mine = ck_mod.adapt_state(G, opt, 0, 0.0)
    assert _skeleton(mine) == _skeleton(ref)
    assert ck_mod.load_adapt(ref, G, opt) == (1, 0.0)
    assert float(G.decoder.last_conv[8].weight.detach().abs().sum()) == 0.0             # the (zero-filled) reference tensors are in
    assert ck_mod.load_adapt({**ref, 'state_dict': {'module.' + k: v for k, v in ref['state_dict'].items()}}, G, ft=True)[0] == 0
    # ---- train.py layout
    ref = _layout_to_state(json.loads(str(golden('feature_loop')['checkpoint_layout'])))
    mods = (sub("modeling.backbone.mobilenet").MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d),
            sub("modeling.assp").ASPP('mobilenet', 16, nn.BatchNorm2d), sub("modeling.decoder").Decoder(19, 'mobilenet', nn.BatchNorm2d),
            sub("modeling.domian").DomainClassifer('mobilenet', nn.BatchNorm2d))
    f_params = list(mods[0].parameters()) + list(mods[1].parameters())
    opts = [torch.optim.Adam(f_params + list(mods[2].parameters()), lr=5e-4), torch.optim.Adam(list(mods[3].parameters()), lr=5e-4),
            torch.optim.Adam(f_params, lr=5e-4)]
    for o in opts:
        _fake_step(o)
    c_opt = torch.optim.Adam(f_params + list(mods[2].parameters()), lr=5e-4)      # exists, never steps (train.py:73-75)
    mine = ck_mod.feature_state(*mods, *opts, 0, 0.5, c_optimizer=c_opt)
    assert _skeleton(mine) == _skeleton(ref)
    epoch, best = ck_mod.load_feature(ref, *mods, *opts)
    assert epoch == 1 and best == ref['best_pred'] and float(mods[3].DC_adnn3.weight.detach().abs().sum()) == 0.0


def test_validation_report_against_reference_run():
    """BASELINE config 5: the report the reference's own `Trainer.validation` (val_adapt.py:117-175, run unmodified on
    the CPU by tests/golden/make_golden.py validation_case) appended to val_info.txt, against the product's metric
    formulas (utils.metrics.Evaluator on the recorded counts) and report text (utils.report.validation_report),
    character for character; the oracle's metrics on the same matrix."""
    fix = golden('validation')
    ev = sub("utils.metrics").Evaluator(19)
    ev._counts = torch.from_numpy(fix['confusion_matrix'].astype(np.int64))       # as accumulated by the kernels
    ev._bad = torch.zeros(1, dtype=torch.int64)
    text = sub("utils.report").validation_report(ev, int(fix['epoch']), int(fix['num_images']), float(fix['test_loss']))
    assert text == str(fix['text'])
    m = O.evaluator_metrics(fix['confusion_matrix'])
    assert "mIoU:{}, fwIoU: {}".format(m['mIoU'], m['fwIoU']) in text and "Acc:{}, Acc_class:{},".format(m['PA'], m['mPA']) in text
    assert 0.9 * 5 * 65 * 97 < int(fix['confusion_matrix'].sum()) < 5 * 65 * 97      # valid pixels: all but the ~5 % ignored


def test_sync_batchnorm_against_reference_protocol_fixture():
    """The oracle's synchronised-BatchNorm branch (batchnorm.py:55-78,113-125 over the whole batch) against the
    reference's protocol executed with two replicas through its own SyncMaster / SlavePipe (tests/golden/
    make_golden.py sync_bn_case): outputs, running statistics, gradients -- including a zero-variance channel, where
    clamp(var, eps)^-1/2 and F.batch_norm's 1/sqrt(var + eps) differ."""
    fix = golden('sync_bn')
    sd = {'bn.weight': torch.from_numpy(fix['weight']).clone().requires_grad_(True),
          'bn.bias': torch.from_numpy(fix['bias']).clone().requires_grad_(True),
          'bn.running_mean': torch.zeros(6), 'bn.running_var': torch.ones(6)}
    x = torch.from_numpy(fix['x']).clone().requires_grad_(True)
    y = O.batch_norm(sd, 'bn', x, O.BNCfg(True, 0.1, 1e-5, sync_clamp=True))
    (y * torch.from_numpy(fix['dy'])).sum().backward()
    assert rel(y.detach(), fix['y']) < 1e-6 and rel(x.grad, fix['dx']) < 1e-5
    assert rel(sd['bn.weight'].grad, fix['dweight']) < 1e-5 and rel(sd['bn.bias'].grad, fix['dbias']) < 1e-5
    assert rel(sd['bn.running_mean'], fix['running_mean']) < 1e-6 and rel(sd['bn.running_var'], fix['running_var']) < 1e-6
    assert abs(float(sd['bn.running_var'][2]) - 0.9) < 1e-6           # the constant channel
    plain = O.batch_norm({k: v.detach().clone() for k, v in sd.items()}, 'bn', x.detach(), O.BNCfg(True, 0.1, 1e-5))
    assert rel(plain[:, [0, 1, 3, 4, 5]], fix['y'][:, [0, 1, 3, 4, 5]]) < 1e-5   # same normalisation elsewhere


def test_feature_step_two_iterations():
    """oracle.feature_step (train.py:173-204, BASELINE config 4) against the run of the reference's own MobileNetV2 /
    ASPP / Decoder / DomainClassifer modules and torch optimizers recorded in tests/golden/feature_step.npz (Adam -- the
    script default -- and SGD); the product's parameter containers, built under the same seed, start from the
    reference's initial weights."""
    fix = golden('feature_step')
    nn = torch.nn
    for flavour in ('Adam', 'SGD'):
        torch.manual_seed(7)
        mods = (sub("modeling.backbone.mobilenet").MobileNetV2(output_stride=16, BatchNorm=nn.BatchNorm2d),
                sub("modeling.assp").ASPP('mobilenet', 16, nn.BatchNorm2d),
                sub("modeling.decoder").Decoder(19, 'mobilenet', nn.BatchNorm2d),
                sub("modeling.domian").DomainClassifer('mobilenet', nn.BatchNorm2d))
        sds = []
        for mod in mods:
            sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
            for v in O.leaf_params(sd).values():
                v.requires_grad_(True)
            sds.append(sd)
        if flavour == 'Adam':
            mk = lambda ps: torch.optim.Adam(ps, lr=5e-4)  # noqa: E731
        else:
            mk = lambda ps: torch.optim.SGD(ps, lr=5e-4, momentum=0.9, weight_decay=5e-4)  # noqa: E731
        fp = list(O.leaf_params(sds[0]).values()) + list(O.leaf_params(sds[1]).values())
        opts = (mk(fp + list(O.leaf_params(sds[2]).values())), mk(list(O.leaf_params(sds[3]).values())), mk(fp))
        g = torch.Generator().manual_seed(11)
        for it in range(2):
            src = torch.randn(2, 3, 64, 96, generator=g)
            tgt = torch.randn(2, 3, 64, 96, generator=g)
            lab = torch.randint(0, 19, (2, 64, 96), generator=g).float()
            for o in opts:
                o.param_groups[0]['lr'] = O.poly_lr(5e-4, it, 10)
            got = O.feature_step(sds[0], sds[1], sds[2], sds[3], opts, src, lab, tgt, O.BNCfg(True), drop=False)
            assert np.allclose(got, fix['losses_' + flavour][it], rtol=2e-4, atol=1e-6), (flavour, it, got)
        for sd, k in ((sds[0], 'features.0.0.weight'), (sds[1], 'conv1.weight'), (sds[2], 'last_conv.8.weight'),
                      (sds[3], 'DC_adnn3.weight')):
            assert rel(sd[k].detach().reshape(-1)[:4096], fix['w_%s:%s' % (flavour, k)]) < 2e-4, (flavour, k)
        assert rel(sds[0]['features.0.1.running_mean'].reshape(-1)[:4096], fix['rm_%s:features.0.1.running_mean' % flavour]) < 1e-5


def test_input_stage_oracle_and_host_tables_against_reference_fixture():
    """oracle/input_stage.py (and the host-side resampling tables of the device input stage) against the outputs of
    the reference's own TrainSet/ValSet pipeline stored by tests/golden/make_golden_input.py."""
    from oracle import input_stage as OI
    fix = golden("input_stage")
    assert np.array_equal(OI.segmap_lut(), fix["lut"])
    for k in fix["cases"]:
        flip, short, crop, x1, y1 = (int(v) for v in fix[k + "_draw"])
        img, lab = OI.train_sample(fix[k + "_src"], fix[k + "_lab"], flip, short, crop, x1, y1)
        tgt, _ = OI.train_sample(fix[k + "_tgt"], fix[k + "_lab"], flip, short, crop, x1, y1)
        assert np.array_equal(img, fix[k + "_out_src"]) and np.array_equal(tgt, fix[k + "_out_tgt"]), k
        assert np.array_equal(lab, fix[k + "_out_lab"]), k
    s = int(fix["val_size"][0])
    assert np.array_equal(OI.normalize_to_tensor(OI.resize_bilinear(fix["val_img"], s, s)), fix["val_out_img"])
    assert np.array_equal(OI.resize_nearest(OI.encode_segmap(fix["val_lab"]), s, s).astype(np.float32), fix["val_out_lab"])
    dt = sub("dataloders.device_transforms")
    assert np.array_equal(dt.segmap_lut(), fix["lut"])
    for (i, o) in [(40, 64), (64, 40), (37, 111), (513, 257), (100, 7), (7, 100), (1024, 730)]:
        b, kk, ks = dt._bilinear_tables(i, o)
        b2, kk2 = OI.precompute_coeffs(i, o)
        assert ks == kk2.shape[1] and np.array_equal(b, b2) and np.array_equal(kk, kk2), (i, o)
        assert np.array_equal(dt._nearest_table(i, o), OI.nearest_table(i, o)), (i, o)


def test_gaussian_blur_oracle_against_reference_fixture_and_pillow():
    """RandomGaussianBlur (custom_transforms.py:92-105): the numpy restatement of Pillow's GaussianBlur (three
    fractional-radius box blurs per axis, libImaging/BoxBlur.c) against the tensors the reference's unmodified TrainSet
    produced on draws where the blur fires, against Pillow itself, and the host mirror's fixed-point weights against the
    oracle's over the reference's range of radii."""
    from oracle import input_stage as OI
    fix = golden("input_stage")
    assert len(fix["blur_cases"]) == 8
    for k in fix["blur_cases"]:
        flip, short, crop, x1, y1 = (int(v) for v in fix[k + "_draw"])
        r_src, r_tgt = (float(v) for v in fix[k + "_radii"])
        img, lab = OI.train_sample(fix[k + "_src"], fix[k + "_lab"], flip, short, crop, x1, y1, blur_radius=r_src)
        tgt, _ = OI.train_sample(fix[k + "_tgt"], fix[k + "_lab"], flip, short, crop, x1, y1, blur_radius=r_tgt)
        assert np.array_equal(img, fix[k + "_out_src"]) and np.array_equal(tgt, fix[k + "_out_tgt"]), k
        assert np.array_equal(lab, fix[k + "_out_lab"]), k
        plain, _ = OI.train_sample(fix[k + "_src"], fix[k + "_lab"], flip, short, crop, x1, y1)
        assert not np.array_equal(plain, img), k                       # the blur does something on these cases
    from PIL import Image, ImageFilter
    rng = np.random.RandomState(7)
    for trial in range(120):                # degenerate sizes and radii far beyond the reference's included
        h, w = int(rng.randint(1, 36)), int(rng.randint(1, 36))
        a = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        r = float(rng.rand()) * (1.0, 1.0, 5.0, 30.0)[trial % 4]
        want = np.array(Image.fromarray(a).filter(ImageFilter.GaussianBlur(radius=r)))
        assert np.array_equal(OI.gaussian_blur(a, r), want), (h, w, r)
    dt = sub("dataloders.device_transforms")
    for r in list(rng.rand(5000)) + [0.0, 1e-30, 1e-10, 0.999999, 1.0, 1.41]:
        ww, fw = dt._gaussian_blur_weights(float(r))
        fr = OI.gaussian_blur_radius(r) if r != 0 else 0
        if fr == 0:
            assert (ww, fw) == (1 << 24, 0), r      # identity: Pillow copies / skips the passes
        else:
            assert (0, ww, fw) == OI.box_blur_weights(fr), r
        assert ww + 2 * fw <= 1 << 24 and 255 * (ww + 2 * fw) + (1 << 23) < 1 << 32     # the kernel's uint32 arithmetic
    with pytest.raises(NotImplementedError):
        dt._gaussian_blur_weights(1.5)


def test_box_blur_c_oracle_against_pillow_and_numpy_oracle():
    """oracle/box_blur.c (the byte path restated in C) against Pillow's GaussianBlur and the numpy restatement, over the
    reference's range of radii and beyond, degenerate sizes included, and on a full 512 x 512 crop."""
    from PIL import Image, ImageFilter
    from oracle import input_stage as OI
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "liboracle_blur.so"))
    lib.gaussian_blur_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                     ctypes.c_uint32, ctypes.c_int]

    def c_blur(a, r):
        out = a.copy()
        fr = OI.gaussian_blur_radius(r) if r != 0 else 0
        if fr != 0:
            radius, ww, fw = OI.box_blur_weights(fr)
            assert lib.gaussian_blur_u8(out.ctypes.data, a.shape[0], a.shape[1], a.shape[2], radius, ww, fw, 3) == 0
        return out

    rng = np.random.RandomState(3)
    for trial in range(160):
        h, w = int(rng.randint(1, 40)), int(rng.randint(1, 40))
        a = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        r = float(rng.rand()) * (1.0, 1.0, 6.0, 40.0)[trial % 4]
        got = c_blur(a, r)
        assert np.array_equal(got, np.array(Image.fromarray(a).filter(ImageFilter.GaussianBlur(radius=r)))), (h, w, r)
        assert np.array_equal(got, OI.gaussian_blur(a, r)), (h, w, r)
    a = rng.randint(0, 256, (512, 512, 3)).astype(np.uint8)
    for r in (0.0, 1e-3, 0.5, 0.999999):
        assert np.array_equal(c_blur(a, r), np.array(Image.fromarray(a).filter(ImageFilter.GaussianBlur(radius=r)))), r


def test_resample_c_oracle_against_pillow_and_numpy_oracle():
    """oracle/resample.c (Pillow's BILINEAR / NEAREST resize restated in C) against Pillow itself and the numpy
    restatement over a sweep of shapes (1-pixel axes included) and on a GTA5-sized image."""
    from PIL import Image
    from oracle import input_stage as OI
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "liboracle_resample.so"))
    lib.resize_bilinear_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    lib.resize_nearest_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    rng = np.random.RandomState(5)
    for (h, w) in [(17, 23), (31, 90), (1, 7), (9, 1)]:
        a = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        m = rng.randint(0, 256, (h, w)).astype(np.uint8)
        for ow in (1, 5, 23, 36, 97):
            for oh in (1, 7, 17, 40):
                o = np.empty((oh, ow, 3), np.uint8)
                assert lib.resize_bilinear_u8(a.ctypes.data, h, w, 3, oh, ow, o.ctypes.data) == 0
                assert np.array_equal(o, np.array(Image.fromarray(a).resize((ow, oh), Image.BILINEAR))), (h, w, ow, oh)
                assert np.array_equal(o, OI.resize_bilinear(a, ow, oh)), (h, w, ow, oh)
                o2 = np.empty((oh, ow), np.uint8)
                lib.resize_nearest_u8(m.ctypes.data, h, w, oh, ow, o2.ctypes.data)
                assert np.array_equal(o2, np.array(Image.fromarray(m).resize((ow, oh), Image.NEAREST))), (h, w, ow, oh)
                assert np.array_equal(o2, OI.resize_nearest(m, ow, oh)), (h, w, ow, oh)
    a = rng.randint(0, 256, (1052, 1914, 3)).astype(np.uint8)
    o = np.empty((700, 1273, 3), np.uint8)
    assert lib.resize_bilinear_u8(a.ctypes.data, 1052, 1914, 3, 700, 1273, o.ctypes.data) == 0
    assert np.array_equal(o, np.array(Image.fromarray(a).resize((1273, 700), Image.BILINEAR)))


def test_loss_module_host_logic_against_reference_fixture(monkeypatch):
    """utils.loss host logic on the CPU -- the focal transform of the mean cross entropy (loss.py:32-46), the loss
    table of build_loss and its NotImplementedError -- with the one device call (functional.cross_entropy) replaced by
    its torch equivalent; values and gradients against the reference's FocalLoss recorded in tests/golden/policy.npz."""
    import torch.nn.functional as F
    loss_mod = sub("utils.loss")

    def ce(logit, target, weight=None, ignore_index=255, **kw):
        return F.cross_entropy(logit, target.long(), weight=weight, ignore_index=ignore_index, reduction='mean')

    monkeypatch.setattr(loss_mod, "cross_entropy", ce)
    fix = golden("policy")
    lab = torch.from_numpy(fix["focal_label"])
    for tag, wgt in (("", None), ("_w", torch.from_numpy(fix["focal_weight"]))):
        x = torch.from_numpy(fix["focal_logit"]).clone().requires_grad_(True)
        losses = loss_mod.SegmentationLosses(weight=wgt)
        loss = losses.build_loss('focal')(x, lab)
        loss.backward()
        assert abs(loss.item() - float(fix["focal_loss" + tag])) <= 1e-6 * abs(float(fix["focal_loss" + tag]))
        assert rel(x.grad, fix["focal_grad" + tag]) < 1e-6
        assert torch.equal(losses.build_loss('ce')(x.detach(), lab), ce(x.detach(), lab, weight=wgt))
    with pytest.raises(NotImplementedError):
        loss_mod.SegmentationLosses().build_loss('dice')


def test_lr_policy_against_reference_fixture():
    """utils.lr_scheduler.LR_Scheduler (poly / cos / step, warm-up, lr to group 0 and 10*lr to the others --
    utils/lr_scheduler.py:43-70, incl. the overwrite of the discriminator's lr at train_adapt.py:133) against the
    values the reference class produced (tests/golden/policy.npz); unknown modes raise like the reference."""
    fix = golden("policy")
    LR = sub("utils.lr_scheduler").LR_Scheduler

    class Opt(object):
        def __init__(self, n):
            self.param_groups = [{'lr': -1.0} for _ in range(n)]

    scheds = []
    for m, lr, e, ipe, st, w in fix["sched_cfgs"]:
        scheds.append(LR(['poly', 'cos', 'step'][int(m)], float(lr), int(e), int(ipe), lr_step=int(st), warmup_epochs=int(w)))
    for row in fix["sched_rows"]:
        ci, ng, epoch, i = (int(v) for v in row[:4])
        o = Opt(ng)
        scheds[ci](o, i, epoch, 0.0)
        assert [g['lr'] for g in o.param_groups] == list(row[4:4 + ng]), (ci, ng, epoch, i)
    import pytest
    with pytest.raises((NotImplementedError, TypeError)):
        LR('linear', 1e-3, 2, 5)(Opt(1), 0, 0)


def test_prediction_export_oracle_against_pillow():
    """oracle/report.py (test_adapt.py:118-157 after np.argmax) against the same statements executed with Pillow itself:
    labelId image and palette image, NEAREST-resized."""
    PIL = __import__("pytest").importorskip("PIL")
    from PIL import Image
    from oracle import report as OR
    rng = np.random.RandomState(5)
    logits = rng.randn(19, 40, 56).astype(np.float32)
    logits[4] = logits[9]                               # ties -> lowest index (np.argmax)
    ids, rgb = OR.imgsaver_arrays(logits, 150, 70)
    im1 = np.uint8(np.argmax(logits[None], axis=1).transpose(1, 2, 0)).squeeze()
    im1_np = np.uint8(np.zeros(im1.shape))
    im2_np = np.uint8(np.zeros(im1.shape + (3,)))
    for c in range(19):
        im1_np[im1 == c] = OR.VALID_CLASSES[c]
        im2_np[im1 == c] = OR.PALETTE[c]
    want_ids = np.array(Image.fromarray(im1_np, mode='L').resize((150, 70), Image.NEAREST))
    want_rgb = np.array(Image.fromarray(im2_np).resize((150, 70), Image.NEAREST))
    assert np.array_equal(ids, want_ids) and np.array_equal(rgb, want_rgb)


def test_prediction_export_oracle_against_reference_fixture():
    """oracle/report.py against the PNG files the reference's own `imgsaver` (test_adapt.py:118-157, compiled unmodified
    out of the script by tests/golden/make_golden_report.py) wrote for the same predictions."""
    import importlib.util
    from oracle import report as OR
    spec = importlib.util.spec_from_file_location("make_golden_report", os.path.join(ROOT, "tests", "golden", "make_golden_report.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fix = golden("report")
    for seed in fix["seeds"]:
        ids, rgb = OR.imgsaver_arrays(gen.logits(int(seed)))
        assert np.array_equal(ids, fix["ids%d" % seed]) and np.array_equal(rgb, fix["rgb%d" % seed]), seed
        assert len(np.unique(ids)) >= 15                       # a real mix of classes, labelIds of the valid classes only
        assert set(np.unique(ids)) <= set(OR.VALID_CLASSES)


def test_batched_input_stage_host_planning_with_emulated_kernels():
    """The REAL host planning of the batched device input stage (window of the scaled image per sample, job tables for
    the column / row / nearest / blur / crop launches) driven through a pure-Python emulation of the kernels reproduces
    the reference's tensors, RandomGaussianBlur cases included (tests/tools/emul_input_stage.py; no GPU)."""
    import sys as _sys
    r = subprocess.run([_sys.executable, os.path.join(ROOT, "tests", "tools", "emul_input_stage.py")], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ALL OK"), r.stdout[-2000:] + r.stderr[-2000:]
