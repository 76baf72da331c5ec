"""Well-posed end-to-end parity on the GPU:
  * teacher-forced stage-wise forward of the composed DeepLab at a BASELINE-shaped input (2x3x512x1024): every one
    of the 20 stages <= 1e-2 against the fp32 oracle, batch statistics of all 60 BatchNorm layers <= 1e-2;
  * the composed (not teacher-forced) train-mode network against the bf16-emulated oracle on statistics that are not
    chaotic: the loss and the batch statistics of every BatchNorm layer;
  * the execution path that bench.py times -- AdaptStep.capture -> stage -> replay_staged -- against the eager step:
    losses, weight DELTAS and BatchNorm buffers over five replays issued without a host synchronisation, with an
    eager-vs-eager control run as the yardstick for run-to-run noise (atomics order);
  * a captured ValStep replayed after captured training steps sees the updated weights.
"""
import numpy as np
import pytest
import torch

from conftest import sub
from emul import emulate_bf16
from oracle import ref_port as O
import stagewise as S

pytestmark = pytest.mark.gpu


def rel(a, b):
    return S.rel(a, b)


def make_deeplab(seed=1):
    torch.manual_seed(seed)
    m = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    m._s2r_no_dropout = True
    return m


def test_teacher_forced_stages_and_bn_statistics_at_512x1024(built_lib):
    m = make_deeplab()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.cuda().train()
    x = torch.randn(2, 3, 512, 1024, generator=torch.Generator().manual_seed(0))
    tr, o_out = S.oracle_trace(sd, x, training=True)
    errs = S.run_stages(m, x, tr, o_out)
    print("stage-wise rel-L2 vs fp32 oracle (teacher-forced, 2x3x512x1024):",
          {k: round(v, 5) for k, v in errs.items()})
    assert len(errs) == 20 and max(errs.values()) <= 1e-2, {k: v for k, v in errs.items() if v > 1e-2}
    stats = S.bn_stat_errors(m, sd)
    assert len(stats) == 60, len(stats)
    worst = sorted(stats.items(), key=lambda kv: -max(kv[1]))[:4]
    print("BatchNorm batch statistics vs fp32 oracle, worst layers (mean err, var err):", worst)
    assert max(max(v) for v in stats.values()) <= 1e-2, worst


def test_composed_train_network_loss_and_bn_statistics_vs_emulation(built_lib):
    """The whole network in train mode, NOT teacher-forced, against the oracle evaluated with the same bf16 storage
    points (tests/emul.py): element-wise logits are chaotic (module docstring of test_gpu_modules.py) but the batch
    statistics of every BatchNorm layer (averages over 10^3..10^6 samples) and the loss are not."""
    m = make_deeplab()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.cuda().train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 256, 512, generator=g)
    lab = torch.randint(0, 20, (2, 256, 512), generator=g).float()
    lab[lab == 19] = 255
    with emulate_bf16(), torch.no_grad():
        e_out = O.deeplab_forward(sd, x, O.BNCfg(True), 16, drop=False)
        e_loss = float(O.seg_cross_entropy(e_out, lab))
    out = m(x.cuda())
    loss = float(sub("utils.loss").SegmentationLosses().build_loss('ce')(out, lab.cuda()))
    stats = S.bn_stat_errors(m, sd)
    worst = sorted(stats.items(), key=lambda kv: -max(kv[1]))[:4]
    med = sorted(max(v) for v in stats.values())[len(stats) // 2]
    print("composed train-mode network vs bf16-emulated oracle: loss %.5f / %.5f, logits rel-L2 %.3f, BN statistics "
          "median %.4f worst" % (loss, e_loss, rel(out.detach(), e_out), med), worst)
    assert abs(loss - e_loss) <= 1e-2 * e_loss
    assert len(stats) == 60 and med <= 1e-2
    # the deepest layers inherit the amplified activation differences: their statistics are bounded, not pinned
    assert max(max(v) for v in stats.values()) <= 1e-1, worst


def _adapt_inputs(it, B, H, W):
    g = torch.Generator().manual_seed(1000 + it)
    src = torch.randn(B, 3, H, W, generator=g)
    tgt = torch.randn(B, 3, H, W, generator=g)
    lab = torch.randint(0, 20, (B, H, W), generator=g).float()
    lab[lab == 19] = 255
    return src, lab, tgt


_KEYS_G = ('decoder.last_conv.8.weight', 'decoder.last_conv.8.bias', 'decoder.last_conv.4.weight', 'aspp.conv1.weight',
           'backbone.features.17.conv.6.weight', 'backbone.features.0.0.weight')
_KEYS_D = ('conv1.weight', 'conv4.weight', 'classifier.weight', 'classifier.bias')
_KEYS_BN = ('backbone.features.0.1.running_mean', 'backbone.features.1.conv.1.running_var',
            'decoder.last_conv.5.running_mean', 'decoder.last_conv.5.running_var')


def _adapt_run(mode, n_it, B, H, W):
    """n_it iterations of the adaptation step from the same seeded weights on the same inputs.
    mode 'eager': AdaptStep.__call__; 'graph': iteration 0 eager inside capture(), then stage() / replay_staged()."""
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    G._s2r_no_dropout = True
    G.cuda().train()
    D.cuda().train()
    init = {k: v.detach().clone() for k, v in list(G.named_parameters()) + [('D.' + k, p) for k, p in D.named_parameters()]}
    step = sub("steps").AdaptStep(G, D, lr=5e-4, epochs=1, iters_per_epoch=10)
    names = ('loss_seg', 'loss_adv', 'loss_D_src', 'loss_D_tgt')
    losses = []
    if mode == 'eager':
        for it in range(n_it):
            src, lab, tgt = (t.cuda() for t in _adapt_inputs(it, B, H, W))
            out = step(src, lab, tgt, i=it, epoch=0)
            losses.append(torch.stack([out[k] for k in names]).clone())
    else:
        src, lab, tgt = (t.cuda() for t in _adapt_inputs(0, B, H, W))
        l0 = sub("_lib").launches
        step.capture(src, lab, tgt, warmup=1)                    # iteration 0 runs eagerly in here
        assert sub("_lib").launches > l0
        losses.append(None)
        host = [tuple(t.pin_memory() for t in _adapt_inputs(it, B, H, W)) for it in range(1, n_it)]
        for it in range(1, n_it):                                 # no host synchronisation inside this loop
            step.stage(*host[it - 1])
            out = step.replay_staged(i=it, epoch=0)
            losses.append(torch.stack([out[k] for k in names]).clone())
    torch.cuda.synchronize()
    losses = [None if t is None else t.cpu().numpy() for t in losses]
    params = dict(list(G.named_parameters()) + [('D.' + k, p) for k, p in D.named_parameters()])
    delta = {k: (params[k].detach() - init[k]).double().cpu() for k in list(_KEYS_G) + ['D.' + k for k in _KEYS_D]}
    bufs = {k: G.state_dict()[k].double().cpu() for k in _KEYS_BN}
    return losses, delta, bufs, (step.optimizer.steps, step.optimizer_D.steps)


def test_adapt_step_capture_stage_replay_matches_eager(built_lib):
    """The path bench.py times against the eager step, 1 eager + 5 replayed iterations on changing inputs.  Yardstick:
    a second eager run from the same seed (run-to-run noise of the fp32 atomics in the weight gradients, amplified by
    the network) -- the graph path must be as close to the eager run as the eager run is to itself (x5, with floors).
    Catches: stale learning rates / Adam bias corrections in the replayed optimizer kernels (hyper-
    parameter race), stale bf16 filter copies, missing kernels in the capture, wrong step counts."""
    n_it, B, H, W = 6, 4, 64, 96
    e1 = _adapt_run('eager', n_it, B, H, W)
    e2 = _adapt_run('eager', n_it, B, H, W)
    gr = _adapt_run('graph', n_it, B, H, W)
    assert gr[3] == e1[3] == (n_it, n_it), (gr[3], e1[3])          # the capture itself does not count as a step

    def dist(a, b):
        dl = max(float(np.max(np.abs(a[0][it] - b[0][it]) / (np.abs(b[0][it]) + 1e-3))) for it in range(1, n_it))
        dd = {k: rel(a[1][k], b[1][k]) for k in b[1]}
        db = {k: rel(a[2][k], b[2][k]) for k in b[2]}
        return dl, dd, db

    c_l, c_d, c_b = dist(e2, e1)
    g_l, g_d, g_b = dist(gr, e1)
    print("losses: graph-vs-eager %.2e (eager-vs-eager %.2e)" % (g_l, c_l))
    print("weight deltas graph-vs-eager:", {k: round(v, 4) for k, v in g_d.items()})
    print("weight deltas eager-vs-eager:", {k: round(v, 4) for k, v in c_d.items()})
    print("BN buffers graph-vs-eager:", {k: "%.1e" % v for k, v in g_b.items()}, "control", {k: "%.1e" % v for k, v in c_b.items()})
    assert g_l <= max(5 * c_l, 2e-3), (g_l, c_l)
    for k in g_d:
        assert g_d[k] <= max(5 * c_d[k], 3e-2), (k, g_d[k], c_d[k])
        assert float(e1[1][k].norm()) > 0, k                      # the optimizers moved every checked tensor
    for k in g_b:
        assert g_b[k] <= max(5 * c_b[k], 2e-3), (k, g_b[k], c_b[k])


def test_val_graph_after_training_replays_sees_updated_weights(built_lib):
    """a ValStep graph captured BEFORE training replays with bf16 filter copies it packed in its warm-up.
    The fused optimizers change the fp32 masters behind torch's version counters; ValStep.replay must refresh the
    copies.  Checked against a fresh module that loads the trained state_dict (new copies by construction)."""
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    G._s2r_no_dropout = True
    G.cuda()
    D.cuda().train()
    g = torch.Generator().manual_seed(7)
    img = torch.randn(1, 3, 64, 96, generator=g).cuda()
    vlab = torch.randint(0, 19, (1, 64, 96), generator=g).float().cuda()
    G.eval()
    val = sub("steps").ValStep(G, 19).capture(img, vlab)
    val.replay(img, vlab)
    before = val.evaluator.confusion_matrix.copy()
    G.train()
    step = sub("steps").AdaptStep(G, D, lr=2e-2, epochs=1, iters_per_epoch=10)   # large lr: the predictions move
    src, lab, tgt = (t.cuda() for t in _adapt_inputs(0, 4, 64, 96))
    step.capture(src, lab, tgt, warmup=1)
    for it in range(1, 4):
        step.replay(src, lab, tgt, i=it, epoch=0)
    G.eval()
    val._evaluator.reset()
    val.replay(img, vlab)                                            # the graph captured before training
    got = val.evaluator.confusion_matrix.copy()
    torch.manual_seed(1)
    G2 = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False)
    G2.load_state_dict(G.state_dict())
    G2.cuda().eval()
    fresh = sub("steps").ValStep(G2, 19)
    fresh(img, vlab)
    want = fresh.evaluator.confusion_matrix
    assert np.array_equal(got, want), int(np.abs(got - want).sum())
    assert not np.array_equal(got, before), "training did not change any prediction: the check would be vacuous"
