"""Kernel-level parity on the GPU: every C-ABI entry point against a plain fp32 restatement of
the same op (torch ops on the same device, or the numpy/C oracle for the integer path).
Tolerances: bit-exact for the confusion matrix; bf16 storage (2^-9 relative per element) bounds
the floating-point kernels, stated per test as a relative L2 error."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, sub
from oracle import ref_port as O

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def eng(built_lib):
    assert torch.cuda.is_available()
    return sub("engine")


@pytest.fixture()
def cx(eng):
    return eng.Ctx(torch.device("cuda", 0), True)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf(t):
    return t.to(torch.bfloat16).float()


def nhwc_act(eng, t_nchw, pitch=None):
    """NCHW fp32 -> Act (bf16, channel pitch padded with zeros)"""
    N, Cc, H, W = t_nchw.shape
    pitch = pitch or eng.round_up(Cc, 8)
    buf = torch.zeros((N, H, W, pitch), dtype=torch.bfloat16, device=t_nchw.device)
    buf[..., :Cc] = t_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    a = eng.Act(buf)
    a.C = Cc
    return a


def to_nchw(a, Cc=None):
    Cc = Cc or a.C
    return a.t[..., a.off:a.off + Cc].float().permute(0, 3, 1, 2).contiguous()


CONV_CASES = [
    # N, H, W, Cin, Cout, R, stride, pad, dil
    (2, 17, 23, 32, 16, 1, 1, 0, 1),
    (2, 16, 24, 16, 96, 1, 1, 0, 1),
    (1, 33, 33, 320, 256, 3, 1, 6, 6),
    (2, 9, 12, 320, 256, 3, 1, 18, 18),
    (2, 20, 28, 304, 256, 3, 1, 1, 1),
    (2, 32, 48, 3, 32, 3, 2, 1, 1),
    (2, 33, 47, 3, 32, 3, 2, 1, 1),
    (2, 32, 48, 19, 64, 4, 2, 1, 1),
    (2, 17, 25, 64, 128, 4, 2, 1, 1),
    (2, 8, 12, 512, 1, 4, 2, 1, 1),
    (2, 16, 24, 256, 19, 1, 1, 0, 1),
    (1, 9, 12, 1024, 2, 3, 1, 1, 1),
    (3, 1, 1, 320, 256, 1, 1, 0, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(eng, cx, case):
    N, H, W, Cin, Cout, R, stride, pad, dil = case
    L = sub("_lib")
    g = torch.Generator(device="cuda").manual_seed(hash(case) & 0xffff)
    x = bf(torch.randn(N, Cin, H, W, device="cuda", generator=g))
    w = torch.randn(Cout, Cin, R, R, device="cuda", generator=g) * (2.0 / (Cin * R * R)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    wq = bf(w)
    xr = x.clone().requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    y_ref = F.conv2d(xr, wr, b, stride, pad, dil)
    OH, OW = y_ref.shape[2:]
    xa = nhwc_act(eng, x)
    wparam = torch.nn.Parameter(w.clone())
    out = cx.new(N, OH, OW, eng.round_up(Cout, 8), zero=True)
    out.C = Cout
    stats = cx.f64(2 * Cout)
    eng.conv_fwd(cx, xa, wparam, out, stride, pad, dil, bias=b, stats=stats)
    torch.cuda.synchronize()
    y = to_nchw(out)
    assert rel(y, y_ref.detach()) < 4e-3
    # the statistics are those of the STORED (bf16) outputs -- what the following BatchNorm normalises
    s_ref = torch.stack([y.double().sum((0, 2, 3)), (y.double() ** 2).sum((0, 2, 3))])
    assert rel(stats.view(2, Cout), s_ref) < 1e-4
    s_ref32 = torch.stack([y_ref.detach().double().sum((0, 2, 3)), (y_ref.detach().double() ** 2).sum((0, 2, 3))])
    assert rel(stats.view(2, Cout), s_ref32) < 5e-3
    # leaky epilogue
    out2 = cx.new(N, OH, OW, eng.round_up(Cout, 8), zero=True)
    out2.C = Cout
    eng.conv_fwd(cx, xa, wparam, out2, stride, pad, dil, bias=b, act=L.ACT_LEAKY, slope=0.2)
    assert rel(to_nchw(out2), F.leaky_relu(y_ref.detach(), 0.2)) < 4e-3
    # gradients
    dy = bf(torch.randn(N, Cout, OH, OW, device="cuda", generator=g))
    y_ref.backward(dy)
    dya = nhwc_act(eng, dy)
    wparam.grad = None
    eng.conv_wgrad(cx, xa, dya, wparam, stride, pad, dil)
    torch.cuda.synchronize()
    assert rel(wparam.grad, wr.grad) < 4e-3
    if Cin % 8 == 0 or True:
        dx = cx.new(N, H, W, eng.round_up(Cin, 8), zero=True)
        dx.C = Cin
        eng.conv_dgrad(cx, dya, wparam, dx, stride, pad, dil)
        torch.cuda.synchronize()
        assert rel(to_nchw(dx), xr.grad) < 4e-3
        # accumulate into an existing gradient
        eng.conv_dgrad(cx, dya, wparam, dx, stride, pad, dil, aux=dx, aux_mode=L.AUX_ADD)
        torch.cuda.synchronize()
        assert rel(to_nchw(dx), 2 * xr.grad) < 6e-3
    bparam = torch.nn.Parameter(b.clone())
    eng.bias_grad(cx, dya, bparam)
    assert rel(bparam.grad, dy.sum((0, 2, 3))) < 1e-4


def test_conv_into_concat_slice_and_leaky_mask(eng, cx):
    L = sub("_lib")
    g = torch.Generator(device="cuda").manual_seed(5)
    x = bf(torch.randn(2, 64, 12, 10, device="cuda", generator=g))
    w = torch.nn.Parameter(torch.randn(48, 64, 1, 1, device="cuda", generator=g) * 0.1)
    cat = cx.new(2, 12, 10, 304, zero=True)
    eng.conv_fwd(cx, nhwc_act(eng, x), w, cat.slice(256, 48))
    ref = F.conv2d(x, bf(w.detach()))
    assert rel(cat.t[..., 256:304].float().permute(0, 3, 1, 2), ref) < 4e-3
    assert float(cat.t[..., :256].abs().sum()) == 0.0
    # leaky mask epilogue of the data gradient
    y_prev = bf(torch.randn(2, 64, 12, 10, device="cuda", generator=g))
    dy = bf(torch.randn(2, 48, 12, 10, device="cuda", generator=g))
    dx = cx.new(2, 12, 10, 64)
    eng.conv_dgrad(cx, nhwc_act(eng, dy), w, dx, aux=nhwc_act(eng, y_prev), aux_mode=L.AUX_LEAKY_MASK, slope=0.2)
    ref = F.conv_transpose2d(dy, bf(w.detach())) * torch.where(y_prev > 0, 1.0, 0.2)
    assert rel(to_nchw(dx), ref) < 4e-3



@pytest.mark.parametrize("shape", [(1, 256, 512, 19, 1024, 2048), (2, 33, 17, 19, 129, 65), (3, 8, 8, 5, 8, 8),
                                   (1, 5, 7, 32, 1, 1), (2, 16, 16, 8, 61, 64)])
def test_fused_upsample_argmax_confusion_bit_exact(eng, cx, shape):
    """s2r_upsample_argmax_confusion_nhwc (deeplab.py:31 + val_adapt.py:133 + metrics.py:34-43 in one launch on the
    low-resolution NHWC bf16 logits) against the two launches it replaces -- s2r_upsample_bilinear_nhwc_to_nchw, then
    s2r_argmax_confusion_nchw on the fp32 logits -- and against numpy's argmax / the oracle's confusion matrix on those
    logits: identical counts, including ties (bf16 inputs make exact ties common), NaN and ignored / out-of-range labels."""
    L = sub("_lib")
    N, Hi, Wi, Cc, Ho, Wo = shape
    g = torch.Generator(device="cuda").manual_seed(Hi * Wi + Cc)
    x = bf(torch.randn(N, Cc, Hi, Wi, device="cuda", generator=g))
    x[:, :, ::3, ::2] = bf(torch.round(x[:, :, ::3, ::2]))          # many exact ties between classes
    if Hi > 4:
        x[0, Cc // 2, 2, 3] = float("nan")
    xa = nhwc_act(eng, x)
    nc = max(Cc, 2)
    gt = torch.randint(-1, nc + 2, (N, Ho, Wo), device="cuda", generator=g).float()
    gt[torch.rand(N, Ho, Wo, device="cuda", generator=g) < 0.1] = 255
    logits = torch.empty((N, Cc, Ho, Wo), dtype=torch.float32, device="cuda")
    L.call("s2r_upsample_bilinear_nhwc_to_nchw", xa.vp(), xa.pitch, N, Hi, Wi, Cc, C.c_void_p(logits.data_ptr()), Ho, Wo, cx.stream)
    two = torch.zeros((nc, nc), dtype=torch.int64, device="cuda")
    L.call("s2r_argmax_confusion_nchw", C.c_void_p(logits.data_ptr()), C.c_void_p(gt.data_ptr()), N, Cc, Ho * Wo, nc,
           C.c_void_p(two.data_ptr()), None, cx.stream)
    one = torch.zeros((nc, nc), dtype=torch.int64, device="cuda")
    L.call("s2r_upsample_argmax_confusion_nhwc", xa.vp(), xa.pitch, N, Hi, Wi, Cc, C.c_void_p(gt.data_ptr()), Ho, Wo, nc,
           C.c_void_p(one.data_ptr()), cx.stream)
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    pred = np.argmax(logits.cpu().numpy(), axis=1)
    want = O.confusion_matrix(gt.cpu().numpy(), pred, nc)
    assert np.array_equal(one.cpu().numpy(), want)
    assert int(one.sum()) == int(((gt >= 0) & (gt < nc)).sum())



def test_prepack_weights_multi_equals_single_pack(eng, cx):
    """prepack_weights (ONE launch re-packing every stale bf16 filter copy; several-tap filters through the
    tap-loop jobs of s2r_pack_weights_multi) against s2r_pack_weight filter by filter: bit-identical buffers for
    3x3 / 4x4 / 1x1 filters, ragged channel counts, forward and data-gradient orientation."""
    L = sub("_lib")
    g = torch.Generator(device="cuda").manual_seed(12)
    shapes = [(24, 19, 3, 3), (320, 257, 3, 3), (64, 24, 4, 4), (96, 16, 1, 1), (33, 70, 3, 3)]
    ws = [torch.nn.Parameter(torch.randn(*sh, device="cuda", generator=g)) for sh in shapes]
    for w in ws:
        for mode in (0, 1):
            eng.packed_weight(cx, w, mode)
    with torch.no_grad():
        for w in ws:
            w.mul_(1.5).add_(0.25)              # bumps the version: every copy is stale
    n = eng.prepack_weights(cx.stream)
    assert n == 2 * len(ws)
    torch.cuda.synchronize()
    for w in ws:
        for mode in (0, 1):
            buf, A_pad, B_pad = eng.packed_weight(cx, w, mode)      # cache hit: the multi launch's result
            Cout, Cin, RS, ns, A2, B2 = eng._pack_dims(w, mode)
            ref = torch.empty((ns, A2, B2), dtype=torch.bfloat16, device="cuda")
            L.call("s2r_pack_weight", C.c_void_p(w.data_ptr()), Cout, Cin, w.shape[2], w.shape[3], mode,
                   C.c_void_p(ref.data_ptr()), A2, B2, cx.stream)
            torch.cuda.synchronize()
            assert (A_pad, B_pad) == (A2, B2) and torch.equal(buf.view(torch.int16), ref.view(torch.int16)), (tuple(w.shape), mode)
            want = w.detach().permute(2, 3, 1, 0) if mode else w.detach().permute(2, 3, 0, 1)
            want = want.reshape(RS, *want.shape[2:]).to(torch.bfloat16)
            assert torch.equal(buf[:, :want.shape[1], :want.shape[2]], want)
            assert float(buf[:, want.shape[1]:, :].float().abs().sum()) == 0 and float(buf[:, :, want.shape[2]:].float().abs().sum()) == 0


DW_CASES = [(2, 18, 26, 32, 1, 1, False), (2, 17, 25, 96, 2, 1, True), (2, 16, 24, 144, 1, 1, True),
            (1, 9, 13, 960, 1, 2, True), (2, 12, 12, 192, 2, 1, True), (1, 10, 10, 384, 1, 1, True),
            # streaming stride-1 kernels: several column tiles / row segments, 32-, 48- and 16-channel chunks
            (2, 70, 75, 64, 1, 1, True), (1, 40, 37, 144, 1, 1, True), (1, 33, 65, 48, 1, 1, False),
            (1, 21, 150, 16, 1, 1, True), (2, 8, 31, 576, 1, 1, True), (1, 1, 1, 32, 1, 1, True),
            # dilation = parity planes of the stride-1 kernels
            (2, 16, 24, 64, 1, 2, True), (2, 11, 10, 32, 1, 2, False), (1, 13, 17, 48, 1, 3, True),
            # streaming stride-2 kernels (parity planes of the input), even / odd sizes, several tiles and segments
            (2, 70, 75, 64, 2, 1, True), (1, 41, 38, 144, 2, 1, True), (1, 34, 130, 16, 2, 1, True),
            (2, 2, 2, 32, 2, 1, True), (1, 3, 5, 96, 2, 1, True)]


@pytest.mark.parametrize("case", DW_CASES)
def test_dwconv_fused_prologue_halo(eng, cx, case):
    """dw3x3 on pad(relu6(bn(z))) where the pad value is relu6(shift) (halo_const) or 0."""
    N, H, W, Cc, stride, dil, halo = case
    L = sub("_lib")
    g = torch.Generator(device="cuda").manual_seed(Cc + stride)
    z = bf(torch.randn(N, Cc, H, W, device="cuda", generator=g) * 2)
    sc = torch.rand(Cc, device="cuda", generator=g) + 0.5
    sh = torch.randn(Cc, device="cuda", generator=g)
    mean = torch.randn(Cc, device="cuda", generator=g) * 0.1
    invstd = torch.rand(Cc, device="cuda", generator=g) + 0.5
    w = torch.randn(Cc, 1, 3, 3, device="cuda", generator=g) * 0.3
    pad = dil
    zr = z.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    if halo:
        # the reference computes bn+relu6 on the zero-padded tensor: border pre-activation = shift
        zp = F.pad(zr, (pad,) * 4)
        a = F.relu6(zp * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    else:
        a = F.pad(F.relu6(zr * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)), (pad,) * 4)
    y_ref = F.conv2d(a, wr, None, stride, 0, dil, Cc)
    st = eng.BNState(torch.cat([sc, sh]).contiguous(), torch.cat([mean, invstd]).contiguous(), 1.0, False)
    za = nhwc_act(eng, z)
    stats = cx.f64(2 * Cc)
    y = eng.dw_fwd(cx, za, st, L.ACT_RELU6, halo, w, stride, dil, pad, stats)
    torch.cuda.synchronize()
    assert rel(to_nchw(y), y_ref.detach()) < 4e-3
    s_ref = torch.stack([y_ref.detach().double().sum((0, 2, 3)), (y_ref.detach().double() ** 2).sum((0, 2, 3))])
    assert rel(stats.view(2, Cc), s_ref) < 1e-3
    dy = bf(torch.randn(*y_ref.shape, device="cuda", generator=g))
    if halo:
        zp.retain_grad()
    y_ref.backward(dy)
    dya = nhwc_act(eng, dy)
    dw = torch.zeros_like(w)
    L.call("s2r_dwconv3x3_wgrad", za.vp(), C.c_void_p(st.ss.data_ptr()), L.ACT_RELU6, 1 if halo else 0, dya.vp(),
           C.c_void_p(dw.data_ptr()), N, H, W, Cc, stride, dil, pad, cx.stream)
    torch.cuda.synchronize()
    assert rel(dw, wr.grad) < 4e-3
    ext = pad if halo else 0
    gbuf = cx.new(N, H + 2 * ext, W + 2 * ext, Cc)
    bsums = cx.f64(2 * Cc)
    L.call("s2r_dwconv3x3_dgrad", dya.vp(), C.c_void_p(w.data_ptr()), za.vp(), C.c_void_p(st.ss.data_ptr()),
           C.c_void_p(st.mi.data_ptr()), L.ACT_RELU6, ext, gbuf.vp(), C.c_void_p(bsums.data_ptr()), N, H, W, Cc, stride,
           dil, pad, cx.stream)
    torch.cuda.synchronize()
    # g = d loss / d (bn output), i.e. grad wrt pre-activation; grad wrt z(p) = g * scale
    g_ref = (zp.grad if halo else zr.grad) / sc.view(1, -1, 1, 1)
    assert rel(to_nchw(gbuf), g_ref) < 5e-3
    zfull = F.pad(z, (ext,) * 4) if halo else z
    xhat = (zfull - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    b_ref = torch.stack([g_ref.double().sum((0, 2, 3)), (g_ref.double() * xhat.double()).sum((0, 2, 3))])
    assert rel(bsums.view(2, Cc), b_ref) < 2e-3
    # both gradients in one pass (the entry point the engine uses)
    dw2 = torch.zeros_like(w)
    gbuf2 = cx.new(N, H + 2 * ext, W + 2 * ext, Cc)
    bsums2 = cx.f64(2 * Cc)
    L.call("s2r_dwconv3x3_bwd", dya.vp(), C.c_void_p(w.data_ptr()), za.vp(), C.c_void_p(st.ss.data_ptr()),
           C.c_void_p(st.mi.data_ptr()), L.ACT_RELU6, 1 if halo else 0, 0, gbuf2.vp(), C.c_void_p(bsums2.data_ptr()),
           C.c_void_p(dw2.data_ptr()), N, H, W, Cc, stride, dil, pad, cx.stream)
    torch.cuda.synchronize()
    assert rel(dw2, wr.grad) < 4e-3
    assert rel(to_nchw(gbuf2), g_ref) < 5e-3
    assert rel(bsums2.view(2, Cc), b_ref) < 2e-3
    # ... with g in the unextended layout (border positions only in the sums)
    gbuf3 = cx.new(N, H, W, Cc)
    bsums3 = cx.f64(2 * Cc)
    dw3 = torch.zeros_like(w)
    L.call("s2r_dwconv3x3_bwd", dya.vp(), C.c_void_p(w.data_ptr()), za.vp(), C.c_void_p(st.ss.data_ptr()),
           C.c_void_p(st.mi.data_ptr()), L.ACT_RELU6, 1 if halo else 0, 1, gbuf3.vp(), C.c_void_p(bsums3.data_ptr()),
           C.c_void_p(dw3.data_ptr()), N, H, W, Cc, stride, dil, pad, cx.stream)
    torch.cuda.synchronize()
    g_int = g_ref[:, :, ext:ext + H, ext:ext + W] if ext else g_ref
    assert rel(to_nchw(gbuf3), g_int) < 5e-3
    assert rel(bsums3.view(2, Cc), b_ref) < 2e-3
    assert rel(dw3, wr.grad) < 4e-3


@pytest.mark.parametrize("Cc,P,clamp", [(32, 5000, 0), (96, 777, 1), (256, 64, 0), (1024, 200, 1)])
def test_batchnorm_forward_backward(eng, cx, Cc, P, clamp):
    L = sub("_lib")
    g = torch.Generator(device="cuda").manual_seed(Cc)
    N, H, W = 1, P, 1
    x = bf(torch.randn(N, Cc, H, W, device="cuda", generator=g) * 1.5 + 0.3)
    bn = torch.nn.BatchNorm2d(Cc).cuda()
    bn.weight.data = torch.rand(Cc, device="cuda", generator=g) + 0.5
    bn.bias.data = torch.randn(Cc, device="cuda", generator=g) * 0.5
    ref = torch.nn.BatchNorm2d(Cc).cuda()
    ref.load_state_dict(bn.state_dict())
    xr = x.clone().requires_grad_(True)
    res = bf(torch.randn(N, Cc, H, W, device="cuda", generator=g))
    y_ref = F.relu6(ref(xr)) + res
    xa = nhwc_act(eng, x)
    sums = cx.f64(2 * Cc)
    L.call("s2r_channel_sums_bf16", xa.vp(), xa.P, Cc, xa.pitch, 0, C.c_void_p(sums.data_ptr()), cx.stream)
    if clamp:
        # the cross-rank branch (clamp(var, eps)^-1/2, batchnorm.py:125) with the all-reduce stubbed:
        # two "ranks" of P/2 elements whose sums are already added up
        bn._s2r_sync = True
        cx.world = 2
        cx.allreduce = lambda t: None
        st = eng.bn_finalize(cx, bn, sums, xa.P / 2)
    else:
        st = eng.bn_finalize(cx, bn, sums, xa.P)
    out = cx.new(N, H, W, Cc)
    eng.bn_apply(cx, xa, st, L.ACT_RELU6, out, residual=nhwc_act(eng, res))
    torch.cuda.synchronize()
    assert rel(to_nchw(out), y_ref.detach()) < 4e-3
    assert rel(bn.running_mean, ref.running_mean) < 1e-4 and rel(bn.running_var, ref.running_var) < 1e-4
    dy = bf(torch.randn(N, Cc, H, W, device="cuda", generator=g))
    y_ref.backward(dy)
    dx = cx.new(N, H, W, Cc)
    eng.bn_backward(cx, bn, nhwc_act(eng, dy), xa, st, L.ACT_RELU6, dx)
    torch.cuda.synchronize()
    assert rel(to_nchw(dx), xr.grad) < 6e-3
    assert rel(bn.weight.grad, ref.weight.grad) < 3e-3 and rel(bn.bias.grad, ref.bias.grad) < 3e-3
    # eval mode uses the running statistics
    ref.eval()
    st_e = eng.bn_eval(cx, bn)
    eng.bn_apply(cx, xa, st_e, L.ACT_NONE, out)
    assert rel(to_nchw(out), ref(x).detach()) < 4e-3


@pytest.mark.parametrize("kind,Cc,H,W", [("apply", 64, 32, 64), ("apply_res", 24, 33, 17), ("apply_relu", 256, 9, 13),
                                          ("dw1", 96, 40, 56), ("dw2", 144, 34, 30), ("dwd2", 64, 18, 22), ("dwgen", 40, 12, 20)])
def test_pending_batchnorm_finalised_by_its_consumer(eng, cx, kind, Cc, H, W):
    """A pending BatchNorm (sums left by the producer, no finalize launch) finalised in the prologue of the kernel that
    consumes the normalised tensor -- bn_apply / the streaming depthwise kernels, csrc/bn_tail.cuh -- against the
    separate bn_finalize launch on the same sums: same output, same published scale / shift / mean / inv-std, same
    running statistics (batchnorm.py:113-125 / F.batch_norm)."""
    L = sub("_lib")
    g = torch.Generator(device="cuda").manual_seed(Cc + H)
    N = 3
    bn_a, bn_b = torch.nn.BatchNorm2d(Cc).cuda(), torch.nn.BatchNorm2d(Cc).cuda()
    bn_a.weight.data = torch.rand(Cc, device="cuda", generator=g) + 0.5
    bn_a.bias.data = torch.randn(Cc, device="cuda", generator=g)
    bn_a.running_mean.data = torch.randn(Cc, device="cuda", generator=g)
    bn_b.load_state_dict(bn_a.state_dict())
    bn_a.train(); bn_b.train()
    x = nhwc_act(eng, bf(torch.randn(N, Cc, H, W, device="cuda", generator=g) * 1.3 + 0.4))
    res = nhwc_act(eng, bf(torch.randn(N, Cc, H, W, device="cuda", generator=g)))
    w = torch.randn(Cc, 1, 3, 3, device="cuda", generator=g) * 0.3

    def consume(st):
        if kind.startswith("apply"):
            out = cx.new(N, H, W, Cc)
            act = L.ACT_RELU if kind == "apply_relu" else L.ACT_RELU6 if kind == "apply" else L.ACT_NONE
            eng.bn_apply(cx, x, st, act, out, residual=res if kind == "apply_res" else None)
            return out
        stride, dil = (2, 1) if kind == "dw2" else (1, 2) if kind == "dwd2" else (1, 1)
        return eng.dw_fwd(cx, x, st, L.ACT_RELU6, True, w, stride, dil, dil, None)

    # A: pending, finalised by the consumer
    sums_a, finish = eng.bn_plan(cx, bn_a, x.P)
    L.call("s2r_channel_sums_bf16", x.vp(), x.P, Cc, x.pitch, 0, C.c_void_p(sums_a.data_ptr()), cx.stream)
    st_a = finish()
    assert st_a.pending is not None
    out_a = consume(st_a)
    assert st_a.pending is None
    # B: the separate launch
    sums_b = cx.f64(2 * Cc)
    L.call("s2r_channel_sums_bf16", x.vp(), x.P, Cc, x.pitch, 0, C.c_void_p(sums_b.data_ptr()), cx.stream)
    st_b = eng.bn_finalize(cx, bn_b, sums_b, x.P)
    out_b = consume(st_b)
    torch.cuda.synchronize()
    for a, b in ((st_a.ss, st_b.ss), (st_a.mi, st_b.mi), (bn_a.running_mean, bn_b.running_mean), (bn_a.running_var, bn_b.running_var)):
        assert rel(a, b) < 2e-6, (kind, rel(a, b))
    assert rel(out_a.t, out_b.t) < 1e-3
    assert st_a.count == st_b.count
    # a second consumer of the same state reads the published scale / shift
    out_c = consume(st_a)
    torch.cuda.synchronize()
    assert torch.equal(out_c.t, out_a.t)
    # any other consumer forces the finalisation with one small launch
    sums_d, finish_d = eng.bn_plan(cx, bn_a, x.P)
    L.call("s2r_channel_sums_bf16", x.vp(), x.P, Cc, x.pitch, 0, C.c_void_p(sums_d.data_ptr()), cx.stream)
    st_d = finish_d().ready(cx)
    torch.cuda.synchronize()
    assert st_d.pending is None and rel(st_d.ss, st_a.ss) < 2e-6


def test_bn_dropout_is_regenerated_in_backward(eng, cx):
    L = sub("_lib")
    Cc, P = 64, 4096
    x = bf(torch.randn(1, Cc, P, 1, device="cuda"))
    bn = torch.nn.BatchNorm2d(Cc).cuda()
    xa = nhwc_act(eng, x)
    st = eng.bn_eval(cx, bn)
    out = cx.new(1, P, 1, Cc)
    eng.bn_apply(cx, xa, st, L.ACT_NONE, out, drop_p=0.5, seed=1234)
    y = to_nchw(out)
    keep = (y != 0)
    frac = float(keep.float().mean())
    assert 0.47 < frac < 0.53
    assert rel(y[keep], (2 * x / (1 + 1e-5) ** 0.5)[keep]) < 4e-3
    ones = nhwc_act(eng, torch.ones_like(x))
    dx = cx.new(1, P, 1, Cc)
    eng.bn_backward(cx, bn, ones, xa, st, L.ACT_NONE, dx, drop_p=0.5, seed=1234)
    d = to_nchw(dx)
    assert bool(((d != 0) == keep).all())


@pytest.mark.parametrize("Hi,Wi,Ho,Wo", [(32, 64, 128, 256), (5, 7, 17, 25), (33, 33, 129, 129), (1, 1, 9, 12),
                                         (20, 300, 77, 1197), (24, 260, 94, 1040), (128, 256, 512, 1024)])
def test_bilinear_align_corners(eng, cx, Hi, Wi, Ho, Wo):
    L = sub("_lib")
    g = torch.Generator(device="cuda").manual_seed(Hi * Wo)
    x = bf(torch.randn(2, 24, Hi, Wi, device="cuda", generator=g))
    xr = x.clone().requires_grad_(True)
    y_ref = F.interpolate(xr, size=(Ho, Wo), mode='bilinear', align_corners=True)
    xa = nhwc_act(eng, x)
    out = cx.new(2, Ho, Wo, 40, zero=True)
    L.call("s2r_upsample_bilinear_nhwc", xa.vp(), 2, Hi, Wi, 24, out.slice(8, 24).vp(), Ho, Wo, 40, 0, cx.stream)
    assert rel(out.t[..., 8:32].float().permute(0, 3, 1, 2), y_ref.detach()) < 4e-3
    dy = bf(torch.randn(2, 24, Ho, Wo, device="cuda", generator=g))
    y_ref.backward(dy)
    dx = cx.new(2, Hi, Wi, 24)
    dya = nhwc_act(eng, dy)
    L.call("s2r_upsample_bilinear_nhwc_bwd", dya.vp(), 24, 0, 2, Hi, Wi, 24, Ho, Wo, dx.vp(), cx.stream)
    assert rel(to_nchw(dx), xr.grad) < 4e-3
    # NHWC bf16 (19 of 24 channels) -> NCHW fp32 and its backward
    y = torch.empty(2, 19, Ho, Wo, device="cuda")
    L.call("s2r_upsample_bilinear_nhwc_to_nchw", xa.vp(), 24, 2, Hi, Wi, 19, C.c_void_p(y.data_ptr()), Ho, Wo, cx.stream)
    assert rel(y, y_ref.detach()[:, :19]) < 1e-5
    gin = torch.randn(2, 19, Ho, Wo, device="cuda", generator=g)
    xr2 = x[:, :19].clone().requires_grad_(True)
    F.interpolate(xr2, size=(Ho, Wo), mode='bilinear', align_corners=True).backward(gin)
    dx2 = cx.new(2, Hi, Wi, 24)
    L.call("s2r_upsample_bilinear_nchw_bwd_to_nhwc", C.c_void_p(gin.data_ptr()), 2, 19, Ho, Wo, dx2.vp(), 24, Hi, Wi,
           cx.stream)
    assert rel(to_nchw(dx2, 19), xr2.grad) < 4e-3
    assert float(dx2.t[..., 19:].abs().sum()) == 0.0


@pytest.mark.parametrize("case", [(2, 3, 33, 40, 3, 3, 2, 1), (2, 19, 32, 64, 4, 4, 2, 1), (1, 19, 17, 23, 4, 4, 2, 1),
                                  (1, 3, 10, 300, 3, 3, 2, 1), (1, 5, 9, 132, 3, 3, 1, 1)])
def test_im2col_patch_gemm(eng, cx, case):
    """Patch matrix from NCHW fp32 (k = (c*R+ky)*S+kx, the OIHW order) and the conv / weight gradient through it."""
    N, Cc, H, W, R, S, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn(N, Cc, H, W, device="cuda", generator=g)
    P = eng.im2col(cx, eng.RawNCHW(x), R, S, stride, pad)
    OH, OW = eng.conv_out_hw(H, W, R, S, stride, pad, 1)
    ref = F.unfold(bf(x), (R, S), padding=pad, stride=stride).view(N, Cc * R * S, OH, OW).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert P.t.shape == (N, OH, OW, eng.round_up(Cc * R * S, 8))
    assert torch.equal(P.t[..., :Cc * R * S].float(), ref)
    assert float(P.t[..., Cc * R * S:].abs().sum()) == 0.0
    w = torch.nn.Parameter(torch.randn(16, Cc, R, S, device="cuda", generator=g) * 0.2)
    out = cx.new(N, OH, OW, 16)
    eng.conv_fwd(cx, P, eng.patch_weight(w), out)
    y_ref = F.conv2d(bf(x), bf(w.detach()), None, stride, pad)
    assert rel(to_nchw(out), y_ref) < 4e-3
    dy = bf(torch.randn(N, 16, OH, OW, device="cuda", generator=g))
    eng.conv_wgrad(cx, P, nhwc_act(eng, dy), eng.patch_weight(w), grad_param=w)
    torch.cuda.synchronize()
    wr = bf(w.detach()).requires_grad_(True)
    F.conv2d(bf(x), wr, None, stride, pad).backward(dy)
    assert rel(w.grad, wr.grad) < 4e-3


def test_avgpool_broadcast_layout(eng, cx):
    L = sub("_lib")
    x = bf(torch.randn(3, 320, 7, 9, device="cuda"))
    xa = nhwc_act(eng, x)
    pooled = cx.new(3, 1, 1, 320)
    L.call("s2r_avgpool_nhwc", xa.vp(), 3, 63, 320, 320, 0, 1.0 / 63, pooled.vp(), None, cx.stream)
    assert rel(to_nchw(pooled), x.mean((2, 3), keepdim=True)) < 4e-3
    dst = cx.new(3, 7, 9, 320, zero=True)
    L.call("s2r_broadcast_nhwc", pooled.vp(), 3, 63, 320, 2.0, 0, dst.vp(), 320, 0, cx.stream)
    L.call("s2r_broadcast_nhwc", pooled.vp(), 3, 63, 320, 1.0, 1, dst.vp(), 320, 0, cx.stream)
    assert rel(to_nchw(dst), 3 * x.mean((2, 3), keepdim=True).expand_as(x)) < 8e-3
    rt = sub("runtime")
    y = torch.randn(2, 19, 5, 6, device="cuda")
    a = rt.to_nhwc(cx, y)
    assert a.pitch == 24 and float(a.t[..., 19:].abs().sum()) == 0.0
    assert rel(rt.to_nchw(cx, a), bf(y)) == 0.0


def test_softmax_dim0_ce_bce(eng):
    fn = sub("functional")
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(4, 19, 13, 21, device="cuda", generator=g, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = fn.softmax_dim0(x), F.softmax(xr, dim=0)
    assert rel(y.detach(), yr.detach()) < 1e-6
    gy = torch.randn_like(y)
    y.backward(gy)
    yr.backward(gy)
    assert rel(x.grad, xr.grad) < 1e-5
    # cross entropy with ignore index, float targets and class weights
    for wgt in (None, torch.rand(19, device="cuda", generator=g) + 0.5):
        lab = torch.randint(0, 20, (4, 13, 21), device="cuda", generator=g).float()
        lab[lab == 19] = 255
        x1 = torch.randn(4, 19, 13, 21, device="cuda", generator=g, requires_grad=True)
        x2 = x1.detach().clone().requires_grad_(True)
        l1 = fn.cross_entropy(x1, lab, weight=wgt) * 1.7
        l2 = F.cross_entropy(x2, lab.long(), weight=wgt, ignore_index=255) * 1.7
        assert abs(l1.item() - l2.item()) < 1e-5 * abs(l2.item())
        l1.backward()
        l2.backward()
        assert rel(x1.grad, x2.grad) < 1e-5
    # odd spatial size (scalar path)
    lab = torch.randint(0, 19, (2, 5, 7), device="cuda", generator=g).float()
    x1 = torch.randn(2, 19, 5, 7, device="cuda", generator=g)
    assert abs(fn.cross_entropy(x1, lab).item() - F.cross_entropy(x1, lab.long()).item()) < 1e-5
    # BCE with logits vs constant targets
    for t in (0.0, 1.0):
        d1 = torch.randn(8, 1, 16, 32, device="cuda", generator=g, requires_grad=True)
        d2 = d1.detach().clone().requires_grad_(True)
        b1 = fn.bce_with_logits(d1, t)
        b2 = F.binary_cross_entropy_with_logits(d2, torch.full_like(d2, t))
        assert abs(b1.item() - b2.item()) < 1e-6
        b1.backward()
        b2.backward()
        assert rel(d1.grad, d2.grad) < 1e-5


def test_domain_loss_known_answer():
    """utils/loss.py:80-87 of the reference: loss = 2 ln(1+e^-1) = 0.626523, accuracy 1.0."""
    crit = sub("utils.loss").DomainLosses().build_loss()
    a, b = torch.ones(1, 1, 7, 7).cuda(), torch.zeros(1, 1, 7, 7).cuda()
    loss, acc = crit(torch.cat([a, b], 1), torch.cat([b, a], 1))
    assert abs(loss.item() - 0.626523) < 1e-5 and acc == 1.0
    with pytest.raises(AssertionError):
        crit(torch.zeros(1, 2, 3, 3).cuda(), torch.zeros(1, 2, 3, 4).cuda())


def test_evaluator_bit_exact():
    fix = golden('evaluator')
    Ev = sub("utils.metrics").Evaluator
    ev = Ev(19)
    ev.add_batch(fix['gt'], fix['pred'])
    ev.add_batch(fix['gt'][:, ::-1].copy(), fix['pred'])
    assert np.array_equal(ev.confusion_matrix, fix['cm'])
    miou, iou = ev.Mean_Intersection_over_Union()
    assert miou == float(fix['mIoU']) and np.array_equal(iou, fix['IoU'])
    assert ev.Pixel_Accuracy() == float(fix['PA']) and ev.Pixel_Accuracy_Class() == float(fix['mPA'])
    assert ev.Frequency_Weighted_Intersection_over_Union() == float(fix['fwIoU'])
    ev.reset()
    assert ev.confusion_matrix.sum() == 0
    # full-size image (1024x2048), labels as float with 255 = ignore, both entry points
    rng = np.random.RandomState(3)
    gt = rng.randint(0, 20, size=(1, 1024, 2048)).astype(np.float32)
    gt[gt == 19] = 255
    logits = torch.randn(1, 19, 1024, 2048, generator=torch.Generator().manual_seed(4))
    pred = np.argmax(logits.numpy(), axis=1)
    want = O.confusion_matrix(gt, pred, 19)
    ev.add_batch(gt, pred)
    assert np.array_equal(ev.confusion_matrix, want)
    ev.reset()
    ev.add_batch_logits(torch.from_numpy(gt).cuda(), logits.cuda())
    assert np.array_equal(ev.confusion_matrix, want)
    assert int(ev.confusion_matrix.sum()) == int(((gt >= 0) & (gt < 19)).sum())   # checksum property
    # edge cases: empty, all ignored, ragged size, int64 labels, out-of-range prediction
    ev.reset()
    ev.add_batch(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.int64))
    ev.add_batch(np.full((3, 5), 255, np.float32), np.zeros((3, 5), np.int64))
    assert ev.confusion_matrix.sum() == 0
    g2 = rng.randint(0, 19, size=(1, 7, 13)).astype(np.int64)
    p2 = rng.randint(0, 19, size=(1, 7, 13)).astype(np.int64)
    ev.add_batch(g2, p2)
    assert np.array_equal(ev.confusion_matrix, O.confusion_matrix(g2, p2, 19))
    with pytest.raises(AssertionError):
        ev.add_batch(np.zeros((2, 2)), np.zeros((2, 3)))
    ev.add_batch(np.zeros((1, 1), np.float32), np.full((1, 1), 19, np.int64))
    with pytest.raises(ValueError):
        ev.confusion_matrix


def test_fused_optimizers_match_torch():
    opt = sub("optim")
    g = torch.Generator(device="cuda").manual_seed(9)
    shapes = [(32, 3, 3, 3), (32,), (96, 16, 1, 1), (1000, 77), (5,)]
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    mine = opt.FusedSGD([{'params': ps[:2], 'lr': 0.01}, {'params': ps[2:], 'lr': 0.1}], lr=0.01, momentum=0.9,
                        weight_decay=5e-4)
    ref = torch.optim.SGD([{'params': qs[:2], 'lr': 0.01}, {'params': qs[2:], 'lr': 0.1}], lr=0.01, momentum=0.9,
                          weight_decay=5e-4)
    for it in range(4):
        mine.zero_grad()
        ref.zero_grad()
        for p, q in zip(ps, qs):
            gr = torch.randn(p.shape, device="cuda", generator=g)
            p.grad.add_(gr)
            q.grad = gr.clone()
        for o in (mine, ref):
            o.param_groups[0]['lr'] = 0.01 * (1 - it / 10)
            o.param_groups[1]['lr'] = 0.1 * (1 - it / 10)
        mine.step()
        ref.step()
    for p, q in zip(ps, qs):
        assert rel(p.detach(), q.detach()) < 1e-6
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    mine = opt.FusedAdam(ps, lr=1e-3, betas=(0.9, 0.99))
    ref = torch.optim.Adam(qs, lr=1e-3, betas=(0.9, 0.99))
    for it in range(5):
        mine.zero_grad()
        for p, q in zip(ps, qs):
            gr = torch.randn(p.shape, device="cuda", generator=g)
            p.grad.add_(gr)
            q.grad = gr.clone()
        mine.step()
        ref.step()
    for p, q in zip(ps, qs):
        assert rel(p.detach(), q.detach()) < 1e-5


def test_softmax0_to_padded_nhwc_and_back(eng):
    """s2r_softmax0_nchw_to_nhwc_pad / _bwd: interior = (batch softmax of) the input in bf16, border pixels and pad
    channels exactly zero; ragged strip widths, batch < 8 and > 8 (conversion only), backward against autograd."""
    import ctypes as C
    L = sub("_lib")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(21)
    for (B, Cc, H, W, sm) in ((8, 19, 6, 260, 1), (3, 19, 4, 130, 1), (1, 5, 2, 2, 1), (11, 19, 3, 40, 0), (2, 64, 2, 128, 0)):
        Cp = (Cc + 7) // 8 * 8
        x = torch.randn(B, Cc, H, W, device="cuda", generator=g)
        yp = torch.full((B, H + 2, W + 2, Cp), float('nan'), device="cuda", dtype=torch.bfloat16)
        L.call("s2r_softmax0_nchw_to_nhwc_pad", C.c_void_p(x.data_ptr()), B, Cc, H, W, sm, C.c_void_p(yp.data_ptr()), Cp, st)
        ref = (F.softmax(x, 0) if sm else x).permute(0, 2, 3, 1)
        got = yp.float()
        assert rel(got[:, 1:-1, 1:-1, :Cc], ref) < 3e-3          # bf16 rounding only
        assert float(got[:, 1:-1, 1:-1, Cc:].abs().sum()) == 0.0 if Cp > Cc else True
        border = got.clone()
        border[:, 1:-1, 1:-1, :] = 0
        assert not torch.isnan(got).any() and float(border.abs().sum()) == 0.0
        # backward: gradient given in the padded layout (border holds garbage that must be ignored)
        gp = torch.randn(B, H + 2, W + 2, Cp, device="cuda", generator=g).to(torch.bfloat16)
        dx = torch.empty_like(x)
        L.call("s2r_softmax0_nhwc_pad_bwd", C.c_void_p(x.data_ptr()) if sm else None, C.c_void_p(gp.data_ptr()), B, Cc, H, W, Cp,
               sm, C.c_void_p(dx.data_ptr()), st)
        gi = gp[:, 1:-1, 1:-1, :Cc].float().permute(0, 3, 1, 2)
        if sm:
            xr = x.clone().requires_grad_(True)
            F.softmax(xr, 0).backward(gi)
            assert rel(dx, xr.grad) < 1e-5
        else:
            assert rel(dx, gi) == 0.0
    # the batch softmax needs the whole batch in one CTA
    x = torch.randn(9, 3, 2, 2, device="cuda")
    yp = torch.empty(9, 4, 4, 8, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(NotImplementedError):
        L.call("s2r_softmax0_nchw_to_nhwc_pad", C.c_void_p(x.data_ptr()), 9, 3, 2, 2, 1, C.c_void_p(yp.data_ptr()), 8, st)


def test_rowtap_conv4x4_s2(eng):
    """Row-tap form of the 4x4 stride-2 pad-1 convolution on a zero-padded NHWC buffer (engine.rowtap_*): forward,
    weight gradient and data gradient against F.conv2d and autograd on bf16-representable operands."""
    import ctypes as C
    L = sub("_lib")
    cx = eng.Ctx(torch.device("cuda", 0), True)
    g = torch.Generator().manual_seed(22)
    for (N, Cin, Cout, H, W) in ((2, 19, 64, 16, 24), (1, 19, 64, 34, 258), (3, 8, 40, 6, 10)):
        Cp = (Cin + 7) // 8 * 8
        x = bf(torch.randn(N, Cin, H, W, generator=g))
        w = torch.nn.Parameter(bf(torch.randn(Cout, Cin, 4, 4, generator=g) * 0.1).cuda())
        b = torch.randn(Cout, generator=g)
        go = bf(torch.randn(N, Cout, H // 2, W // 2, generator=g))
        xr = x.clone().requires_grad_(True)
        wr = w.detach().cpu().clone().requires_grad_(True)
        yr = F.leaky_relu(F.conv2d(xr, wr, b, stride=2, padding=1), 0.2)
        pre_grad = go * torch.where(yr > 0, 1.0, 0.2)      # gradient w.r.t. the conv output
        F.conv2d(xr, wr, None, stride=2, padding=1).backward(pre_grad)
        xp = eng.PadAct(torch.empty(N, H + 2, W + 2, Cp, device="cuda", dtype=torch.bfloat16), H, W, Cin)
        L.call("s2r_softmax0_nchw_to_nhwc_pad", C.c_void_p(x.cuda().data_ptr()), N, Cin, H, W, 0, C.c_void_p(xp.ptr), Cp, cx.stream)
        out = cx.new(N, H // 2, W // 2, eng.round_up(Cout, 8))
        out.C = Cout
        eng.rowtap_fwd(cx, xp, w, out, bias=b.cuda(), act=L.ACT_LEAKY, slope=0.2)
        assert rel(out.t[..., :Cout].float().permute(0, 3, 1, 2), yr.detach()) < 4e-3
        dy = eng.Act(pre_grad.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda())
        dyf = dy.t.float().permute(0, 3, 1, 2).cpu()       # what the kernels see (bf16)
        xr2 = x.clone().requires_grad_(True)
        wr2 = w.detach().cpu().clone().requires_grad_(True)
        F.conv2d(xr2, wr2, None, stride=2, padding=1).backward(dyf)
        w.grad = None
        eng.rowtap_wgrad(cx, xp, dy, w)
        assert rel(w.grad, wr2.grad) < 2e-3
        dxp = eng.rowtap_dgrad(cx, dy, w, H, W)
        assert rel(dxp.t[:, 1:-1, 1:-1, :Cin].float().permute(0, 3, 1, 2), xr2.grad) < 4e-3


def test_device_input_stage_bit_exact_against_reference_fixture(eng):
    """dataloders.device_transforms (flip, PIL-exact bilinear / nearest resize, pad, crop, Normalize, ToTensor,
    encode_segmap) against the tensors the reference's TrainSet / ValSet produced for the same files and draws
    (tests/golden/input_stage.npz): bit-exact."""
    import numpy as np
    from conftest import golden
    dt = sub("dataloders.device_transforms")
    fix = golden("input_stage")
    for k in fix["cases"]:
        flip, short, crop, x1, y1 = (int(v) for v in fix[k + "_draw"])
        tr = dt.DeviceTrainTransform(base_size=1, crop_size=crop)
        src = torch.from_numpy(fix[k + "_src"]).cuda()[None].contiguous()
        tgt = torch.from_numpy(fix[k + "_tgt"]).cuda()[None].contiguous()
        lab = torch.from_numpy(fix[k + "_lab"]).cuda()[None].contiguous()
        out = tr(src, tgt, lab, draws=[(bool(flip), short, x1, y1)])
        assert np.array_equal(out['src_image'][0].cpu().numpy(), fix[k + "_out_src"]), k
        assert np.array_equal(out['tgt_image'][0].cpu().numpy(), fix[k + "_out_tgt"]), k
        assert np.array_equal(out['src_label'][0].cpu().numpy(), fix[k + "_out_lab"]), k
    # the same cases two samples at a time (different flip / scale / window per sample), through the batched launches
    # and through the per-sample launches
    names = [str(k) for k in fix["cases"]]
    for a, b in zip(names[0::2], names[1::2]):
        da, db = [int(v) for v in fix[a + "_draw"]], [int(v) for v in fix[b + "_draw"]]
        assert da[2] == db[2] and fix[a + "_src"].shape == fix[b + "_src"].shape
        tr = dt.DeviceTrainTransform(base_size=1, crop_size=da[2])
        src = torch.from_numpy(np.stack([fix[a + "_src"], fix[b + "_src"]])).cuda()
        tgt = torch.from_numpy(np.stack([fix[a + "_tgt"], fix[b + "_tgt"]])).cuda()
        lab = torch.from_numpy(np.stack([fix[a + "_lab"], fix[b + "_lab"]])).cuda()
        draws = [(bool(da[0]), da[1], da[3], da[4]), (bool(db[0]), db[1], db[3], db[4])]
        for batched in (True, False):
            out = tr(src, tgt, lab, draws=draws, batched=batched)
            for n, k in enumerate((a, b)):
                assert np.array_equal(out['src_image'][n].cpu().numpy(), fix[k + "_out_src"]), (k, batched)
                assert np.array_equal(out['tgt_image'][n].cpu().numpy(), fix[k + "_out_tgt"]), (k, batched)
                assert np.array_equal(out['src_label'][n].cpu().numpy(), fix[k + "_out_lab"]), (k, batched)
    s = int(fix["val_size"][0])
    va = dt.DeviceValTransform(s)
    img = torch.from_numpy(np.stack([fix["val_img"]] * 2)).cuda()
    lab = torch.from_numpy(np.stack([fix["val_lab"]] * 2)).cuda()
    out = va(img, lab)
    for n in range(2):
        assert np.array_equal(out['image'][n].cpu().numpy(), fix["val_out_img"])
        assert np.array_equal(out['label'][n].cpu().numpy(), fix["val_out_lab"])
    # the random stream: draws are made in the reference's order and reported
    import random
    random.seed(3)
    tr = dt.DeviceTrainTransform(base_size=40, crop_size=32, gaussian_blur=False)   # blur: tests/test_input_stage_blur.py
    x = torch.randint(0, 256, (2, 40, 64, 3), dtype=torch.uint8, device="cuda")
    lb = torch.randint(0, 34, (2, 40, 64), dtype=torch.uint8, device="cuda")
    o = tr(x, x, lb)
    assert o['src_image'].shape == (2, 3, 32, 32) and len(tr.last_draws) == 2 and torch.isfinite(o['src_image']).all()
    # size-independent property at full resolution: labels map through the table, valid trainIds or 255 only
    big = torch.randint(0, 256, (1, 1024, 2048), dtype=torch.uint8, device="cuda")
    img = torch.randint(0, 256, (1, 1024, 2048, 3), dtype=torch.uint8, device="cuda")
    tr = dt.DeviceTrainTransform(base_size=1024, crop_size=512)
    o = tr(img, img, big, draws=[(True, 1024, 700, 300)])     # no resize: pure flip + crop + table
    lut = torch.from_numpy(dt.segmap_lut()).cuda()
    want = lut[big[0].flip(1)[300:812, 700:1212].long()].float()
    assert torch.equal(o['src_label'][0], want)
    with pytest.raises(sub("_lib").S2RError):
        tr(img.cpu(), img.cpu(), big.cpu())


def test_prediction_export_and_validation_report(eng):
    """utils.report: fused argmax + labelId / palette tables + PIL NEAREST resize against the numpy restatement of
    test_adapt.py:118-157 (bit-exact, ties included) and against Pillow's own resize; report text of val_adapt.py:159-166."""
    import numpy as np
    from oracle import report as OR
    rep = sub("utils.report")
    g = torch.Generator().manual_seed(31)
    logits = torch.randn(2, 19, 64, 96, generator=g)
    logits[:, 3] = logits[:, 7]                       # ties: the lower class index must win
    logits[0, :, :8] = 0.0                            # all-equal pixels -> class 0
    ex = rep.PredictionExporter(out_size=(150, 80))
    ids, rgb = ex(logits.cuda())
    for n in range(2):
        want_ids, want_rgb = OR.imgsaver_arrays(logits[n].numpy(), 150, 80)
        assert np.array_equal(ids[n].cpu().numpy(), want_ids) and np.array_equal(rgb[n].cpu().numpy(), want_rgb)
    try:
        from PIL import Image
        im1 = np.uint8(np.argmax(logits[1].numpy(), 0))
        lab = np.zeros_like(im1)
        for c in range(19):
            lab[im1 == c] = OR.VALID_CLASSES[c]
        assert np.array_equal(np.array(Image.fromarray(lab, mode='L').resize((150, 80), Image.NEAREST)), ids[1].cpu().numpy())
    except ImportError:
        pass
    # full-size case through the default (1280, 640) output: every value is a valid labelId / palette colour
    big = torch.randn(1, 19, 512, 512, generator=g).cuda()
    ids, rgb = rep.PredictionExporter()(big)
    assert ids.shape == (1, 640, 1280) and set(ids.unique().tolist()) <= set(OR.VALID_CLASSES)
    pal = {tuple(p) for p in OR.PALETTE}
    assert {tuple(v) for v in rgb.reshape(-1, 3).unique(dim=0).tolist()} <= pal
    # report text
    ev = sub("utils.metrics").Evaluator(19)
    gt = torch.randint(0, 19, (2, 33, 47), generator=g).float()
    pr = torch.randint(0, 19, (2, 33, 47), generator=g)
    ev.add_batch(gt.cuda(), pr.cuda())
    text = rep.validation_report(ev, 3, 2, 1.23456)
    lines = text.split('\n')
    mIoU, IoU = ev.Mean_Intersection_over_Union()
    assert lines[0] == 'Validation:' and lines[1] == '[Epoch: 3, numImages:     2]' and lines[3] == 'Loss: 1.235'
    assert lines[2] == "Acc:{}, Acc_class:{}, mIoU:{}, fwIoU: {}".format(ev.Pixel_Accuracy(), ev.Pixel_Accuracy_Class(), mIoU,
                                                                        ev.Frequency_Weighted_Intersection_over_Union())
    assert lines[6] == '\troad: \t\t' + str(IoU[0]) and lines[7] == '\tsidewalk: \t' + str(IoU[1]) and len(lines) == 6 + 19 + 1


def test_focal_loss_against_reference_fixture(eng):
    """SegmentationLosses.build_loss('focal') (utils/loss.py:32-46: a scalar transform of the mean cross entropy) --
    value and gradient against the reference's own output (tests/golden/policy.npz), with and without class weights."""
    from conftest import golden
    fix = golden("policy")
    loss_mod = sub("utils.loss")
    lab = torch.from_numpy(fix["focal_label"]).cuda()
    for tag, wgt in (("", None), ("_w", torch.from_numpy(fix["focal_weight"]))):
        x = torch.from_numpy(fix["focal_logit"]).cuda().requires_grad_(True)
        loss = loss_mod.SegmentationLosses(weight=wgt).build_loss('focal')(x, lab)
        loss.backward()
        assert abs(loss.item() - float(fix["focal_loss" + tag])) <= 1e-5 * abs(float(fix["focal_loss" + tag]))
        assert rel(x.grad, torch.from_numpy(fix["focal_grad" + tag])) < 1e-5
    with pytest.raises(NotImplementedError):
        loss_mod.SegmentationLosses().build_loss('dice')
