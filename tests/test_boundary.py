"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, ctypes structs match the C layout, the tap tables the host builds for the
tap-GEMM kernels reproduce F.conv2d and its gradients (emulated in numpy, no GPU), and the
product path refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import PKG, ROOT, sub


def header_symbols():
    src = open(os.path.join(ROOT, "include", "s2r_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2r_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(built_lib, s), s
    assert built_lib.s2r_version() >= 100
    L = sub("_lib")
    assert set(L.PROTOTYPES) | set(L.PLAIN) == set(syms)


def test_struct_layout_matches_c(built_lib, tmp_path):
    """sizeof() of the ctypes mirrors == sizeof() in C (compiled from the header with gcc)."""
    L = sub("_lib")
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include "s2r_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(s2r_tap),'
                 ' sizeof(s2r_conv_args), sizeof(s2r_wgrad_args), sizeof(s2r_param_slot), sizeof(s2r_bn_tail));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(L.Tap), ctypes.sizeof(L.ConvArgs), ctypes.sizeof(L.WgradArgs),
                     ctypes.sizeof(L.ParamSlot), ctypes.sizeof(L.BnTail)]


def test_struct_size_is_checked(built_lib):
    L = sub("_lib")
    a = L.ConvArgs()
    a.struct_size = 4
    with pytest.raises(ValueError):
        L.call("s2r_conv_fwd", ctypes.byref(a), None)
    assert "struct size" in L.last_error()


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    L = sub("_lib")
    m = sub("modeling.discriminator").FCDiscriminator(19)
    with pytest.raises(L.S2RError):
        m(torch.zeros(1, 19, 32, 32))
    with pytest.raises(L.S2RError):
        sub("functional").softmax_dim0(torch.zeros(2, 3))
    with pytest.raises(L.S2RError):
        sub("utils.metrics").Evaluator(19).add_batch(np.zeros((2, 2), np.float32), np.zeros((2, 2), np.int64))


def test_reference_error_conventions():
    nn = torch.nn
    with pytest.raises(NotImplementedError):
        sub("modeling.assp").ASPP('mobilenet', 32, nn.BatchNorm2d)
    with pytest.raises(NotImplementedError):
        sub("modeling.decoder").Decoder(19, 'vgg', nn.BatchNorm2d)
    with pytest.raises(NotImplementedError):
        sub("modeling.domian").DomainClassifer('resnet', nn.BatchNorm2d)
    with pytest.raises(NotImplementedError):
        sub("utils.loss").SegmentationLosses().build_loss('dice')
    with pytest.raises(NotImplementedError):
        sub("modeling.backbone").build_backbone('resnet', 16, nn.BatchNorm2d)
    with pytest.raises(AssertionError):
        sub("modeling.backbone.mobilenet").InvertedResidual(8, 8, 3, 1, 6, nn.BatchNorm2d)


# ----------------------------------------------------------------------------- tap tables
class FakeAct:
    """Stand-in for engine.Act on a numpy buffer: ptr is a byte offset (2 bytes per element)."""

    def __init__(self, arr, base):
        self.arr = arr
        self.N, self.H, self.W, self.pitch = arr.shape
        self.C = self.pitch
        self.off = 0
        self.ptr = base


def read_view(mem, tap, n, y, x, Cn):
    if not (0 <= y < tap.H and 0 <= x < tap.W):
        return np.zeros(Cn, dtype=np.float64)
    e = (tap.base or 0) // 2 + n * tap.sn + y * tap.sh + x * tap.sw
    return mem[e:e + Cn]


@pytest.mark.parametrize("R,stride,pad,dil,H,W", [(3, 1, 1, 1, 7, 9), (3, 1, 6, 6, 9, 11), (3, 2, 1, 1, 9, 8),
                                                  (4, 2, 1, 1, 8, 10), (4, 2, 1, 1, 9, 11), (1, 1, 0, 1, 5, 4)])
def test_tap_tables_reproduce_conv2d_and_gradients(R, stride, pad, dil, H, W):
    eng = sub("engine")
    L = sub("_lib")
    rng = np.random.RandomState(R * 100 + stride * 10 + dil)
    N, Cin, Cout = 2, 8, 5
    x = rng.randn(N, H, W, Cin)
    w = rng.randn(Cout, Cin, R, R)
    OH, OW = eng.conv_out_hw(H, W, R, R, stride, pad, dil)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).clone().requires_grad_(True)
    wt = torch.from_numpy(w).clone().requires_grad_(True)
    yt = F.conv2d(xt, wt, None, stride, pad, dil)
    dy = rng.randn(*yt.shape)
    yt.backward(torch.from_numpy(dy))

    # forward: out[m][co] = sum_t sum_ci view_t[pix + d_t][ci] * w[co][ci][t]
    mem = x.reshape(-1).copy()
    xa = FakeAct(x, 0)
    taps = (L.Tap * 16)()
    nt = eng.fwd_taps(taps, xa, R, R, stride, pad, dil)
    assert nt == R * R
    out = np.zeros((N, OH, OW, Cout))
    for t in range(nt):
        kh, kw = divmod(taps[t].wslice, R)
        for n in range(N):
            for oh in range(OH):
                for ow in range(OW):
                    v = read_view(mem, taps[t], n, oh + taps[t].dh, ow + taps[t].dw, Cin)
                    out[n, oh, ow] += w[:, :, kh, kw] @ v
    assert np.allclose(out, yt.detach().permute(0, 2, 3, 1).numpy(), atol=1e-9)

    # weight gradient from the same taps
    dyn = dy.transpose(0, 2, 3, 1)
    dw = np.zeros_like(w)
    for t in range(nt):
        kh, kw = divmod(taps[t].wofs, R)
        for n in range(N):
            for oh in range(OH):
                for ow in range(OW):
                    v = read_view(mem, taps[t], n, oh + taps[t].dh, ow + taps[t].dw, Cin)
                    dw[:, :, kh, kw] += np.outer(dyn[n, oh, ow], v)
    assert np.allclose(dw, wt.grad.numpy(), atol=1e-8)

    # data gradient: one tap-GEMM per input parity class, recorded instead of launched
    calls = []
    real_call = L.call

    def fake_call(name, *args):
        a = args[0]._obj
        calls.append((a.ntaps, [(a.taps[i].dh, a.taps[i].dw, a.taps[i].wslice, a.taps[i].H, a.taps[i].W) for i in
                                range(a.ntaps)], a.OH, a.OW, a.out, a.on, a.oh, a.ow))
        return 0

    class FakeCx:
        stream = None

    dya = FakeAct(np.ascontiguousarray(dyn_pad(dyn)), 1 << 20)
    dxa = FakeAct(np.zeros((N, H, W, Cin)), 0)
    wparam = torch.zeros(Cout, Cin, R, R)
    eng_pack = eng.packed_weight
    eng.packed_weight = lambda cx, w_, tr: (torch.zeros(1), 16, 64)
    L.call = fake_call
    try:
        eng.conv_dgrad(FakeCx(), dya, wparam, dxa, stride, pad, dil)
    finally:
        L.call = real_call
        eng.packed_weight = eng_pack
    dx = np.zeros((N, H, W, Cin))
    dymem = dya.arr
    for ntaps, tl, OHc, OWc, outp, on, oh_s, ow_s in calls:
        for n in range(N):
            for i in range(OHc):
                for j in range(OWc):
                    acc = np.zeros(Cin)
                    for dh, dw_, ws, TH, TW in tl:
                        y, xx = i + dh, j + dw_
                        if 0 <= y < TH and 0 <= xx < TW:
                            kh, kw = divmod(ws, R)
                            acc += w[:, :, kh, kw].T @ dymem[n, y, xx, :Cout]
                    e = (outp or 0) // 2 + n * on + i * oh_s + j * ow_s
                    dx.reshape(-1)[e:e + Cin] += acc
    assert np.allclose(dx, xt.grad.permute(0, 2, 3, 1).numpy(), atol=1e-8)


def dyn_pad(dyn):
    """pad the channel axis of dy to a multiple of 8 (zeros) as the kernels require"""
    c = dyn.shape[-1]
    p = (8 - c % 8) % 8
    return np.pad(dyn, ((0, 0), (0, 0), (0, 0), (0, p)))


def _pack_numpy(w, mode):
    """numpy restatement of s2r_pack_weight's index maps for the row-tap modes (csrc/conv_mma.cu pack_src):
    mode 2: [kh][co][kw*Cp + c]; mode 3/4 (ph = mode - 3): [2*ta + tb][pw*Cp + c][co] with (kh, kw) = (ph+2ta, pw+2tb)."""
    Cout, Cin = w.shape[:2]
    Cp = (Cin + 7) // 8 * 8
    if mode == 2:
        out = np.zeros((4, Cout, 4 * Cp))
        for kh in range(4):
            for kw in range(4):
                out[kh, :, kw * Cp:kw * Cp + Cin] = w[:, :, kh, kw]
        return out
    ph = mode - 3
    out = np.zeros((4, 2 * Cp, Cout))
    for ta in range(2):
        for tb in range(2):
            for pw in range(2):
                out[2 * ta + tb, pw * Cp:pw * Cp + Cin, :] = w[:, :, ph + 2 * ta, pw + 2 * tb].T
    return out


@pytest.mark.parametrize("H,W,Cin,Cout", [(8, 12, 19, 6), (6, 6, 8, 5)])
def test_rowtap_views_and_pack_maps_reproduce_conv4x4_s2(H, W, Cin, Cout):
    """Host logic of the discriminator's row-tap convolution (engine.PadAct / rowtap_*): the four overlapping strided
    views over the zero-padded NHWC buffer, with the packed-filter index maps, reproduce F.conv2d(k=4, s=2, p=1), its
    weight gradient and its data gradient (numpy, no GPU)."""
    eng = sub("engine")
    L = sub("_lib")
    rng = np.random.RandomState(H * W + Cin)
    N = 2
    Cp = (Cin + 7) // 8 * 8
    x = rng.randn(N, Cin, H, W)
    w = rng.randn(Cout, Cin, 4, 4)
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    wt = torch.from_numpy(w).clone().requires_grad_(True)
    yt = F.conv2d(xt, wt, None, 2, 1)
    dy = rng.randn(*yt.shape)
    yt.backward(torch.from_numpy(dy))
    OH, OW = H // 2, W // 2
    xp = np.zeros((N, H + 2, W + 2, Cp))
    xp[:, 1:-1, 1:-1, :Cin] = x.transpose(0, 2, 3, 1)
    mem = xp.reshape(-1)

    class FakePad:
        ptr = 0

    fp = FakePad()
    fp.N, fp.H, fp.W, fp.C, fp.Cp = N, H, W, Cin, Cp
    taps = (L.Tap * 16)()
    assert eng._rowtap_views(taps, fp) == (OH, OW)
    # forward and weight gradient through the four views (K = 4*Cp contiguous channels per tap)
    wp = _pack_numpy(w, 2)
    out = np.zeros((N, OH, OW, Cout))
    G = np.zeros((Cout, 4, 4 * Cp))
    dyn = dy.transpose(0, 2, 3, 1)
    for t in range(4):
        assert taps[t].wslice == t and taps[t].wofs == t * 4 * Cp
        for n in range(N):
            for oh in range(OH):
                for ow in range(OW):
                    v = read_view(mem, taps[t], n, oh, ow, 4 * Cp)
                    out[n, oh, ow] += wp[t] @ v
                    G[:, t, :] += np.outer(dyn[n, oh, ow], v)
    assert np.allclose(out, yt.detach().permute(0, 2, 3, 1).numpy(), atol=1e-9)
    dw = np.zeros_like(w)                      # s2r_rowtap_wgrad_scatter: w.grad[co][c][kh][kw] += G[co][kh][kw*Cp + c]
    for kh in range(4):
        for kw in range(4):
            dw[:, :, kh, kw] = G[:, kh, kw * Cp:kw * Cp + Cin]
    assert np.allclose(dw, wt.grad.numpy(), atol=1e-8)
    # data gradient: one launch per padded-row parity, output "pixel" = two adjacent padded pixels
    calls = []
    real_call, real_pack = L.call, eng.packed_weight

    def fake_call(name, *args):
        a = args[0]._obj
        calls.append(([(a.taps[i].dh, a.taps[i].dw, a.taps[i].wslice) for i in range(a.ntaps)], a.OH, a.OW, a.out, a.on, a.oh,
                      a.ow, a.Cout))
        return 0

    class FakeCx:
        stream = None
        device = torch.device("cpu")

    modes = []
    eng.packed_weight = lambda cx, w_, mode: (modes.append(mode) or torch.zeros(1), 48, 64)
    L.call = fake_call
    real_empty = torch.empty
    try:
        torch.empty = lambda *a, **k: real_empty(*a, **{kk: vv for kk, vv in k.items() if kk != 'device'})
        dya = FakeAct(np.ascontiguousarray(dyn_pad(dyn)), 1 << 20)
        dxp = eng.rowtap_dgrad(FakeCx(), dya, torch.zeros(Cout, Cin, 4, 4), H, W)
    finally:
        torch.empty = real_empty
        L.call, eng.packed_weight = real_call, real_pack
    assert modes == [3, 4] and len(calls) == 2
    base = dxp.ptr
    dxmem = np.zeros((N, H + 2, W + 2, Cp))
    for ph, (tl, OHc, OWc, outp, on, oh_s, ow_s, nout) in enumerate(calls):
        wd = _pack_numpy(w, 3 + ph)
        assert nout == 2 * Cp and (OHc, OWc) == ((H + 2) // 2, (W + 2) // 2)
        for n in range(N):
            for i in range(OHc):
                for j in range(OWc):
                    acc = np.zeros(2 * Cp)
                    for dh, dw_, ws in tl:
                        y, xx = i + dh, j + dw_
                        if 0 <= y < OH and 0 <= xx < OW:
                            acc += wd[ws] @ dyn[n, y, xx]
                    e = (outp - base) // 2 + n * on + i * oh_s + j * ow_s
                    dxmem.reshape(-1)[e:e + 2 * Cp] += acc
    assert np.allclose(dxmem[:, 1:-1, 1:-1, :Cin], xt.grad.permute(0, 2, 3, 1).numpy(), atol=1e-8)


def test_blur_kernel_text_compiled_for_the_host(tmp_path):
    """The RandomGaussianBlur kernels (csrc/input_stage.cu: blur_tap, blur_three_passes, blur_rows_kernel,
    blur_cols_kernel) taken as TEXT out of the .cu, stripped of their CUDA qualifiers, with the grid-stride loop reduced
    to a plain loop, compiled with g++ and run on random windows / mirrors / paddings / radii against the oracle's
    restatement of Pillow's GaussianBlur: the arithmetic and the addressing of the kernels without a GPU."""
    import subprocess
    from oracle import input_stage as OI
    src = open(os.path.join(ROOT, PKG, "csrc", "input_stage.cu")).read()
    core = src[src.index("__device__ __forceinline__ uint32_t blur_tap"):src.index("// grid = (blocks over H*W*3 bytes of the crop, njobs)")]
    kernels = src[src.index("__global__ void __launch_bounds__(kThreads)\nblur_rows_kernel"):
                  src.index('}  // namespace\n\nextern "C" int s2r_gaussian_blur3_u8_multi')]
    core = core.replace("__device__ __forceinline__", "static inline")
    kernels = (kernels.replace("__global__ void __launch_bounds__(kThreads)\n", "void ").replace("__restrict__", "")
               .replace("jobs[blockIdx.y]", "jobs[block_y]").replace("blockIdx.x * kThreads + threadIdx.x", "0")
               .replace("gridDim.x * kThreads", "1"))
    assert "blockIdx" not in kernels and "threadIdx" not in kernels
    header = open(os.path.join(ROOT, "include", "s2r_b200.h")).read()
    job = header[header.index("typedef struct s2r_blur_job {"):header.index("} s2r_blur_job;") + len("} s2r_blur_job;")]
    cpp = tmp_path / "blur_host.cpp"
    cpp.write_text("#include <cstdint>\nconstexpr int kThreads = 256;\nstatic int block_y = 0;\n" + job + "\n" + core + kernels +
                   'extern "C" void run(const s2r_blur_job* jobs, int n, int H, int W) {\n'
                   "  for (block_y = 0; block_y < n; ++block_y) blur_rows_kernel(jobs, H, W);\n"
                   "  for (block_y = 0; block_y < n; ++block_y) blur_cols_kernel(jobs, H, W);\n}\n")
    so = tmp_path / "blur_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(cpp)])
    lib = ctypes.CDLL(str(so))
    L = sub("_lib")
    dt = sub("dataloders.device_transforms")
    rng = np.random.default_rng(1)
    for t in range(120):
        Hs, Ws, cs = int(rng.integers(1, 50)), int(rng.integers(1, 50)), int(rng.integers(2, 40))
        flip = int(rng.integers(0, 2))
        x1 = int(rng.integers(0, max(1, max(Ws, cs) - cs + 1)))
        y1 = int(rng.integers(0, max(1, max(Hs, cs) - cs + 1)))
        img = rng.integers(0, 256, (Hs, Ws, 3), dtype=np.uint8)
        r = float(rng.random()) * (1.0 if t % 4 else 1.41)
        ww, fw = dt._gaussian_blur_weights(r)
        tmp, out = np.zeros((cs, cs, 3), np.uint8), np.zeros((cs, cs, 3), np.uint8)
        job_s = L.BlurJob(img.ctypes.data, tmp.ctypes.data, out.ctypes.data, Hs, Ws, flip, x1, y1, ww, fw, 0)
        lib.run(ctypes.byref(job_s), 1, cs, cs)
        a = img[:, ::-1] if flip else img
        pad = np.zeros((max(Hs, y1 + cs), max(Ws, x1 + cs), 3), np.uint8)
        pad[:Hs, :Ws] = a
        assert np.array_equal(out, OI.gaussian_blur(pad[y1:y1 + cs, x1:x1 + cs], r)), (t, Hs, Ws, cs, flip, x1, y1, r)
