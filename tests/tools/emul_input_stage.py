"""CPU emulation (numpy / pure Python) of the batched input-stage kernels driven by the REAL host planning
(DeviceTrainTransform._run_batched) and compared with tests/golden/input_stage.npz: checks the window / job-table logic
without a GPU.  Every emulated entry point follows its kernel in csrc/input_stage.cu statement by statement, reading and
writing host memory through the raw pointers of the job tables.

    python tests/tools/emul_input_stage.py      (about half a minute; ends with "ALL OK")
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
dt = importlib.import_module('synthetic-to-real-semantic-segmentation_b200.dataloders.device_transforms')
L = importlib.import_module('synthetic-to-real-semantic-segmentation_b200._lib')
from oracle import input_stage as OI  # noqa: E402

PB = 22                       # PRECISION_BITS of libImaging/Resample.c
_CTYPE = {np.uint8: C.c_uint8, np.int32: C.c_int32, np.float32: C.c_float}


def arr(ptr, n, dtype=np.uint8):
    """n elements of host memory at address `ptr` as a writable numpy view."""
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(_CTYPE[dtype])), shape=(n,))


def clip8(v):
    return np.uint8(min(255, max(0, v >> PB)))


def resize_bilinear_multi(tab, n, max_elems, stream):
    """resize_multi_kernel"""
    for j in (L.ResizeJob * n).from_address(tab):
        bnd = arr(j.bounds, 1 << 20, np.int32)
        if j.axis == 1:
            for line in range(j.lines):
                for xo in range(j.on):
                    xx = j.o0 + xo
                    xmin, cnt = int(bnd[2 * xx]), int(bnd[2 * xx + 1])
                    k = arr(j.kk + 4 * xx * j.ksize, j.ksize, np.int32)
                    for c in range(j.C):
                        acc = 1 << (PB - 1)
                        for x in range(cnt):
                            sx = (j.W - 1 - (xmin + x)) if j.flip else xmin + x
                            acc += int(arr(j.inp + line * j.in_pitch + sx * j.C + c, 1)[0]) * int(k[x])
                        arr(j.out + (line * j.on + xo) * j.C + c, 1)[0] = clip8(acc)
        else:
            for yo in range(j.on):
                yy = j.o0 + yo
                ymin, cnt = int(bnd[2 * yy]), int(bnd[2 * yy + 1])
                k = arr(j.kk + 4 * yy * j.ksize, j.ksize, np.int32)
                for x in range(j.lines):
                    acc = 1 << (PB - 1)
                    for y in range(cnt):
                        acc += int(arr(j.inp + (ymin - j.base + y) * j.in_pitch + x, 1)[0]) * int(k[y])
                    arr(j.out + yo * j.lines + x, 1)[0] = clip8(acc)


def resize_nearest_multi(tab, n, max_elems, stream):
    """nearest_multi_kernel"""
    for j in (L.NearestJob * n).from_address(tab):
        xt, yt = arr(j.xtab, 1 << 16, np.int32), arr(j.ytab, 1 << 16, np.int32)
        for y in range(j.OH):
            for x in range(j.OW):
                sx, sy = int(xt[j.x0 + x]), int(yt[j.y0 + y])
                v = 0
                if sx >= 0 and sy >= 0:
                    v = arr(j.inp + sy * j.W + ((j.W - 1 - sx) if j.flip else sx), 1)[0]
                arr(j.out + y * j.OW + x, 1)[0] = v


def input_stage_multi(tab, n, mean, std, lut, fill, H, W, stream):
    """input_stage_multi_kernel (the normalisation table is the oracle's Normalize applied to every byte value)"""
    lutv = arr(lut, 256)
    norm = OI.normalize_to_tensor(np.arange(256, dtype=np.uint8).reshape(256, 1, 1).repeat(3, 2).reshape(256, 1, 3))   # [3][256][1]
    for j in (L.StageJob * n).from_address(tab):
        for y in range(H):
            for x in range(W):
                sy, sx0 = j.y1 + y, j.x1 + x
                inside = sy < j.Hs and sx0 < j.Ws
                sx = (j.Ws - 1 - sx0) if j.flip else sx0
                if j.img:
                    px = [int(v) for v in arr(j.img + (sy * j.Ws + sx) * 3, 3)] if inside else [0, 0, 0]
                    o = arr(j.out_img, 3 * H * W, np.float32)
                    for c in range(3):
                        o[c * H * W + y * W + x] = norm[c, px[c], 0]
                if j.label:
                    v = float(lutv[arr(j.label + sy * j.Ws + sx, 1)[0]]) if inside else float(fill)
                    arr(j.out_label, H * W, np.float32)[y * W + x] = v


def gaussian_blur3_multi(tab, n, H, W, stream):
    """blur_rows_kernel, then blur_cols_kernel"""
    def tap(l, c, r, ww, fw):
        return ((c * ww + (l + r) * fw + (1 << 23)) & 0xffffffff) >> 24

    def three_passes(p0, x, last, ww, fw):
        lo = lambda j: 0 if j < 0 else j                # noqa: E731
        hi = lambda j: last if j > last else j          # noqa: E731
        p1 = lambda j: tap(p0(lo(j - 1)), p0(j), p0(hi(j + 1)), ww, fw)   # noqa: E731
        p2 = lambda j: tap(p1(lo(j - 1)), p1(j), p1(hi(j + 1)), ww, fw)   # noqa: E731
        return tap(p2(lo(x - 1)), p2(x), p2(hi(x + 1)), ww, fw)

    for j in (L.BlurJob * n).from_address(tab):
        tmp, out = arr(j.tmp, H * W * 3), arr(j.out, H * W * 3)
        for i in range(H * W * 3):
            c, x, y = i % 3, (i // 3) % W, i // (3 * W)
            sy = j.y1 + y

            def px(xx):
                sx0 = j.x1 + xx
                if not (sy < j.Hs) or sx0 >= j.Ws:
                    return 0
                return int(arr(j.img + sy * j.Ws * 3 + c + ((j.Ws - 1 - sx0) if j.flip else sx0) * 3, 1)[0])

            tmp[i] = three_passes(px, x, W - 1, j.ww, j.fw)
        pitch = W * 3
        for i in range(H * pitch):
            xb, y = i % pitch, i // pitch
            out[i] = three_passes(lambda yy: int(tmp[xb + yy * pitch]), y, H - 1, j.ww, j.fw)


KERNELS = {"s2r_resize_bilinear_u8_multi": resize_bilinear_multi, "s2r_resize_nearest_u8_multi": resize_nearest_multi,
           "s2r_input_stage_u8_multi": input_stage_multi, "s2r_gaussian_blur3_u8_multi": gaussian_blur3_multi}


def emu(name, *args):
    KERNELS[name](*args)
    return 0


L.call = emu
dt.L.call = emu


def run_pair(fix, a, b):
    """Samples a and b of the fixture as one batch of two through the real _run_batched; RandomGaussianBlur fires on a
    sample iff the fixture holds radii for it.  Returns True when all six tensors equal the reference's."""
    da, db = [int(v) for v in fix[a + '_draw']], [int(v) for v in fix[b + '_draw']]
    tr = dt.DeviceTrainTransform(1, da[2])
    src = torch.from_numpy(np.stack([fix[a + '_src'], fix[b + '_src']]))
    tgt = torch.from_numpy(np.stack([fix[a + '_tgt'], fix[b + '_tgt']]))
    lab = torch.from_numpy(np.stack([fix[a + '_lab'], fix[b + '_lab']]))
    N, H, W, _ = src.shape
    cs = da[2]
    out = {'src_image': torch.zeros(N, 3, cs, cs), 'tgt_image': torch.zeros(N, 3, cs, cs), 'src_label': torch.zeros(N, cs, cs)}
    plan, blur = [], []
    for d, k in ((da, a), (db, b)):
        ow, oh = dt._scale_size(W, H, d[1])
        plan.append((bool(d[0]), ow, oh, d[3], d[4]))
        if (k + '_radii') in fix.files:
            r = fix[k + '_radii']
            blur.append({'src_image': dt._gaussian_blur_weights(float(r[0])), 'tgt_image': dt._gaussian_blur_weights(float(r[1]))})
        else:
            blur.append(None)
    keep = tr._run_batched(src, tgt, lab, plan, out, None, blur)     # noqa: F841  (scratch buffers stay alive)
    ok = True
    for n, k in enumerate((a, b)):
        e = [np.array_equal(out['src_image'][n].numpy(), fix[k + '_out_src']),
             np.array_equal(out['tgt_image'][n].numpy(), fix[k + '_out_tgt']),
             np.array_equal(out['src_label'][n].numpy(), fix[k + '_out_lab'])]
        print(k, plan[n], 'blur' if blur[n] else '-', e)
        ok &= all(e)
    return ok


def main():
    fix = np.load(os.path.join(ROOT, 'tests', 'golden', 'input_stage.npz'))
    plain = [str(k) for k in fix['cases']]
    blurred = [str(k) for k in fix['blur_cases']]
    ok = True
    for a, b in zip(plain[0::2], plain[1::2]):
        ok &= run_pair(fix, a, b)
    for a, b in zip(blurred[0::2], blurred[1::2]):          # the blur fires on both samples (own radius per image)
        ok &= run_pair(fix, a, b)
    ok &= run_pair(fix, blurred[0], plain[1])               # mixed batches: one sample blurred, one not
    ok &= run_pair(fix, plain[4], blurred[5])
    print("ALL OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
