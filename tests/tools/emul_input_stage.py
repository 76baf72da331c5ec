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


def view(ptr, shape, dtype=np.uint8):
    return arr(ptr, int(np.prod(shape)), dtype).reshape(shape)


def resize_bilinear(inp, N, H, W, Cc, axis, out_size, bounds, kk, ksize, flip, out, stream):
    """resize_h_kernel (axis 1, mirrored source columns when flip) / resize_v_kernel (axis 0), vectorised over the lines"""
    x = view(inp, (N, H, W, Cc)).astype(np.int64)
    bnd, k = view(bounds, (out_size, 2), np.int32), view(kk, (out_size, ksize), np.int32)
    if axis == 1:
        x = x[:, :, ::-1] if flip else x
        o = view(out, (N, H, out_size, Cc))
        for xx in range(out_size):
            acc = np.full((N, H, Cc), 1 << (PB - 1), np.int64)
            for t in range(bnd[xx, 1]):
                acc += x[:, :, bnd[xx, 0] + t] * int(k[xx, t])
            o[:, :, xx] = np.clip(acc >> PB, 0, 255)
    else:
        o = view(out, (N, out_size, W, Cc))
        for yy in range(out_size):
            acc = np.full((N, W, Cc), 1 << (PB - 1), np.int64)
            for t in range(bnd[yy, 1]):
                acc += x[:, bnd[yy, 0] + t] * int(k[yy, t])
            o[:, yy] = np.clip(acc >> PB, 0, 255)


def resize_nearest(inp, N, H, W, xtab, ytab, OH, OW, flip, out, stream):
    """resize_nearest_kernel"""
    x, o = view(inp, (N, H, W)), view(out, (N, OH, OW))
    xt, yt = view(xtab, (OW,), np.int32), view(ytab, (OH,), np.int32)
    o[:] = 0
    for y in range(OH):
        for xx in range(OW):
            if xt[xx] >= 0 and yt[y] >= 0:
                o[:, y, xx] = x[:, yt[y], (W - 1 - xt[xx]) if flip else xt[xx]]


def input_stage(img, label, N, Hs, Ws, flip, x1, y1, mean, std, lut, fill, out_img, out_label, H, W, stream):
    """input_stage_kernel: window (x1, y1) of the mirrored image, zero / fill padded on the right and bottom"""
    for n in range(N):
        if img:
            x = view(img, (N, Hs, Ws, 3))[n]
            x = x[:, ::-1] if flip else x
            pad = np.zeros((max(Hs, y1 + H), max(Ws, x1 + W), 3), np.uint8)
            pad[:Hs, :Ws] = x
            view(out_img, (N, 3, H, W), np.float32)[n] = OI.normalize_to_tensor(pad[y1:y1 + H, x1:x1 + W])
        if label:
            x = view(label, (N, Hs, Ws))[n]
            x = x[:, ::-1] if flip else x
            if lut:
                x = view(lut, (256,))[x]
            pad = np.full((max(Hs, y1 + H), max(Ws, x1 + W)), fill, np.uint8)
            pad[:Hs, :Ws] = x
            view(out_label, (N, H, W), np.float32)[n] = pad[y1:y1 + H, x1:x1 + W].astype(np.float32)


KERNELS = {"s2r_resize_bilinear_u8": resize_bilinear, "s2r_resize_nearest_u8": resize_nearest, "s2r_input_stage_u8": input_stage,
           "s2r_resize_bilinear_u8_multi": resize_bilinear_multi, "s2r_resize_nearest_u8_multi": resize_nearest_multi,
           "s2r_input_stage_u8_multi": input_stage_multi, "s2r_gaussian_blur3_u8_multi": gaussian_blur3_multi}


def emu(name, *args):
    KERNELS[name](*args)
    return 0


L.call = emu
dt.L.call = emu


class on_cpu(object):
    """Lets DeviceTrainTransform / DeviceValTransform .__call__ run on CPU tensors: the device checks and the stream
    lookup are the only CUDA touches on the host side (the kernels are the emulations above)."""

    def __enter__(self):
        import contextlib
        import types
        self.saved = (dt._Stage._check, torch.cuda.device, torch.cuda.current_stream)
        dt._Stage._check = staticmethod(lambda t, n, what: None)
        torch.cuda.device = lambda d: contextlib.nullcontext()
        torch.cuda.current_stream = lambda d=None: types.SimpleNamespace(cuda_stream=0)

    def __exit__(self, *exc):
        dt._Stage._check, torch.cuda.device, torch.cuda.current_stream = self.saved


def run_calls(fix):
    """The public __call__ paths: validation pipelines (FixedResize; FixScaleCrop landscape / portrait) and the
    per-sample (batched=False) training path with and without RandomGaussianBlur."""
    ok = True
    with on_cpu():
        s = int(fix['val_size'][0])
        out = dt.DeviceValTransform(s)(torch.from_numpy(fix['val_img'])[None], torch.from_numpy(fix['val_lab'])[None])
        e = [np.array_equal(out['image'][0].numpy(), fix['val_out_img']), np.array_equal(out['label'][0].numpy(), fix['val_out_lab'])]
        print('val fixed_resize', e)
        ok &= all(e)
        for tag in ('land', 'port'):
            out = dt.DeviceValTransform(36, mode='fix_scale_crop')(torch.from_numpy(fix['fsc_%s_img' % tag])[None],
                                                                   torch.from_numpy(fix['fsc_%s_lab' % tag])[None])
            e = [np.array_equal(out['image'][0].numpy(), fix['fsc_%s_out_img' % tag]),
                 np.array_equal(out['label'][0].numpy(), fix['fsc_%s_out_lab' % tag])]
            print('val fix_scale_crop', tag, e)
            ok &= all(e)
        for k in (str(fix['cases'][0]), str(fix['cases'][3]), str(fix['blur_cases'][1]), str(fix['blur_cases'][2])):
            flip, short, crop, x1, y1 = (int(v) for v in fix[k + '_draw'])
            r = fix[k + '_radii'] if (k + '_radii') in fix.files else None
            d = (bool(flip), short, x1, y1) + ((True, float(r[0]), float(r[1])) if r is not None else ())
            out = dt.DeviceTrainTransform(1, crop)(torch.from_numpy(fix[k + '_src'])[None], torch.from_numpy(fix[k + '_tgt'])[None],
                                                   torch.from_numpy(fix[k + '_lab'])[None], draws=[d], batched=False)
            e = [np.array_equal(out['src_image'][0].numpy(), fix[k + '_out_src']),
                 np.array_equal(out['tgt_image'][0].numpy(), fix[k + '_out_tgt']),
                 np.array_equal(out['src_label'][0].numpy(), fix[k + '_out_lab'])]
            print('per-sample path', k, e)
            ok &= all(e)
    return ok


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
    ok &= run_calls(fix)
    print("ALL OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
