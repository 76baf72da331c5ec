"""The reference's sample pipeline on the device.

`DeviceTrainTransform` is `TrainSet.transform_tr` (dataloders/datasets/gtav2cityscapes.py:66-74) -- RandomHorizontalFlip,
RandomScaleCrop, RandomGaussianBlur, Normalize, ToTensor from dataloders/custom_transforms.py -- plus `encode_segmap`
(:76-83), applied to uint8 images / labelId maps that are already in HBM; `DeviceValTransform` is
`ValSet.transform_val` (:139-146: FixedResize, Normalize, ToTensor).  The host computes what the reference computes on
the host -- the random draws, in the reference's order, Pillow's resampling tables and the fixed-point weights of its
box blur -- and the bytes are moved by csrc/input_stage.cu.  Outputs are bit-identical to the reference's tensors
(tests/golden/input_stage.npz).

No CPU path: tensors must be CUDA tensors and the C-ABI library must be present.
"""
import ctypes as C
import math
import random

import numpy as np
import torch

from .. import _lib as L

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
VOID_CLASSES = [0, 1, 2, 3, 4, 5, 6, 9, 10, 14, 15, 16, 18, 29, 30, 34, -1]
VALID_CLASSES = [7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33]
_PRECISION_BITS = 32 - 8 - 2


def segmap_lut(ignore_index=255):
    """encode_segmap (gtav2cityscapes.py:76-83) as a byte table: the sequential relabelling applied to 0..255."""
    m = np.arange(256, dtype=np.uint8)
    for v in VOID_CLASSES:
        if 0 <= v <= 255:
            m[m == v] = ignore_index
    for i, v in enumerate(VALID_CLASSES):
        m[m == v] = i
    return m


def _bilinear_tables(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c) for the bilinear filter, vectorised
    over the output positions with the per-position operations kept in Pillow's order (double precision)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = 0 + (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size) - xmin
    k = np.zeros((out_size, ksize), np.float64)
    ww = np.zeros(out_size, np.float64)
    for x in range(ksize):
        arg = np.abs((x + xmin - center + 0.5) * ss)
        w = np.where(arg < 1.0, 1.0 - arg, 0.0)
        w = np.where(x < xmax, w, 0.0)
        k[:, x] = w
        ww = ww + w
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz][:, None]
    kk = np.where(k < 0, np.trunc(-0.5 + k * (1 << _PRECISION_BITS)), np.trunc(0.5 + k * (1 << _PRECISION_BITS))).astype(np.int32)
    bounds = np.stack([xmin, xmax], 1).astype(np.int32)
    return bounds, kk, ksize


def _nearest_table(in_size, out_size):
    """Source index per output position of Pillow's ImagingScaleAffine: repeated double additions, truncation."""
    a = float(in_size) / out_size
    xo = a * 0.5
    tab = np.empty(out_size, np.int32)
    for x in range(out_size):
        xin = -1 if xo < 0.0 else int(xo)
        tab[x] = xin if 0 <= xin < in_size else -1
        xo += a
    return tab


def _gaussian_blur_weights(radius, passes=3):
    """The 24-bit weights (ww: centre pixel, fw: each neighbour) of one box-blur pass of PIL's
    ImageFilter.GaussianBlur(radius): libImaging/BoxBlur.c `_gaussian_blur_radius` (box radius whose `passes`-fold
    convolution has the Gaussian's variance; float locals, sqrt/floor in double) and ImagingHorizontalBoxBlur
    (`ww = (UINT32)(1 << 24) / (floatRadius * 2 + 1)` in single precision).  The kernels implement the case where the
    integer part of the box radius is 0, i.e. radius < sqrt(2); the reference draws radius from [0, 1)
    (custom_transforms.py:97-100)."""
    f = np.float32
    if radius == 0:                                    # ImageFilter.GaussianBlur.filter: plain copy
        return 1 << 24, 0
    r = f(radius)
    sigma2 = f(f(r * r) / f(passes))
    big_l = f(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = f(math.floor((float(big_l) - 1.0) / 2.0))
    if l != 0:
        raise NotImplementedError("GaussianBlur radius %r: box radius >= 1 (only radius < sqrt(2) is implemented; "
                                  "the reference draws [0, 1))" % (radius,))
    a = f(f(f(2) * l + f(1)) * f(f(l * f(l + f(1))) - f(f(3) * sigma2)))
    a = f(a / f(f(6) * f(sigma2 - f(f(l + f(1)) * f(l + f(1))))))
    fr = f(l + a)
    if fr == 0:                                        # ImagingBoxBlur skips an axis whose radius is 0
        return 1 << 24, 0
    ww = int(f(16777216.0) / f(f(fr * f(2)) + f(1)))
    fw = (((1 << 24) - ww) & 0xffffffff) // 2
    return ww, fw


def _scale_size(w, h, short_size):
    if h > w:
        ow = short_size
        oh = int(1.0 * h * ow / w)
    else:
        oh = short_size
        ow = int(1.0 * w * oh / h)
    return ow, oh


class _PinnedRing(object):
    """Staging slots in page-locked memory for the per-batch job tables.  A table that goes to the device from pageable
    memory makes the host wait until the copy has run, i.e. until everything queued on the stream before it has
    finished: with four or five tables per batch the stage ran at the pace of those waits (0.45 ms per batch for
    ~0.1 ms of kernels).  From a pinned slot the copy is asynchronous; a slot is reused only after its copy is done."""
    SLOTS, SIZE = 32, 1 << 16

    def __init__(self):
        self.buf, self.events, self.next = None, [None] * self.SLOTS, 0

    def put(self, data, dev):
        n = len(data)
        if self.buf is None:
            self.buf = torch.empty((self.SLOTS, self.SIZE), dtype=torch.uint8, pin_memory=True)
        k, self.next = self.next, (self.next + 1) % self.SLOTS
        if self.events[k] is not None:
            self.events[k].synchronize()
        self.buf[k, :n] = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        t = torch.empty(n, dtype=torch.uint8, device=dev)
        t.copy_(self.buf[k, :n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self.events[k] = ev
        return t


_RING = _PinnedRing()


def _upload(jobs, dev):
    """A ctypes job table -> device memory (asynchronously, through a pinned staging slot)."""
    data = bytes((type(jobs[0]) * len(jobs))(*jobs))
    dev = torch.device(dev)
    if dev.type != "cuda" or len(data) > _PinnedRing.SIZE:
        return torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    return _RING.put(data, dev)


class _Stage(object):
    def __init__(self, mean, std):
        self.mean = (C.c_double * 3)(*mean)
        self.std = (C.c_double * 3)(*std)
        self._tables = {}
        self._lut = {}

    @staticmethod
    def _check(t, shape_len, what):
        if t.device.type != "cuda":
            raise L.S2RError("%s must be a CUDA tensor (got %s); there is no CPU path" % (what, t.device))
        if t.dtype != torch.uint8 or t.dim() != shape_len or not t.is_contiguous():
            raise ValueError("%s: expected a contiguous uint8 tensor with %d dimensions" % (what, shape_len))

    def _dev(self, key, build, device):
        ent = self._tables.get((key, device))
        if ent is None:
            ent = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device) if isinstance(a, np.ndarray) else a for a in build())
            if len(self._tables) > 64:
                self._tables.clear()
            self._tables[(key, device)] = ent
        return ent

    def lut(self, device):
        t = self._lut.get(device)
        if t is None:
            t = torch.from_numpy(segmap_lut()).to(device)
            self._lut[device] = t
        return t

    def resize_image(self, img, ow, oh, flip, st):
        """PIL resize((ow, oh), BILINEAR) of the (mirrored) u8 [n,H,W,3] image; returns (tensor, flip still pending)."""
        n, H, W, _ = img.shape
        if W != ow:
            b, kk, ks = self._dev(("bl", W, ow), lambda: _bilinear_tables(W, ow), img.device)
            out = torch.empty((n, H, ow, 3), dtype=torch.uint8, device=img.device)
            L.call("s2r_resize_bilinear_u8", img.data_ptr(), n, H, W, 3, 1, ow, b.data_ptr(), kk.data_ptr(), ks, int(flip),
                   out.data_ptr(), st)
            img, flip, W = out, False, ow
        if H != oh:
            b, kk, ks = self._dev(("bl", H, oh), lambda: _bilinear_tables(H, oh), img.device)
            out = torch.empty((n, oh, W, 3), dtype=torch.uint8, device=img.device)
            L.call("s2r_resize_bilinear_u8", img.data_ptr(), n, H, W, 3, 0, oh, b.data_ptr(), kk.data_ptr(), ks, 0,
                   out.data_ptr(), st)
            img = out
        return img, flip

    def resize_label(self, lab, ow, oh, flip, st):
        n, H, W = lab.shape
        if (W, H) == (ow, oh):
            return lab, flip
        (xt,) = self._dev(("nn", W, ow), lambda: (_nearest_table(W, ow),), lab.device)
        (yt,) = self._dev(("nn", H, oh), lambda: (_nearest_table(H, oh),), lab.device)
        out = torch.empty((n, oh, ow), dtype=torch.uint8, device=lab.device)
        L.call("s2r_resize_nearest_u8", lab.data_ptr(), n, H, W, xt.data_ptr(), yt.data_ptr(), oh, ow, int(flip),
               out.data_ptr(), st)
        return out, False

    def finish(self, img, lab, flip_img, flip_lab, x1, y1, out_img, out_lab, H, W, st, use_lut=True, fill=255):
        n = img.shape[0] if img is not None else lab.shape[0]
        if img is not None:
            L.call("s2r_input_stage_u8", img.data_ptr(), None, n, img.shape[1], img.shape[2], int(flip_img), x1, y1,
                   C.cast(self.mean, C.c_void_p), C.cast(self.std, C.c_void_p), None, 255, out_img.data_ptr(), None, H, W, st)
        if lab is not None:
            L.call("s2r_input_stage_u8", None, lab.data_ptr(), n, lab.shape[1], lab.shape[2], int(flip_lab), x1, y1, None, None,
                   self.lut(lab.device).data_ptr() if use_lut else None, int(fill), None, out_lab.data_ptr(), H, W, st)


class DeviceTrainTransform(object):
    """transform_tr of the reference's TrainSet on a batch of equally sized uint8 device tensors:
    src/tgt images [N,H,W,3] (RGB, HWC as PIL decodes them) and labelId maps [N,H,W] ->
    {'src_image': f32 [N,3,crop,crop], 'tgt_image': ..., 'src_label': f32 [N,crop,crop]}.
    Per sample the draws are made with python's `random` in the reference's order (flip, short edge, x1, y1, blur and,
    when it fires, the source and the target image's blur radius) unless `draws` = [(flip, short_size, x1, y1[, blur,
    src_radius, tgt_radius]), ...] is given.  `last_draws` holds the draws of the last call.  gaussian_blur=False drops
    the blur (its draws are still consumed, so the random stream stays aligned with the reference's)."""

    def __init__(self, base_size, crop_size, mean=MEAN, std=STD, fill=255, gaussian_blur=True):
        self.base_size, self.crop_size, self.fill, self.gaussian_blur = base_size, crop_size, fill, gaussian_blur
        self.stage = _Stage(mean, std)
        self.last_draws = []

    def draw(self, w, h):
        flip = random.random() < 0.5                                           # custom_transforms.py:64
        short = random.randint(int(self.base_size * 0.5), int(self.base_size * 2.0))   # :119
        ow, oh = _scale_size(w, h, short)
        pw = max(ow, self.crop_size) if short < self.crop_size else ow         # :130-135
        ph = max(oh, self.crop_size) if short < self.crop_size else oh
        x1 = random.randint(0, pw - self.crop_size)                            # :138-139
        y1 = random.randint(0, ph - self.crop_size)
        blur = random.random() < 0.5                                           # :96
        r_src = random.random() if blur else None                              # :97-98 source image radius
        r_tgt = random.random() if blur else None                              # :99-100 target image radius
        return flip, short, x1, y1, blur, r_src, r_tgt

    def __call__(self, src_image, tgt_image, src_label, draws=None, batched=True):
        st_ = self.stage
        st_._check(src_image, 4, "src_image"); st_._check(tgt_image, 4, "tgt_image"); st_._check(src_label, 3, "src_label")
        N, H, W, _ = src_image.shape
        assert tgt_image.shape == src_image.shape and tuple(src_label.shape) == (N, H, W)
        dev, cs = src_image.device, self.crop_size
        out = {'src_image': torch.empty((N, 3, cs, cs), dtype=torch.float32, device=dev),
               'tgt_image': torch.empty((N, 3, cs, cs), dtype=torch.float32, device=dev),
               'src_label': torch.empty((N, cs, cs), dtype=torch.float32, device=dev)}
        self.last_draws = []
        plan, blur = [], []      # blur[n]: None or the (ww, fw) box-blur weights of sample n's source and target image
        for n in range(N):
            d = draws[n] if draws is not None else self.draw(W, H)
            flip, short, x1, y1 = d[:4]
            self.last_draws.append(tuple(d))
            ow, oh = _scale_size(W, H, short)
            if x1 + cs > max(ow, cs) or y1 + cs > max(oh, cs):
                raise ValueError("crop window (%d,%d)+%d outside the %dx%d scaled image" % (x1, y1, cs, ow, oh))
            plan.append((bool(flip), ow, oh, int(x1), int(y1)))
            fires = self.gaussian_blur and len(d) >= 7 and bool(d[4])
            blur.append({'src_image': _gaussian_blur_weights(d[5]), 'tgt_image': _gaussian_blur_weights(d[6])} if fires else None)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if batched:
                self._run_batched(src_image, tgt_image, src_label, plan, out, st, blur)
                return out
            for n in range(N):
                flip, ow, oh, x1, y1 = plan[n]
                for key, t in (('src_image', src_image), ('tgt_image', tgt_image)):
                    im, fl = st_.resize_image(t[n:n + 1], ow, oh, flip, st)
                    if blur[n] is not None:
                        im, _keep = self._blur([(im.data_ptr(), im.shape[1], im.shape[2], fl, x1, y1) + blur[n][key]], dev, st)
                        im, fl, x1_, y1_ = im.view(1, cs, cs, 3), False, 0, 0
                    else:
                        x1_, y1_ = x1, y1
                    st_.finish(im, None, fl, False, x1_, y1_, out[key][n:n + 1], None, cs, cs, st)
                lb, fl = st_.resize_label(src_label[n:n + 1], ow, oh, flip, st)
                st_.finish(None, lb, False, fl, x1, y1, None, out['src_label'][n:n + 1], cs, cs, st, fill=self.fill)
        return out


    def _blur(self, entries, dev, st):
        """RandomGaussianBlur on crops: entries = [(image pointer, Hs, Ws, mirror, x1, y1, ww, fw), ...] in the terms of
        a StageJob; returns (u8 [n][crop][crop][3] blurred crops, tensors to keep alive).  Two launches."""
        cs = self.crop_size
        n = len(entries)
        buf = torch.empty((2, n, cs, cs, 3), dtype=torch.uint8, device=dev)
        jobs = [L.BlurJob(p, buf[0, k].data_ptr(), buf[1, k].data_ptr(), int(hs), int(ws), int(fl), int(x1), int(y1),
                          int(ww), int(fw), 0) for k, (p, hs, ws, fl, x1, y1, ww, fw) in enumerate(entries)]
        tab = _upload(jobs, dev)
        L.call("s2r_gaussian_blur3_u8_multi", tab.data_ptr(), n, cs, cs, st)
        return buf[1], [buf, tab]

    def _run_batched(self, src_image, tgt_image, src_label, plan, out, st, blur=None):
        """At most four launches for the whole batch (column pass, row pass, nearest, crop/normalise) through device
        tables of per-sample jobs -- six when RandomGaussianBlur fires on a sample of the batch -- and only the WINDOW
        of every scaled image that its crop keeps is resampled: the column pass produces the window's columns for the
        source rows the row pass will read, the row pass the window's rows.  Same arithmetic per byte as the
        per-sample path (bit-identical)."""
        st_ = self.stage
        N, H, W, _ = src_image.shape
        dev, cs = src_image.device, self.crop_size
        blur = blur if blur is not None else [None] * N

        def upload(jobs):
            return _upload(jobs, dev)

        def pool(sizes):
            offs, tot = [], 0
            for sz in sizes:
                offs.append(tot)
                tot += (sz + 15) & ~15
            return torch.empty(max(tot, 16), dtype=torch.uint8, device=dev), offs

        def tables(a, b):
            """(device bounds, device coefficients, ksize, host bounds) of the bilinear resampling a -> b"""
            key = ("blh", a, b)
            ent = st_._tables.get((key, dev))
            if ent is None:
                bnd, kk, ks = _bilinear_tables(a, b)
                ent = (torch.from_numpy(bnd).to(dev), torch.from_numpy(kk).to(dev), ks, bnd)
                if len(st_._tables) > 64:
                    st_._tables.clear()
                st_._tables[(key, dev)] = ent
            return ent

        keep = []
        win = []      # per sample: window (cx0, cx1, cy0, cy1) in scaled coordinates and the source rows [r0, r1) it needs
        for flip, ow, oh, x1, y1 in plan:
            cx0, cx1, cy0, cy1 = x1, min(x1 + cs, ow), y1, min(y1 + cs, oh)
            if oh != H:
                bv = tables(H, oh)[3]
                r0, r1 = int(bv[cy0][0]), int(bv[cy1 - 1][0] + bv[cy1 - 1][1])
            else:
                r0, r1 = cy0, cy1
            win.append((cx0, cx1, cy0, cy1, r0, r1))
        images = [(key, t[n].data_ptr(), n) for n in range(N) for key, t in (('src_image', src_image), ('tgt_image', tgt_image))]
        # per image: [pointer to (row r0, first window column), row pitch in bytes, mirror still pending, windowed?]
        cur = {}
        for key, ptr, n in images:
            flip, ow, oh, x1, y1 = plan[n]
            cx0, cx1, cy0, cy1, r0, r1 = win[n]
            coloff = (W - cx1) if flip else cx0          # without a column pass the window is cut from the source columns
            cur[(key, n)] = [ptr + (r0 * W + coloff) * 3, W * 3, flip, False]
        # column pass
        todo = [(key, ptr, n) for key, ptr, n in images if plan[n][1] != W]
        if todo:
            buf, offs = pool([(win[n][5] - win[n][4]) * (win[n][1] - win[n][0]) * 3 for _, _, n in todo])
            keep.append(buf)
            jobs, mx = [], 0
            for (key, ptr, n), off in zip(todo, offs):
                flip, ow = plan[n][0], plan[n][1]
                cx0, cx1, cy0, cy1, r0, r1 = win[n]
                b, kk, ks, _ = tables(W, ow)
                ww, lines = cx1 - cx0, r1 - r0
                jobs.append(L.ResizeJob(ptr + r0 * W * 3, buf.data_ptr() + off, b.data_ptr(), kk.data_ptr(), W, 3, ks, int(flip), 1,
                                        cx0, ww, lines, W * 3, 0))
                cur[(key, n)] = [buf.data_ptr() + off, ww * 3, False, True]
                mx = max(mx, lines * ww * 3)
            tab = upload(jobs); keep.append(tab)
            L.call("s2r_resize_bilinear_u8_multi", tab.data_ptr(), len(jobs), mx, st)
        # row pass
        todo = [(key, n) for key, _, n in images if plan[n][2] != H]
        if todo:
            buf, offs = pool([(win[n][3] - win[n][2]) * (win[n][1] - win[n][0]) * 3 for _, n in todo])
            keep.append(buf)
            jobs, mx = [], 0
            for (key, n), off in zip(todo, offs):
                oh = plan[n][2]
                cx0, cx1, cy0, cy1, r0, r1 = win[n]
                b, kk, ks, _ = tables(H, oh)
                c = cur[(key, n)]
                ww, wh = cx1 - cx0, cy1 - cy0
                jobs.append(L.ResizeJob(c[0], buf.data_ptr() + off, b.data_ptr(), kk.data_ptr(), 0, 3, ks, 0, 0,
                                        cy0, wh, ww * 3, c[1], r0))
                cur[(key, n)] = [buf.data_ptr() + off, ww * 3, c[2], True]
                mx = max(mx, wh * ww * 3)
            tab = upload(jobs); keep.append(tab)
            L.call("s2r_resize_bilinear_u8_multi", tab.data_ptr(), len(jobs), mx, st)
        # labels: nearest, window only
        lab = {}
        todo = [n for n in range(N) if (plan[n][1], plan[n][2]) != (W, H)]
        if todo:
            buf, offs = pool([(win[n][3] - win[n][2]) * (win[n][1] - win[n][0]) for n in todo])
            keep.append(buf)
            jobs, mx = [], 0
            for n, off in zip(todo, offs):
                flip, ow, oh = plan[n][0], plan[n][1], plan[n][2]
                cx0, cx1, cy0, cy1, r0, r1 = win[n]
                (xt,) = st_._dev(("nn", W, ow), lambda W=W, ow=ow: (_nearest_table(W, ow),), dev)
                (yt,) = st_._dev(("nn", H, oh), lambda H=H, oh=oh: (_nearest_table(H, oh),), dev)
                jobs.append(L.NearestJob(src_label[n].data_ptr(), buf.data_ptr() + off, xt.data_ptr(), yt.data_ptr(), W,
                                         cy1 - cy0, cx1 - cx0, cx0, cy0, int(flip)))
                lab[n] = buf.data_ptr() + off
                mx = max(mx, (cy1 - cy0) * (cx1 - cx0))
            tab = upload(jobs); keep.append(tab)
            L.call("s2r_resize_nearest_u8_multi", tab.data_ptr(), len(jobs), mx, st)
        # crop window + Normalize/ToTensor (images) and the labelId table (labels): one launch.  A resampled image is
        # already cut to its window (origin 0, 0); an image that was not resampled at all is read in place.
        # (a crop on which RandomGaussianBlur fires is first cut and blurred as uint8, then normalised like an image of
        # the crop's size)
        geom = {}
        for key, ptr, n in images:
            flip, ow, oh, x1, y1 = plan[n]
            cx0, cx1, cy0, cy1, r0, r1 = win[n]
            c = cur[(key, n)]
            geom[(key, n)] = (c[0], cy1 - cy0, cx1 - cx0, int(c[2]), 0, 0) if c[3] else (ptr, H, W, int(flip), x1, y1)
        fired = [(key, n) for key, _, n in images if blur[n] is not None]
        if fired:
            crops, kept = self._blur([geom[kn] + blur[kn[1]][kn[0]] for kn in fired], dev, st)
            keep.extend(kept)
            for k, kn in enumerate(fired):
                geom[kn] = (crops[k].data_ptr(), cs, cs, 0, 0, 0)
        jobs = []
        for key, ptr, n in images:
            p, hs, ws, fl, gx1, gy1 = geom[(key, n)]
            jobs.append(L.StageJob(p, None, out[key][n].data_ptr(), None, hs, ws, fl, gx1, gy1, 0))
        for n in range(N):
            flip, ow, oh, x1, y1 = plan[n]
            cx0, cx1, cy0, cy1, r0, r1 = win[n]
            if n in lab:
                jobs.append(L.StageJob(None, lab[n], None, out['src_label'][n].data_ptr(), cy1 - cy0, cx1 - cx0, 0, 0, 0, 0))
            else:
                jobs.append(L.StageJob(None, src_label[n].data_ptr(), None, out['src_label'][n].data_ptr(), H, W, int(flip),
                                       x1, y1, 0))
        tab = upload(jobs); keep.append(tab)
        L.call("s2r_input_stage_u8_multi", tab.data_ptr(), len(jobs), C.cast(st_.mean, C.c_void_p), C.cast(st_.std, C.c_void_p),
               st_.lut(dev).data_ptr(), int(self.fill), cs, cs, st)
        return keep


def _fix_scale_crop_geometry(w, h, crop_size):
    """custom_transforms_eval.py:132-145 (FixScaleCrop): short edge -> crop_size, then the centre window; python's
    round() (half to even) as the reference calls it."""
    if w > h:
        oh = crop_size
        ow = int(1.0 * w * oh / h)
    else:
        ow = crop_size
        oh = int(1.0 * h * ow / w)
    return ow, oh, int(round((ow - crop_size) / 2.)), int(round((oh - crop_size) / 2.))


class DeviceValTransform(object):
    """The reference's evaluation pipelines + encode_segmap on a batch: images u8 [N,H,W,3], labelId maps u8 [N,H,W] ->
    {'image': f32 [N,3,size,size], 'label': f32 [N,size,size]}.
    mode='fixed_resize': FixedResize((size, size)), Normalize, ToTensor -- ValSet / TestSet.transform_val of
    gtav2cityscapes.py:139-146,213-219 and gta5.py's transform_ts;
    mode='fix_scale_crop': FixScaleCrop(size), Normalize, ToTensor -- gta5.py:81-88 (short edge to `size`, centre crop)."""

    def __init__(self, crop_size, mean=MEAN, std=STD, mode='fixed_resize'):
        if mode not in ('fixed_resize', 'fix_scale_crop'):
            raise NotImplementedError(mode)
        self.size, self.mode = crop_size, mode
        self.stage = _Stage(mean, std)

    def __call__(self, image, label):
        st_ = self.stage
        st_._check(image, 4, "image"); st_._check(label, 3, "label")
        N, H, W, _ = image.shape
        assert tuple(label.shape) == (N, H, W)                               # custom_transforms_eval.py:159
        dev, s = image.device, self.size
        ow, oh, x1, y1 = (s, s, 0, 0) if self.mode == 'fixed_resize' else _fix_scale_crop_geometry(W, H, s)
        out = {'image': torch.empty((N, 3, s, s), dtype=torch.float32, device=dev),
               'label': torch.empty((N, s, s), dtype=torch.float32, device=dev)}
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            im, _ = st_.resize_image(image, ow, oh, False, st)
            # the reference relabels BEFORE the nearest resize; a per-pixel table commutes with a nearest-neighbour copy
            lb, _ = st_.resize_label(label, ow, oh, False, st)
            st_.finish(im, lb, False, False, x1, y1, out['image'], out['label'], s, s, st)
        return out
