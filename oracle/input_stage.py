"""CPU restatement (numpy, test infrastructure only) of the reference's per-sample input pipeline -- the row
"input stage" of SURVEY.md section 8(f):

  dataloders/datasets/gtav2cityscapes.py:76-83   encode_segmap (labelId -> trainId, void -> 255)
  dataloders/custom_transforms.py:59-71           RandomHorizontalFlip
  dataloders/custom_transforms.py:108-147         RandomScaleCrop  (resize, pad right/bottom, crop window)
  dataloders/custom_transforms.py:150-176         FixScaleCrop
  dataloders/custom_transforms.py:92-105          RandomGaussianBlur (PIL ImageFilter.GaussianBlur)
  dataloders/custom_transforms.py:17-33           Normalize
  dataloders/custom_transforms.py:36-56           ToTensor (HWC -> CHW float32)

The reference does the geometry on PIL images; the resampling arithmetic therefore lives in Pillow (third party, not
vendored; pinned here against Pillow 12.2.0 as installed in the build container, algorithm unchanged since 4.x):
`Image.resize(BILINEAR)` is the two-pass fixed-point convolution of libImaging/Resample.c (22-bit coefficients, uint8
intermediate), `Image.resize(NEAREST)` the incremental-coordinate copy of libImaging/Geometry.c (ImagingScaleAffine).
`Image.filter(GaussianBlur(radius))` is three passes of the fractional-radius box blur of libImaging/BoxBlur.c per axis
(24-bit fixed-point weights derived in single precision, uint8 intermediate after every pass, edge pixels replicated).
All are restated below and pinned by tests/golden/make_golden_input.py, which runs the REAL reference classes on PIL
images and stores their outputs (tests/golden/input_stage.npz).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import math

import numpy as np

VOID_CLASSES = [0, 1, 2, 3, 4, 5, 6, 9, 10, 14, 15, 16, 18, 29, 30, 34, -1]
VALID_CLASSES = [7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33]
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2


def encode_segmap(mask, ignore_index=255):
    """gtav2cityscapes.py:76-83, statement by statement (sequential in-place relabelling of a uint8 array)."""
    mask = np.array(mask, dtype=np.uint8)
    class_map = dict(zip(VALID_CLASSES, range(19)))
    for v in VOID_CLASSES:
        if 0 <= v <= 255:            # `mask == -1` never matches a uint8 array
            mask[mask == v] = ignore_index
    for v in VALID_CLASSES:
        mask[mask == v] = class_map[v]
    return mask


def segmap_lut(ignore_index=255):
    """The relabelling as a 256-entry table: what the sequential algorithm does to every possible byte."""
    return encode_segmap(np.arange(256, dtype=np.uint8), ignore_index)


def normalize_to_tensor(img_u8, mean=MEAN, std=STD):
    """Normalize (custom_transforms.py:17-33) then ToTensor (:36-56) on an HWC uint8 image.  numpy semantics kept:
    the /255 runs in float32, the tuple operands make `-= mean` and `/= std` run in float64 and round to float32."""
    x = np.array(img_u8).astype(np.float32)
    x /= 255.0
    x -= mean
    x /= std
    return np.array(x).astype(np.float32).transpose((2, 0, 1))


# ------------------------------------------------------------------------------------------------ Pillow resampling
def _bilinear_filter(x):
    if x < 0.0:
        x = -x
    if x < 1.0:
        return 1.0 - x
    return 0.0


def precompute_coeffs(in_size, out_size):
    """libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0) over the
    whole input range: (bounds[out][2] = (xmin, count), integer coefficients [out][ksize])."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = 0 + (xx + 0.5) * scale
        ww = 0.0
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = _bilinear_filter((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(a, out_size, axis):
    """One pass of ImagingResampleHorizontal/Vertical_8bpc along `axis` of a uint8 array (int32 accumulation starting
    at 1 << (PRECISION_BITS - 1), arithmetic shift, clip to [0, 255])."""
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    bounds, kk = precompute_coeffs(a.shape[0], out_size)
    out = np.empty((out_size,) + a.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.full(a.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += a[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear(img_u8, ow, oh):
    """PIL.Image.resize((ow, oh), Image.BILINEAR) on an HWC (or HW) uint8 array: horizontal pass first, then vertical,
    each skipped when the size does not change (Resample.c ImagingResample)."""
    a = np.asarray(img_u8, dtype=np.uint8)
    if a.shape[1] != ow:
        a = _resample_axis(a, ow, 1)
    if a.shape[0] != oh:
        a = _resample_axis(a, oh, 0)
    return a


def nearest_table(in_size, out_size):
    """Source index per output position of ImagingScaleAffine (Geometry.c): xo starts at scale/2 and is advanced by
    repeated addition in double precision; COORD truncates."""
    a = float(in_size) / out_size
    xo = 0.0 + a * 0.5
    tab = np.zeros(out_size, np.int32)
    for x in range(out_size):
        xin = -1 if xo < 0.0 else int(xo)
        tab[x] = min(max(xin, 0), in_size - 1) if 0 <= xin < in_size else -1
        xo += a
    return tab


def resize_nearest(mask_u8, ow, oh):
    """PIL.Image.resize((ow, oh), Image.NEAREST) on an HW uint8 array (positions outside the source stay 0)."""
    m = np.asarray(mask_u8, dtype=np.uint8)
    xt, yt = nearest_table(m.shape[1], ow), nearest_table(m.shape[0], oh)
    out = np.zeros((oh, ow), np.uint8)
    ys, xs = np.nonzero(yt >= 0)[0], np.nonzero(xt >= 0)[0]
    out[np.ix_(ys, xs)] = m[np.ix_(yt[ys], xt[xs])]
    return out


# ------------------------------------------------------------------------------------------------ Pillow Gaussian blur
_F = np.float32


def gaussian_blur_radius(radius, passes=3):
    """libImaging/BoxBlur.c _gaussian_blur_radius: the (fractional) box radius whose `passes`-fold convolution has the
    variance of a Gaussian of standard deviation `radius`.  C semantics kept: the arguments and locals are floats,
    sqrt/floor are evaluated in double on double literals."""
    r = _F(radius)
    sigma2 = _F(_F(r * r) / _F(passes))
    big_l = _F(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = _F(math.floor((float(big_l) - 1.0) / 2.0))
    a = _F(_F(_F(2) * l + _F(1)) * _F(_F(l * _F(l + _F(1))) - _F(_F(3) * sigma2)))
    a = _F(a / _F(_F(6) * _F(sigma2 - _F(_F(l + _F(1)) * _F(l + _F(1))))))
    return _F(l + a)


def box_blur_weights(float_radius):
    """ImagingHorizontalBoxBlur's fixed-point weights: (radius, ww, fw) = integer radius, the weight of the 2*radius+1
    inner pixels and of the two fractional outer pixels, in units of 2^-24 (the division runs in single precision)."""
    fr = _F(float_radius)
    radius = int(fr)
    ww = int(_F(16777216.0) / _F(_F(fr * _F(2)) + _F(1)))
    fw = (((1 << 24) - (radius * 2 + 1) * ww) & 0xffffffff) // 2
    return radius, ww, fw


def box_blur_axis1(a, float_radius):
    """One ImagingHorizontalBoxBlur (ImagingLineBoxBlur32 per line) along axis 1 of a uint8 [H][W][C] array: running
    sum `acc` over the inner window in uint32, out = (acc*ww + (left + right)*fw + 2^23) >> 24, indices beyond the
    line replaced by its first / last pixel."""
    radius, ww, fw = box_blur_weights(float_radius)
    w = a.shape[1]
    lastx = w - 1
    edge_a, edge_b = min(radius + 1, w), max(w - radius - 1, 0)
    src = a.astype(np.int64)
    out = np.empty_like(a)
    m32 = 0xffffffff
    acc = src[:, 0] * (radius + 1)
    for x in range(edge_a - 1):
        acc = acc + src[:, x]
    acc = (acc + src[:, lastx] * (radius - edge_a + 1)) & m32

    def step(acc, sub, add, left, right, x):
        acc = (acc + src[:, add] - src[:, sub]) & m32
        bulk = (acc * ww + (src[:, left] + src[:, right]) * fw) & m32
        out[:, x] = (((bulk + (1 << 23)) & m32) >> 24).astype(np.uint8)
        return acc

    if edge_a <= edge_b:
        for x in range(0, edge_a):
            acc = step(acc, 0, x + radius, 0, x + radius + 1, x)
        for x in range(edge_a, edge_b):
            acc = step(acc, x - radius - 1, x + radius, x - radius - 1, x + radius + 1, x)
        for x in range(edge_b, lastx + 1):
            acc = step(acc, x - radius - 1, lastx, x - radius - 1, lastx, x)
    else:
        for x in range(0, edge_b):
            acc = step(acc, 0, x + radius, 0, x + radius + 1, x)
        for x in range(edge_b, edge_a):
            acc = step(acc, 0, lastx, 0, lastx, x)
        for x in range(edge_a, lastx + 1):
            acc = step(acc, x - radius - 1, lastx, x - radius - 1, lastx, x)
    return out


def gaussian_blur(img_u8, radius, passes=3):
    """PIL.Image.filter(ImageFilter.GaussianBlur(radius)) on an HWC uint8 array (ImageFilter.py GaussianBlur.filter ->
    ImagingGaussianBlur -> ImagingBoxBlur: `passes` box blurs along x, transpose, `passes` along y, transpose)."""
    a = np.asarray(img_u8, dtype=np.uint8)
    if radius == 0:
        return a.copy()
    fr = gaussian_blur_radius(radius, passes)
    if fr != 0:
        for _ in range(passes):
            a = box_blur_axis1(a, fr)
        t = np.ascontiguousarray(a.transpose(1, 0, 2))
        for _ in range(passes):
            t = box_blur_axis1(t, fr)
        a = np.ascontiguousarray(t.transpose(1, 0, 2))
    return a


# ------------------------------------------------------------------------------------------------ geometry
def scale_size(w, h, short_size):
    """custom_transforms.py:117-123 (RandomScaleCrop) -- also FixScaleCrop's :157-162 with short_size = crop_size and
    the branches swapped the way the reference writes them."""
    if h > w:
        ow = short_size
        oh = int(1.0 * h * ow / w)
    else:
        oh = short_size
        ow = int(1.0 * w * oh / h)
    return ow, oh


def flip_scale_pad_crop(img_u8, mask_u8, flip, short_size, crop_size, x1, y1, fill=255):
    """RandomHorizontalFlip (:59-71) + RandomScaleCrop (:108-147) with the random draws given: returns the
    crop_size x crop_size uint8 image and mask."""
    img, mask = np.asarray(img_u8), np.asarray(mask_u8)
    if flip:
        img, mask = img[:, ::-1], mask[:, ::-1]
    h, w = mask.shape
    ow, oh = scale_size(w, h, short_size)
    img, mask = resize_bilinear(img, ow, oh), resize_nearest(mask, ow, oh)
    if short_size < crop_size:
        padh = crop_size - oh if oh < crop_size else 0
        padw = crop_size - ow if ow < crop_size else 0
        img = np.pad(img, ((0, padh), (0, padw), (0, 0)), constant_values=0)
        mask = np.pad(mask, ((0, padh), (0, padw)), constant_values=fill)
    return img[y1:y1 + crop_size, x1:x1 + crop_size], mask[y1:y1 + crop_size, x1:x1 + crop_size]


def train_sample(img_u8, label_ids_u8, flip, short_size, crop_size, x1, y1, blur_radius=None):
    """TrainSet.__getitem__ + transform_tr (gtav2cityscapes.py:49-74) for one image/label pair with the random draws
    given: (float32 CHW image, float32 HW label).  blur_radius = the radius RandomGaussianBlur drew for this image
    (custom_transforms.py:96-100) or None when it did not fire; the label is never blurred."""
    mask = encode_segmap(label_ids_u8)
    img, mask = flip_scale_pad_crop(img_u8, mask, flip, short_size, crop_size, x1, y1, fill=255)
    if blur_radius is not None:
        img = gaussian_blur(img, blur_radius)
    return normalize_to_tensor(img), np.array(mask).astype(np.float32)
