"""Device input stage (dataloders.device_transforms) on GTA5 -> Cityscapes shaped uint8 batches: time per batch of 8
source/target pairs (1052x1914 RGB + labelIds -> 3x512x512 crops, random scale in [0.5, 2] x base 512) with CUDA
events, next to the same pipeline on the host through Pillow + numpy (what the reference's DataLoader workers run,
custom_transforms.py:59-147,17-56) on one core.  GPU box:  python tests/tools/input_stage_bench.py"""
import os, sys, time, random
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sub
dt = sub("dataloders.device_transforms")
dev = torch.device("cuda", 0)
N, H, W, BASE, CROP = 8, 1052, 1914, 512, 512
g = torch.Generator().manual_seed(0)
src = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, generator=g)
tgt = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, generator=g)
lab = torch.randint(0, 35, (N, H, W), dtype=torch.uint8, generator=g)
d_src, d_tgt, d_lab = src.to(dev), tgt.to(dev), lab.to(dev)
tr = dt.DeviceTrainTransform(BASE, CROP)
random.seed(0)
draws = [tr.draw(W, H)[:4] for _ in range(N)]
for _ in range(2):
    out = tr(d_src, d_tgt, d_lab, draws=draws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = tr(d_src, d_tgt, d_lab, draws=draws)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
in_bytes = N * H * W * 7
print("device input stage (batched launches): %.2f ms per batch of %d pairs (%.0f pairs/s), %.1f GB/s of source bytes; draws %s"
      % (ms, N, N / ms * 1e3, in_bytes / ms / 1e6, draws[:2]))
for _ in range(2):
    tr(d_src, d_tgt, d_lab, draws=draws, batched=False)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    out1 = tr(d_src, d_tgt, d_lab, draws=draws, batched=False)
e1.record(); torch.cuda.synchronize()
print("device input stage (per-sample launches): %.2f ms per batch; identical to the batched result: %s"
      % (e0.elapsed_time(e1) / 5, all(torch.equal(out[k], out1[k]) for k in out)))

# the same arithmetic on the host through Pillow (one core), sample 0..N-1
from PIL import Image, ImageOps
mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
lut = dt.segmap_lut()
t0 = time.time()
ok = True
for n in range(N):
    flip, short, x1, y1 = draws[n]
    ims = [Image.fromarray(src[n].numpy()), Image.fromarray(tgt[n].numpy()), Image.fromarray(lut[lab[n].numpy()])]
    if flip:
        ims = [im.transpose(Image.FLIP_LEFT_RIGHT) for im in ims]
    ow, oh = dt._scale_size(W, H, short)
    ims = [ims[0].resize((ow, oh), Image.BILINEAR), ims[1].resize((ow, oh), Image.BILINEAR), ims[2].resize((ow, oh), Image.NEAREST)]
    if short < CROP:
        padh, padw = (CROP - oh if oh < CROP else 0), (CROP - ow if ow < CROP else 0)
        ims = [ImageOps.expand(ims[0], border=(0, 0, padw, padh), fill=0), ImageOps.expand(ims[1], border=(0, 0, padw, padh), fill=0),
               ImageOps.expand(ims[2], border=(0, 0, padw, padh), fill=255)]
    ims = [im.crop((x1, y1, x1 + CROP, y1 + CROP)) for im in ims]
    res = []
    for im in ims[:2]:
        a = np.array(im).astype(np.float32); a /= 255.0; a -= mean; a /= std
        res.append(a.transpose((2, 0, 1)))
    m = np.array(ims[2]).astype(np.float32)
    ok &= np.array_equal(res[0], out['src_image'][n].cpu().numpy()) and np.array_equal(res[1], out['tgt_image'][n].cpu().numpy())
    ok &= np.array_equal(m, out['src_label'][n].cpu().numpy())
cpu_ms = (time.time() - t0) * 1e3
print("Pillow + numpy on 1 host core: %.1f ms per batch (includes the comparison); device output bit-identical: %s" % (cpu_ms, ok))

# RandomGaussianBlur on every sample of the batch (the reference fires it on half of them), own radius per image,
# against Pillow's GaussianBlur on the host
from PIL import ImageFilter
bdraws = [tuple(d) + (True, random.random(), random.random()) for d in draws]
for _ in range(2):
    outb = tr(d_src, d_tgt, d_lab, draws=bdraws)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    outb = tr(d_src, d_tgt, d_lab, draws=bdraws)
e1.record(); torch.cuda.synchronize()
ms_b = e0.elapsed_time(e1) / 5
ok = True
for n in range(N):
    flip, short, x1, y1, _, r_src, r_tgt = bdraws[n]
    for key, t, r in (('src_image', src, r_src), ('tgt_image', tgt, r_tgt)):
        im = Image.fromarray(t[n].numpy())
        if flip:
            im = im.transpose(Image.FLIP_LEFT_RIGHT)
        ow, oh = dt._scale_size(W, H, short)
        im = im.resize((ow, oh), Image.BILINEAR)
        if short < CROP:
            im = ImageOps.expand(im, border=(0, 0, CROP - ow if ow < CROP else 0, CROP - oh if oh < CROP else 0), fill=0)
        im = im.crop((x1, y1, x1 + CROP, y1 + CROP)).filter(ImageFilter.GaussianBlur(radius=r))
        a = np.array(im).astype(np.float32); a /= 255.0; a -= mean; a /= std
        ok &= np.array_equal(a.transpose((2, 0, 1)), outb[key][n].cpu().numpy())
print("with RandomGaussianBlur on all %d pairs: %.2f ms per batch (+%.2f ms for the two blur launches); identical to Pillow: %s"
      % (N, ms_b, ms_b - ms, ok))

# device time of the two blur launches alone (job table resident, CUDA-graph replay of 10 calls): sixteen 512x512
# crops cut out of 8 + 8 scaled images
L = sub("_lib")
scaled = [torch.randint(0, 256, (700, 1273, 3), dtype=torch.uint8, generator=g).to(dev) for _ in range(2 * N)]
buf = torch.empty((2, 2 * N, CROP, CROP, 3), dtype=torch.uint8, device=dev)
ww, fw = dt._gaussian_blur_weights(0.7)
jobs = [L.BlurJob(scaled[k].data_ptr(), buf[0, k].data_ptr(), buf[1, k].data_ptr(), 700, 1273, k & 1, 37 + k, 11 + k,
                  int(ww), int(fw), 0) for k in range(2 * N)]
tab = dt._upload(jobs, dev)
st = torch.cuda.current_stream().cuda_stream
call = lambda: L.call("s2r_gaussian_blur3_u8_multi", tab.data_ptr(), 2 * N, CROP, CROP, st)
call(); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(10):
        call()
gr.replay(); torch.cuda.synchronize()
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
nbytes = 2 * N * CROP * CROP * 3
print("blur launches alone (device time, 16 crops of %dx%d): %.1f us per batch = %.0f GB/s over the 4 x %.1f MB read + written"
      % (CROP, CROP, us, 4 * nbytes / us / 1e3, nbytes / 1e6))

# device time of the whole batched stage: the launches of 5 calls are enqueued behind a 30 ms spin kernel, so the events
# see the GPU run them back to back (the 0.45 ms above is what the Python call costs end to end: four launches plus the
# job tables and their upload)
for rnd, with_blur, d in ((0, False, draws), (0, True, bdraws), (1, False, draws), (1, True, bdraws)):   # round 0: warm-up
    torch.cuda.synchronize()
    torch.cuda._sleep(int(30e-3 * 1.9e9))
    e0.record()
    for _ in range(5):
        tr(d_src, d_tgt, d_lab, draws=d)
    e1.record(); torch.cuda.synchronize()
    if rnd:
        print("batched stage%s, device time: %.3f ms per batch" % (" with RandomGaussianBlur on all pairs" if with_blur else "", e0.elapsed_time(e1) / 5))
