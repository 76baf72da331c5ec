"""Generates tests/golden/input_stage.npz from the REAL reference input pipeline (build container only):
`dataloders.datasets.gtav2cityscapes.TrainSet.__getitem__` / `ValSet.__getitem__` run UNMODIFIED on small synthetic
PNG files (PIL resize / flip / pad / crop, GaussianBlur, Normalize, ToTensor), with `random` seeded so that the draws
are known: per image configuration two cases on which RandomGaussianBlur does not fire (train*) and two on which it
does (trainb*).  Also checks oracle/input_stage.py against every case and against Pillow itself over a sweep of sizes
and blur radii.  Run:  python tests/golden/make_golden_input.py
"""
import os
import random
import sys
import tempfile
import types

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("S2R_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from dataloders.datasets import gtav2cityscapes as ref_ds  # noqa: E402
from oracle import input_stage as OI  # noqa: E402


def draws(seed, n_tgt, w, h, base_size, crop_size):
    """Replays the reference's calls to `random` in TrainSet.__getitem__ + transform_tr."""
    random.seed(seed)
    tgt_index = random.randint(0, n_tgt - 1)
    flip = random.random() < 0.5
    short = random.randint(int(base_size * 0.5), int(base_size * 2.0))
    ow, oh = OI.scale_size(w, h, short)
    pw = max(ow, crop_size) if short < crop_size else ow
    ph = max(oh, crop_size) if short < crop_size else oh
    x1 = random.randint(0, pw - crop_size)
    y1 = random.randint(0, ph - crop_size)
    blur = random.random() < 0.5
    radii = (random.random(), random.random()) if blur else None      # custom_transforms.py:97-100: src, then tgt
    return tgt_index, flip, short, x1, y1, blur, radii


def main():
    rng = np.random.RandomState(1234)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for sub in ("src", "lab", "tgt", "vimg", "vlab"):
            os.makedirs(os.path.join(d, sub))
        cases, blur_cases = [], []
        # (H, W, base_size, crop_size): up-scaling, down-scaling with padding, portrait (h > w)
        for ci, (H, W, base, crop) in enumerate([(40, 64, 40, 32), (48, 36, 20, 32), (33, 57, 36, 24), (64, 96, 24, 40)]):
            for f in os.listdir(os.path.join(d, "src")):
                os.remove(os.path.join(d, "src", f)); os.remove(os.path.join(d, "lab", f)); os.remove(os.path.join(d, "tgt", f))
            src = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
            tgt = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
            lab = rng.randint(0, 36, (H, W)).astype(np.uint8)
            lab[rng.rand(H, W) < 0.03] = 255
            # smooth the images a little so that bilinear weights matter more than noise
            Image.fromarray(src).save(os.path.join(d, "src", "a.png"))
            Image.fromarray(lab).save(os.path.join(d, "lab", "a.png"))
            Image.fromarray(tgt).save(os.path.join(d, "tgt", "a.png"))
            args = types.SimpleNamespace(src_img_root=os.path.join(d, "src"), src_label_root=os.path.join(d, "lab"),
                                         tgt_img_root=os.path.join(d, "tgt"), base_size=base, crop_size=crop)
            ds = ref_ds.TrainSet(args)
            found = 0
            for seed in range(1000):
                _, flip, short, x1, y1, blur, _ = draws(seed, 1, W, H, base, crop)
                if blur:
                    continue
                want_flip = found % 2 == 0
                if flip != want_flip:
                    continue
                random.seed(seed)
                sample = ds[0]                      # the reference, unmodified
                got_img, got_lab = OI.train_sample(src, lab, flip, short, crop, x1, y1)
                got_tgt, _ = OI.train_sample(tgt, lab, flip, short, crop, x1, y1)
                assert np.array_equal(sample['src_image'].numpy(), got_img), ("src image", ci, seed)
                assert np.array_equal(sample['tgt_image'].numpy(), got_tgt), ("tgt image", ci, seed)
                assert np.array_equal(sample['src_label'].numpy(), got_lab), ("label", ci, seed)
                k = "train%d_%d" % (ci, found)
                out[k + "_src"], out[k + "_tgt"], out[k + "_lab"] = src, tgt, lab
                out[k + "_draw"] = np.array([int(flip), short, crop, x1, y1], np.int32)
                out[k + "_out_src"] = sample['src_image'].numpy()
                out[k + "_out_tgt"] = sample['tgt_image'].numpy()
                out[k + "_out_lab"] = sample['src_label'].numpy()
                cases.append(k)
                found += 1
                if found == 2:
                    break
            assert found == 2
            # the same images on draws where RandomGaussianBlur fires (own radius for the source and the target image)
            found = 0
            for seed in range(1000, 2000):
                _, flip, short, x1, y1, blur, radii = draws(seed, 1, W, H, base, crop)
                if not blur or flip != (found % 2 == 1):
                    continue
                random.seed(seed)
                sample = ds[0]                      # the reference, unmodified
                got_img, got_lab = OI.train_sample(src, lab, flip, short, crop, x1, y1, blur_radius=radii[0])
                got_tgt, _ = OI.train_sample(tgt, lab, flip, short, crop, x1, y1, blur_radius=radii[1])
                assert np.array_equal(sample['src_image'].numpy(), got_img), ("blurred src image", ci, seed)
                assert np.array_equal(sample['tgt_image'].numpy(), got_tgt), ("blurred tgt image", ci, seed)
                assert np.array_equal(sample['src_label'].numpy(), got_lab), ("label", ci, seed)
                k = "trainb%d_%d" % (ci, found)
                out[k + "_src"], out[k + "_tgt"], out[k + "_lab"] = src, tgt, lab
                out[k + "_draw"] = np.array([int(flip), short, crop, x1, y1], np.int32)
                out[k + "_radii"] = np.array(radii, np.float64)
                out[k + "_out_src"] = sample['src_image'].numpy()
                out[k + "_out_tgt"] = sample['tgt_image'].numpy()
                out[k + "_out_lab"] = sample['src_label'].numpy()
                blur_cases.append(k)
                found += 1
                if found == 2:
                    break
            assert found == 2
        # validation pipeline: FixedResize((size, size)) + Normalize + ToTensor (ValSet.transform_val)
        img = rng.randint(0, 256, (50, 70, 3)).astype(np.uint8)
        lab = rng.randint(0, 36, (50, 70)).astype(np.uint8)
        Image.fromarray(img).save(os.path.join(d, "vimg", "x_leftImg8bit.png"))
        Image.fromarray(lab).save(os.path.join(d, "vlab", "x_gtFine_labelIds.png"))
        vargs = types.SimpleNamespace(val_img_root=os.path.join(d, "vimg"), val_label_root=os.path.join(d, "vlab"), crop_size=36)
        vs = ref_ds.ValSet(vargs)[0]
        want_img = OI.normalize_to_tensor(OI.resize_bilinear(img, 36, 36))
        want_lab = OI.resize_nearest(OI.encode_segmap(lab), 36, 36).astype(np.float32)
        assert np.array_equal(vs['image'].numpy(), want_img) and np.array_equal(vs['label'].numpy(), want_lab)
        out["val_img"], out["val_lab"], out["val_size"] = img, lab, np.array([36], np.int32)
        out["val_out_img"], out["val_out_lab"] = vs['image'].numpy(), vs['label'].numpy()
        # gta5.py:81-88 transform_val: FixScaleCrop + Normalize + ToTensor -- the reference's own transform classes on
        # a landscape and a portrait image (the label is relabelled first, as ValSet.__getitem__ does)
        from torchvision import transforms as tv_transforms
        from dataloders import custom_transforms_eval as tr_e
        fsc = tv_transforms.Compose([tr_e.FixScaleCrop(crop_size=36), tr_e.Normalize(mean=OI.MEAN, std=OI.STD), tr_e.ToTensor()])
        for tag, (a, m) in (("land", (img, lab)), ("port", (np.ascontiguousarray(img.transpose(1, 0, 2)), np.ascontiguousarray(lab.T)))):
            res = fsc({'image': Image.fromarray(a), 'label': Image.fromarray(OI.encode_segmap(m))})
            h, w = m.shape
            ow, oh = (int(1.0 * w * 36 / h), 36) if w > h else (36, int(1.0 * h * 36 / w))
            x1, y1 = int(round((ow - 36) / 2.)), int(round((oh - 36) / 2.))
            want_img = OI.normalize_to_tensor(OI.resize_bilinear(a, ow, oh)[y1:y1 + 36, x1:x1 + 36])
            want_lab = OI.resize_nearest(OI.encode_segmap(m), ow, oh)[y1:y1 + 36, x1:x1 + 36].astype(np.float32)
            assert np.array_equal(res['image'].numpy(), want_img) and np.array_equal(res['label'].numpy(), want_lab), tag
            out["fsc_%s_img" % tag], out["fsc_%s_lab" % tag] = a, m
            out["fsc_%s_out_img" % tag], out["fsc_%s_out_lab" % tag] = res['image'].numpy(), res['label'].numpy()
        out["cases"] = np.array(cases)
        out["blur_cases"] = np.array(blur_cases)
    # the label table against the reference's own relabelling of every byte value
    holder = types.SimpleNamespace(void_classes=OI.VOID_CLASSES, valid_classes=OI.VALID_CLASSES, ignore_index=255,
                                   class_map=dict(zip(OI.VALID_CLASSES, range(19))))
    lut_ref = ref_ds.TrainSet.encode_segmap(holder, np.arange(256, dtype=np.uint8))
    assert np.array_equal(lut_ref, OI.segmap_lut())
    out["lut"] = lut_ref
    # Pillow itself over a sweep of sizes (not stored): both resampling restatements are bit-exact
    n = 0
    for (h, w) in [(17, 23), (40, 64), (31, 90)]:
        a = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        m = rng.randint(0, 256, (h, w)).astype(np.uint8)
        for ow in list(range(5, 40, 3)) + [64, 97, 150]:
            for oh in (7, 17, 40, 83):
                assert np.array_equal(np.array(Image.fromarray(a).resize((ow, oh), Image.BILINEAR)), OI.resize_bilinear(a, ow, oh)), (h, w, ow, oh)
                assert np.array_equal(np.array(Image.fromarray(m).resize((ow, oh), Image.NEAREST)), OI.resize_nearest(m, ow, oh)), (h, w, ow, oh)
                n += 1
    # Pillow's GaussianBlur over radii of the reference's range [0, 1) and beyond, on small and degenerate sizes
    from PIL import ImageFilter
    nb = 0
    for trial in range(600):
        h, w = int(rng.randint(1, 40)), int(rng.randint(1, 40))
        a = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        r = float(rng.rand()) * (1.0, 6.0, 40.0)[trial % 3]
        assert np.array_equal(np.array(Image.fromarray(a).filter(ImageFilter.GaussianBlur(radius=r))), OI.gaussian_blur(a, r)), (h, w, r)
        nb += 1
    print("oracle == Pillow on %d resize shapes and %d blurs; %d pipeline cases written" % (n, nb, len(cases) + len(blur_cases) + 1))
    np.savez_compressed(os.path.join(HERE, "input_stage.npz"), **out)


if __name__ == "__main__":
    main()
