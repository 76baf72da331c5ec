"""Device input stage, the transforms added last:
* RandomGaussianBlur (dataloders/custom_transforms.py:92-105 of the reference): bit-exact against the tensors the
  reference's unmodified TrainSet produced on draws where the blur fires (tests/golden/input_stage.npz, trainb* cases)
  and against the numpy restatement of Pillow's GaussianBlur;
* FixScaleCrop evaluation pipeline (gta5.py:81-88): bit-exact against the reference's own transform classes."""
import random

import numpy as np
import pytest
import torch

from conftest import golden, sub

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dt(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return sub("dataloders.device_transforms")


def _draw(fix, k):
    flip, short, crop, x1, y1 = (int(v) for v in fix[k + "_draw"])
    r = fix[k + "_radii"] if (k + "_radii") in fix.files else None
    d = (bool(flip), short, x1, y1) + ((True, float(r[0]), float(r[1])) if r is not None else (False, None, None))
    return crop, d


def _stack(fix, names, key):
    return torch.from_numpy(np.stack([fix[k + key] for k in names])).cuda()


def _check(out, fix, names, tag):
    for n, k in enumerate(names):
        assert np.array_equal(out['src_image'][n].cpu().numpy(), fix[k + "_out_src"]), (k, tag, "src")
        assert np.array_equal(out['tgt_image'][n].cpu().numpy(), fix[k + "_out_tgt"]), (k, tag, "tgt")
        assert np.array_equal(out['src_label'][n].cpu().numpy(), fix[k + "_out_lab"]), (k, tag, "label")


def test_blurred_samples_bit_exact_against_reference_fixture(dt):
    fix = golden("input_stage")
    blurred = [str(k) for k in fix["blur_cases"]]
    plain = [str(k) for k in fix["cases"]]
    assert len(blurred) == 8
    for k in blurred:                                   # one sample at a time
        crop, d = _draw(fix, k)
        tr = dt.DeviceTrainTransform(base_size=1, crop_size=crop)
        for batched in (True, False):
            out = tr(_stack(fix, [k], "_src"), _stack(fix, [k], "_tgt"), _stack(fix, [k], "_lab"), draws=[d], batched=batched)
            _check(out, fix, [k], batched)
    # two blurred samples per batch, and mixed batches (the blur fires on one sample only)
    for i in range(0, 8, 2):
        for names in ((blurred[i], blurred[i + 1]), (blurred[i], plain[i + 1]), (plain[i], blurred[i + 1])):
            crop, d0 = _draw(fix, names[0])
            _, d1 = _draw(fix, names[1])
            tr = dt.DeviceTrainTransform(base_size=1, crop_size=crop)
            for batched in (True, False):
                out = tr(_stack(fix, names, "_src"), _stack(fix, names, "_tgt"), _stack(fix, names, "_lab"), draws=[d0, d1],
                         batched=batched)
                _check(out, fix, names, batched)
    # gaussian_blur=False drops the blur only
    k = blurred[0]
    crop, d = _draw(fix, k)
    out = dt.DeviceTrainTransform(1, crop, gaussian_blur=False)(_stack(fix, [k], "_src"), _stack(fix, [k], "_tgt"),
                                                                 _stack(fix, [k], "_lab"), draws=[d])
    assert np.array_equal(out['src_label'][0].cpu().numpy(), fix[k + "_out_lab"])
    assert not np.array_equal(out['src_image'][0].cpu().numpy(), fix[k + "_out_src"])


def test_blur_random_stream_and_full_size_against_oracle(dt):
    from oracle import input_stage as OI
    # the reference's order of draws: flip, short edge, x1, y1, blur, source radius, target radius
    random.seed(3)
    tr = dt.DeviceTrainTransform(base_size=40, crop_size=32)
    g = torch.Generator().manual_seed(5)
    src = torch.randint(0, 256, (2, 40, 64, 3), dtype=torch.uint8, generator=g)
    tgt = torch.randint(0, 256, (2, 40, 64, 3), dtype=torch.uint8, generator=g)
    lab = torch.randint(0, 34, (2, 40, 64), dtype=torch.uint8, generator=g)
    out = tr(src.cuda(), tgt.cuda(), lab.cuda())
    assert [bool(d[4]) for d in tr.last_draws] == [False, True]
    for n, (flip, short, x1, y1, blur, r_src, r_tgt) in enumerate(tr.last_draws):
        want_src, want_lab = OI.train_sample(src[n].numpy(), lab[n].numpy(), flip, short, 32, x1, y1, r_src if blur else None)
        want_tgt, _ = OI.train_sample(tgt[n].numpy(), lab[n].numpy(), flip, short, 32, x1, y1, r_tgt if blur else None)
        assert np.array_equal(out['src_image'][n].cpu().numpy(), want_src), n
        assert np.array_equal(out['tgt_image'][n].cpu().numpy(), want_tgt), n
        assert np.array_equal(out['src_label'][n].cpu().numpy(), want_lab), n
    # full-size crops (512 x 512 out of 1024 x 2048, no resize) with the extreme radii of the reference's range
    img = torch.randint(0, 256, (2, 1024, 2048, 3), dtype=torch.uint8, generator=g)
    big = torch.randint(0, 34, (2, 1024, 2048), dtype=torch.uint8, generator=g)
    tr = dt.DeviceTrainTransform(base_size=1024, crop_size=512)
    draws = [(True, 1024, 700, 300, True, 0.999999, 1e-3), (False, 1024, 1536, 512, True, 0.0, 0.5)]
    out = tr(img.cuda(), img.cuda(), big.cuda(), draws=draws)
    for n, (flip, short, x1, y1, _, r_src, r_tgt) in enumerate(draws):
        a = img[n].numpy()[:, ::-1] if flip else img[n].numpy()
        crop = a[y1:y1 + 512, x1:x1 + 512]
        assert np.array_equal(out['src_image'][n].cpu().numpy(), OI.normalize_to_tensor(OI.gaussian_blur(crop, r_src))), n
        assert np.array_equal(out['tgt_image'][n].cpu().numpy(), OI.normalize_to_tensor(OI.gaussian_blur(crop, r_tgt))), n
    with pytest.raises(NotImplementedError):
        tr(img.cuda(), img.cuda(), big.cuda(), draws=[(False, 1024, 0, 0, True, 2.0, 0.1)] * 2)


def test_fix_scale_crop_validation_pipeline_against_reference_fixture(dt):
    """DeviceValTransform(mode='fix_scale_crop') = FixScaleCrop + Normalize + ToTensor (gta5.py:81-88,
    custom_transforms_eval.py:125-149) against the reference's own transform classes on a landscape and a portrait
    image: bit-exact."""
    fix = golden("input_stage")
    for tag in ("land", "port"):
        img = torch.from_numpy(np.stack([fix["fsc_%s_img" % tag]] * 2)).cuda()
        lab = torch.from_numpy(np.stack([fix["fsc_%s_lab" % tag]] * 2)).cuda()
        out = dt.DeviceValTransform(36, mode='fix_scale_crop')(img, lab)
        for n in range(2):
            assert np.array_equal(out['image'][n].cpu().numpy(), fix["fsc_%s_out_img" % tag]), tag
            assert np.array_equal(out['label'][n].cpu().numpy(), fix["fsc_%s_out_lab" % tag]), tag
    with pytest.raises(NotImplementedError):
        dt.DeviceValTransform(36, mode='random')
