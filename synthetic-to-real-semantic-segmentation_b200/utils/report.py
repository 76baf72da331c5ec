"""Validation reporting and prediction export around the Evaluator.

* `validation_report` / `write_val_info`: the text block val_adapt.py:159-166 appends to val_info.txt (per-class IoU
  table included), from an `Evaluator` whose confusion matrix was accumulated on the device.
* `PredictionExporter`: test_adapt.py:118-157 (`imgsaver`) -- trainId -> Cityscapes labelId image and palette image,
  both NEAREST-resized to the output size -- fused with the host argmax of test_adapt.py:170-171 into one kernel over
  the logits (csrc/evaluator.cu); only PNG encoding stays on the host.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _lib as L
from ..dataloders.device_transforms import _nearest_table

# val_adapt.py:140-158
CLASS_NAMES = ["road", "sidewalk", "building", "wall", "fence", "pole", "light", "sign", "vegetation", "terrain", "sky",
               "person", "rider", "car", "truck", "bus", "train", "motocycle", "bicycle"]
# test_adapt.py:122 / :131-149
VALID_CLASSES = [7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33]
PALETTE = [[128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153], [250, 170, 30],
           [220, 220, 0], [107, 142, 35], [152, 251, 152], [70, 130, 180], [220, 20, 60], [255, 0, 0], [0, 0, 142], [0, 0, 70],
           [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32]]


def validation_report(evaluator, epoch, num_images, test_loss):
    """The block of val_adapt.py:159-166, character for character."""
    Acc = evaluator.Pixel_Accuracy()
    Acc_class = evaluator.Pixel_Accuracy_Class()
    mIoU, IoU = evaluator.Mean_Intersection_over_Union()
    FWIoU = evaluator.Frequency_Weighted_Intersection_over_Union()
    out = ['Validation:' + '\n',
           '[Epoch: %d, numImages: %5d]' % (epoch, num_images) + '\n',
           "Acc:{}, Acc_class:{}, mIoU:{}, fwIoU: {}".format(Acc, Acc_class, mIoU, FWIoU) + '\n',
           'Loss: %.3f' % test_loss + '\n' + '\n',
           'Class IOU: ' + '\n']
    for idx in range(19):
        out.append('\t' + CLASS_NAMES[idx] + (': \t' if len(CLASS_NAMES[idx]) > 5 else ': \t\t') + str(IoU[idx]) + '\n')
    return ''.join(out)


def write_val_info(evaluator, epoch, num_images, test_loss, path='val_info.txt'):
    text = validation_report(evaluator, epoch, num_images, test_loss)
    with open(path, 'a') as f1:
        f1.write(text)
    return text


class PredictionExporter(object):
    def __init__(self, out_size=(1280, 640), valid_classes=VALID_CLASSES, palette=PALETTE):
        self.out_w, self.out_h = out_size          # PIL size convention (width, height), test_adapt.py:127
        self.ids = np.asarray(valid_classes, np.uint8)
        self.rgb = np.asarray(palette, np.uint8).reshape(-1, 3)
        assert len(self.ids) == len(self.rgb)
        self._dev = {}

    def _tables(self, device, H, W):
        key = (device, H, W)
        t = self._dev.get(key)
        if t is None:
            t = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in
                      (_nearest_table(W, self.out_w), _nearest_table(H, self.out_h), self.ids, self.rgb))
            self._dev[key] = t
        return t

    def __call__(self, logits):
        """logits: f32 [N,C,H,W] on the device -> (labelId image u8 [N,out_h,out_w], colour image u8 [N,out_h,out_w,3])."""
        if logits.device.type != "cuda":
            raise L.S2RError("PredictionExporter runs on CUDA tensors only (got %s); there is no CPU path" % logits.device)
        if logits.dim() != 4:
            raise ValueError("expected [N,C,H,W] logits")
        x = logits.detach().contiguous().float()
        N, Cc, H, W = x.shape
        xt, yt, idt, rgbt = self._tables(x.device, H, W)
        ids = torch.empty((N, self.out_h, self.out_w), dtype=torch.uint8, device=x.device)
        rgb = torch.empty((N, self.out_h, self.out_w, 3), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            L.call("s2r_export_prediction_nchw", x.data_ptr(), N, Cc, H, W, xt.data_ptr(), yt.data_ptr(), self.out_h, self.out_w,
                   idt.data_ptr(), rgbt.data_ptr(), len(self.ids), ids.data_ptr(), rgb.data_ptr(),
                   C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        return ids, rgb

    def save(self, ids, rgb, names, out_dir='result', miou=None):
        """PNG files as test_adapt.py:128,156 (and val_adapt.py:217 when miou is given) name them."""
        from PIL import Image
        os.makedirs(out_dir, exist_ok=True)
        ids, rgb = ids.cpu().numpy(), rgb.cpu().numpy()
        for n, name in enumerate(names):
            if miou is None:
                Image.fromarray(ids[n], mode='L').save(os.path.join(out_dir, name))
                Image.fromarray(rgb[n]).save(os.path.join(out_dir, name[:-4] + '_color.png'))
            else:
                Image.fromarray(rgb[n]).save(os.path.join(out_dir, name[:-4] + '_color_' + str(miou) + '_.png'))
