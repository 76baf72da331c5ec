"""CPU restatement (numpy; test infrastructure only) of the reference's prediction export, test_adapt.py:118-157
(`imgsaver`) after the host argmax of test_adapt.py:170-171; PIL's NEAREST resize is oracle.input_stage.resize_nearest
(pinned against Pillow by tests/golden/make_golden_input.py).  Pinned against the reference's own `imgsaver`, compiled
unmodified out of test_adapt.py, by tests/golden/make_golden_report.py (tests/golden/report.npz)."""
import numpy as np

from .input_stage import nearest_table

VALID_CLASSES = [7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33]
PALETTE = [[128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153], [250, 170, 30],
           [220, 220, 0], [107, 142, 35], [152, 251, 152], [70, 130, 180], [220, 20, 60], [255, 0, 0], [0, 0, 142], [0, 0, 70],
           [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32]]


def imgsaver_arrays(logits_chw, out_w=1280, out_h=640):
    """(labelId image [out_h,out_w] u8, colour image [out_h,out_w,3] u8) for one prediction [C,H,W]."""
    pred = np.argmax(logits_chw[None], axis=1)                 # test_adapt.py:170-171
    im1 = np.uint8(pred.transpose(1, 2, 0)).squeeze()            # :119
    class_map = dict(zip(range(19), VALID_CLASSES))
    im1_np = np.uint8(np.zeros(im1.shape))                       # the reference hard-codes [512,512]
    for c in range(19):
        im1_np[im1 == c] = class_map[c]
    class_color_map = dict(zip(range(19), PALETTE))
    im2_np = np.uint8(np.zeros(im1.shape + (3,)))
    for c in range(19):
        im2_np[im1 == c] = class_color_map[c]
    xt, yt = nearest_table(im1.shape[1], out_w), nearest_table(im1.shape[0], out_h)
    return im1_np[np.ix_(yt, xt)], im2_np[np.ix_(yt, xt)]
