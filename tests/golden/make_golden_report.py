"""Generates tests/golden/report.npz from the REAL reference prediction export (build container only): the method
`Trainer.imgsaver` is taken UNMODIFIED out of /root/reference/test_adapt.py (the script itself cannot be imported --
tensorboardX is missing and it builds CUDA loaders at import -- so the function's source segment is compiled on its
own), run on the host argmax of test_adapt.py:170-171, and the two PNG files it writes are read back.  Also checks
oracle/report.py against them.  Run:  python tests/golden/make_golden_report.py
"""
import ast
import os
import sys
import tempfile
import types

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("S2R_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import report as OR  # noqa: E402


def reference_imgsaver():
    src = open(os.path.join(REF, "test_adapt.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "imgsaver":
            ns = {"np": np, "Image": Image}
            exec(compile(ast.Module(body=[node], type_ignores=[]), "test_adapt.py", "exec"), ns)
            return ns["imgsaver"]
    raise RuntimeError("imgsaver not found")


def logits(seed):
    """[19, 512, 512] float32: blocky class evidence + noise, with exact ties (np.argmax keeps the lowest index)."""
    rng = np.random.RandomState(seed)
    coarse = rng.rand(19, 16, 16).astype(np.float32) * 8
    x = np.kron(coarse, np.ones((32, 32), np.float32)) + 0.5 * rng.rand(19, 512, 512).astype(np.float32)
    x[3] = x[7]
    x[:, :16] = 0.0
    return x


def main():
    imgsaver = reference_imgsaver()
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "result"))
        os.chdir(d)
        try:
            for seed in (3, 4):
                x = logits(seed)
                pred = np.argmax(x[None], axis=1)                       # test_adapt.py:170-171
                imgsaver(types.SimpleNamespace(), pred, "a%d.png" % seed)
                ids = np.array(Image.open("result/a%d.png" % seed))
                rgb = np.array(Image.open("result/a%d_color.png" % seed))
                want_ids, want_rgb = OR.imgsaver_arrays(x)
                assert ids.shape == (640, 1280) and rgb.shape == (640, 1280, 3)
                assert np.array_equal(ids, want_ids) and np.array_equal(rgb, want_rgb), seed
                out["ids%d" % seed], out["rgb%d" % seed] = ids, rgb
        finally:
            os.chdir(cwd)
    out["seeds"] = np.array([3, 4], np.int32)
    np.savez_compressed(os.path.join(HERE, "report.npz"), **out)
    print("oracle == reference imgsaver on 2 predictions; fixture written")


if __name__ == "__main__":
    main()
