#!/usr/bin/env python
"""Writes the raw grayscale buffers and the manifest tools/refdump/main.go consumes (oracle/_ref/inputs/).

The buffers are the ones the parity tests already use (the committed 400x320 luma of testdata/sample.jpg and synthetic
frames built from it), so a dump of the true Go outputs pins the oracle on exactly the inputs the GPU path is checked on."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pigo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "oracle", "_ref", "inputs")
TEST, DOC, CLI = (20, 1000, 0.2, 1.1), (20, 1000, 0.1, 1.1), (20, 1000, 0.15, 1.15)


def run(prm, angle=0.0, iou=(0.0, 0.1, 0.15, 0.2)):
    return {"min_size": prm[0], "max_size": prm[1], "shift_factor": prm[2], "scale_factor": prm[3], "angle": angle, "iou": list(iou)}


def main():
    os.makedirs(OUT, exist_ok=True)
    sample = synth.sample_gray()
    wide = synth.frame_faces(sample, 300, 900, noise_seed=1)
    f1080 = synth.frame_faces(sample, 1080, 1920)
    smooth = synth.frame_smooth(540, 960, seed=3)
    man = []

    def add(name, img, rows, cols, dim, runs, pupils=(), landmarks=()):
        img.tofile(os.path.join(OUT, name))
        man.append({"file": name, "rows": rows, "cols": cols, "dim": dim, "runs": list(runs), "pupils": list(pupils), "landmarks": list(landmarks)})

    pupils = [{"cascade": "puploc", "row": 186, "col": 118, "scale": 60.0, "perturbs": 50, "angle": 0.0, "flipv": False, "seed": 1}]
    for k, (ang, fl) in enumerate([(0.0, False), (0.0, True), (0.2, False), (0.93, True)]):
        pupils.append({"cascade": "puploc", "row": 188, "col": 205, "scale": 55.5, "perturbs": 63, "angle": ang, "flipv": fl, "seed": 10 + k})
    lms = [{"cascade": "lps/" + n, "left_row": 186, "left_col": 118, "right_row": 188, "right_col": 205, "perturbs": 63, "flipv": fl, "seed": 100 + i}
           for i, (n, fl) in enumerate([("lp42", False), ("lp42", True), ("lp93", False), ("lp84", True)])]
    add("sample_gray_400x320.u8", sample, 400, 320, 320,
        [run(TEST), run(DOC), run(CLI)] + [run(TEST, k / 32.0, iou=(0.1,)) for k in (1, 8, 13, 16, 27, 32)], pupils, lms)
    add("wide_300x900.u8", wide, 300, 900, 900, [run(TEST), run(TEST, 0.05, iou=(0.1,)), run(TEST, 0.5, iou=(0.1,)), run(TEST, 0.97, iou=(0.1,))])
    add("faces_1080x1920.u8", f1080, 1080, 1920, 1920, [run(TEST), run(DOC, iou=(0.2,))])
    add("smooth_540x960.u8", smooth, 540, 960, 960, [run(TEST), run((0, 40, 0.3, 1.4), iou=(0.1,))])
    json.dump(man, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
    print("wrote", len(man), "inputs to", OUT)


if __name__ == "__main__":
    main()
