#!/usr/bin/env python
"""Generates tests/golden/* (run in the BUILD container, where /root/reference exists).

1. sample_gray_400x320.u8 -- testdata/sample.jpg decoded with Pillow (libjpeg; NOT bit-identical
   to Go's image/jpeg decoder) and converted with the formula of core/grayscale.go:8-23.
   The parity contract is defined on the grayscale buffer, so this file *is* the input.
2. oracle_vectors.npz -- outputs of the C oracle on that buffer and on small synthetic frames.
   These are ORACLE-generated regression vectors (the reference is Go and cannot run here):
   they pin the oracle against drift and travel to the GPU box; they are not reference outputs.
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from oracle import np_oracle as NP  # noqa: E402
from pigo_b200 import synth  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
CASC = os.path.join(ROOT, "pigo_b200", "data", "cascade")


def main():
    os.makedirs(GOLD, exist_ok=True)
    img = Image.open(os.path.join(REF, "testdata", "sample.jpg")).convert("RGB")
    rgb = np.asarray(img, dtype=np.uint8)
    gray = NP.rgb_to_grayscale(rgb)
    rows, cols = gray.shape
    assert (rows, cols) == (400, 320)
    gray.tofile(os.path.join(GOLD, "sample_gray_400x320.u8"))

    face = O.OracleFace(open(os.path.join(CASC, "facefinder"), "rb").read())
    vec = {}
    # (a) reference test params on sample.jpg (core/pigo_test.go:44-50)
    d = face.run_cascade(gray, rows, cols, cols, 20, 1000, 0.2, 1.1, 0.0)
    vec["sample_test_dets"] = d
    vec["sample_test_clusters"] = O.cluster(d, 0.1)[1]
    # (b) doc params (README: shift 0.1) and CLI defaults (cmd/pigo/main.go:110-111)
    vec["sample_doc_dets"] = face.run_cascade(gray, rows, cols, cols, 20, 1000, 0.1, 1.1, 0.0)
    vec["sample_cli_dets"] = face.run_cascade(gray, rows, cols, cols, 20, 1000, 0.15, 1.15, 0.0)
    # (c) rotated path: every table slot k/32 on the sample
    for k in (1, 5, 8, 16, 27, 32):
        vec[f"sample_rot{k}_dets"] = face.run_cascade(gray, rows, cols, cols, 20, 1000, 0.2, 1.1, k / 32.0)
    # (d) a 1080p tiled-faces frame (class F) at test params
    f1080 = synth.frame_faces(gray, 1080, 1920, shift=(0, 0))
    d = face.run_cascade(f1080, 1080, 1920, 1920, 20, 1000, 0.2, 1.1, 0.0)
    vec["f1080_test_dets"] = d
    vec["f1080_test_clusters"] = O.cluster(d, 0.2)[1]
    # (e) puploc / flploc with injected randoms on the sample's face
    pl = O.OraclePuploc(open(os.path.join(CASC, "puploc"), "rb").read())
    rnd = np.random.default_rng(7).random(3 * 63, dtype=np.float32)
    vec["puploc_randoms"] = rnd
    cl = vec["sample_test_clusters"]
    big = [c for c in cl if c["scale"] > 50]
    assert len(big) == 1
    det = big[0]
    r0 = int(det["row"]) - int(np.float32(0.075) * np.float32(det["scale"]))
    cL = int(det["col"]) - int(np.float32(0.175) * np.float32(det["scale"]))
    cR = int(det["col"]) + int(np.float32(0.185) * np.float32(det["scale"]))
    sc = np.float32(det["scale"]) * np.float32(0.25)
    res = []
    for P in (63, 50, 1):
        for (cc, ang, fl) in ((cL, 0.0, False), (cR, 0.0, False), (cL, 0.1, False), (cR, 0.0, True), (cR, 0.37, True)):
            o = pl.run_detector(r0, cc, float(sc), P, rnd, gray, rows, cols, cols, ang, fl)
            res.append((r0, cc, float(sc), P, ang, int(fl), o[0], o[1], float(o[2])))
    vec["puploc_cases"] = np.array(res, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "oracle_vectors.npz"), **vec)
    print("sample dets:", vec["sample_test_dets"])
    print("sample clusters:", vec["sample_test_clusters"])
    print("f1080 dets:", len(vec["f1080_test_dets"]), "clusters:", len(vec["f1080_test_clusters"]))


if __name__ == "__main__":
    main()
