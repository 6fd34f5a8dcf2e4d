#!/usr/bin/env python
"""Exercises every kernel of libpigo_b200 on small inputs, for use under compute-sanitizer:

    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py
    compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py --quick

No torch (ctypes mirror only), a few hundred thousand windows in total, so that the instrumented run ends in minutes.
Covers: fused kernel (tile + gather role, aligned TMA fill and byte-wise fill), gather-v2 with Q1, deep kernel with Q2,
queue-overflow fallbacks (tiny KS, tail policy on), universal gather kernel (rotated, scan_mode=1), finalize, cluster,
RunDetector / landmark kernels over resident frames, grayscale.  Prints the detection counts (compare with a plain run).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pigo_b200  # noqa: E402
from pigo_b200 import CascadeParams, ImageParams, pipeline, synth  # noqa: E402

PRM = (20, 1000, 0.2, 1.1)


def cp_of(img, rows, cols, dim, prm=PRM):
    return CascadeParams(ImageParams(img, rows, cols, dim), *prm)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="scan kernels only, fewer variants (racecheck is slow)")
    args = ap.parse_args()
    pigo_b200.init(0)
    clf = pigo_b200.NewPigo().Unpack(pigo_b200.load_cascade("facefinder"))
    sample = synth.sample_gray()
    counts = []

    variants = [{}, {"tile_warps": 4, "tile_ks": 5, "gather_ks": 7, "tile_tail_min": 33},
                {"scan_mode": 3, "gather_ks": 4}, {"scan_mode": 1}, {"tile_ni": 2, "tile_warps": 8, "gather_warps": 0, "tile_prefetch": 1}]
    if args.quick:
        variants = variants[:2]
    keys = sorted({k for v in variants for k in v})
    saved = {k: pigo_b200.get_option(k) for k in keys}
    frames = np.stack([synth.frame_faces(sample, 360, 640, shift=(20 * i, 10 * i), noise_seed=3 + i) for i in range(2)])
    odd = np.zeros((203, 331), dtype=np.uint8)
    odd[:, :320] = synth.frame_faces(None, 203, 320, shift=(3, 1), noise_seed=8)
    for v in variants:
        for k, val in saved.items():
            pigo_b200.set_option(k, val)
        for k, val in v.items():
            pigo_b200.set_option(k, val)
        counts.append(len(clf.run_cascade_array(cp_of(sample, 400, 320, 320), 0.0)))
        counts.append(len(clf.run_cascade_array(cp_of(odd, 203, 320, 331), 0.0)))       # Dim % 16 != 0: byte-wise tile fill
        dets, cnt = clf.RunCascadeBatch(frames, cp_of(None, 360, 640, 640), 0.0, cap_per_frame=256)
        counts += [int(c) for c in cnt]
    for k, val in saved.items():
        pigo_b200.set_option(k, val)
    for a in (0.3, 0.97):                                                                # rotated: universal gather kernel
        counts.append(len(clf.run_cascade_array(cp_of(sample, 400, 320, 320), a)))

    if not args.quick:
        dets = clf.RunCascade(cp_of(sample, 400, 320, 320), 0.0)
        counts.append(len(clf.ClusterDetections(dets, 0.1)))
        plc = pigo_b200.NewPuplocCascade().UnpackCascade(pigo_b200.load_cascade("puploc"))
        names = sorted(set(pipeline.EYE_CASCADES + pipeline.MOUTH_CASCADES))
        flp = {n: pigo_b200.NewPuplocCascade().UnpackCascade(pigo_b200.load_cascade("lps/" + n)) for n in names}
        faces = pipeline.detect_batch(clf, plc, flp, frames, cp_of(None, 360, 640, 640), iou=0.1)
        counts.append(sum(len(f) for f in faces))
        counts.append(sum(1 for f in faces for fc in f if fc.landmarks))
        rgba = np.random.default_rng(0).integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
        counts.append(int(pigo_b200.RgbToGrayscale(rgba).sum()))
    print("sanitize_run counts:", counts, "launches:", pigo_b200.launch_count() if hasattr(pigo_b200, "launch_count") else "n/a")


if __name__ == "__main__":
    main()
