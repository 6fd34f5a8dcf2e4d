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
    ap.add_argument("--new", action="store_true", help="only the kernels added late in round 2 (offset tables, shared-memory deep, queue overflow)")
    args = ap.parse_args()
    pigo_b200.init(0)
    clf = pigo_b200.NewPigo().Unpack(pigo_b200.load_cascade("facefinder"))
    sample = synth.sample_gray()
    counts = []

    variants = [{}, {"tile_warps": 4, "tile_ks": 5, "gather_ks": 7, "tile_tail_min": 33},
                {"scan_mode": 3, "gather_ks": 4}, {"scan_mode": 1}, {"tile_ni": 2, "tile_warps": 8, "gather_warps": 0, "tile_prefetch": 1},
                {"tile_head": 2, "tile_warps": 22}, {"tile_head": 1, "tile_warps": 6, "tile_ks": 5, "tile_tail_min": 33},
                {"host_stream": 1, "copy_chunk": 1, "sub_batch": 1, "deep_flat": 1}, {"host_stream": 0}, {"walk_stats": 1}]
    late = [{"tile_ptab": 1}, {"tile_ptab": 1, "ptab_kt": 4, "ptab_ks": 6, "tile_warps": 7, "gather_warps": 3, "tile_tail_min": 33},
            {"deep_smem": 1, "deep_smem_k": 468, "deep_smem_lo": 0}, {"deep_smem": 1, "deep_group": 4, "gather_limit": 4, "tile_ks": 6, "deep_smem_k": 30, "deep_smem_lo": 11},
            {"queue_cap": 9, "tile_ks": 4, "gather_ks": 6, "tile_tail_min": 33}, {"queue_cap": 7, "tile_ptab": 1, "ptab_kt": 3, "ptab_ks": 5, "tile_tail_min": 33}]
    variants = late if args.new else variants + late
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
    for a in (0.3, 0.97):                                                                # rotated: table-driven block + deep kernels
        counts.append(len(clf.run_cascade_array(cp_of(sample, 400, 320, 320), a)))
        dets, cnt = clf.RunCascadeBatch(frames, cp_of(None, 360, 640, 640), a, cap_per_frame=256)
        counts += [int(c) for c in cnt]
    pigo_b200.set_option("rot_mode", 1)                                                  # rotated: universal gather kernel
    counts.append(len(clf.run_cascade_array(cp_of(sample, 400, 320, 320), 0.3)))
    pigo_b200.set_option("rot_mode", 0)

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
        # device-sequenced pipeline (seed kernels + pair kernel, staged and unstaged, rotated eyes), sharded entry point
        for stage, ang in ((1, 0.0), (0, 0.0), (1, 0.05)):
            pigo_b200.set_option("puploc_stage", stage)
            f, n, p = pipeline.detect_batch_device(clf, plc, flp, frames, cp_of(None, 360, 640, 640), face_cap=16, rng_seed=3, angle=ang, raw=True)
            counts += [int(n.sum()), int((p["row"] > 0).sum())]
        pigo_b200.set_option("puploc_stage", 1)
        pigo_b200.init_devices(1)
        f, n, p = pipeline.detect_batch_device(clf, plc, flp, frames, cp_of(None, 360, 640, 640), face_cap=16, rng_seed=3, raw=True, sharded=True)
        counts.append(int(n.sum()))
        pigo_b200.set_option("puploc_mode", 1)
        faces2 = pipeline.detect_batch(clf, plc, flp, frames[:1], cp_of(None, 360, 640, 640), iou=0.1)
        counts.append(sum(len(x) for x in faces2))
        pigo_b200.set_option("puploc_mode", 0)
        yy = np.random.default_rng(1).integers(0, 256, size=(21, 40), dtype=np.uint8)
        cb = np.random.default_rng(2).integers(0, 256, size=(11, 20), dtype=np.uint8)
        cr = np.random.default_rng(3).integers(0, 256, size=(11, 20), dtype=np.uint8)
        out, gy = pigo_b200.YCbCrToNRGBA(yy, cb, cr, 2, 37, 21, want_gray=True)
        counts += [int(out.sum()), int(gy.sum())]
    print("sanitize_run counts:", counts, "launches:", pigo_b200.launch_count() if hasattr(pigo_b200, "launch_count") else "n/a")


if __name__ == "__main__":
    main()
