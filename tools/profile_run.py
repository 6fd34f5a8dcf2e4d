#!/usr/bin/env python
"""Small, fixed workloads for ncu captures of every kernel at the SHIPPING (default) configuration.

    ncu --set full --clock-control none --import-source on -k regex:'scan_|deep_' -s <warm-up launches> -c <n> \
        -o gpurun_out/prof_scan python tools/profile_run.py scan

  scan : one 128-frame group of 1080p U/S/F frames, device resident (what one fused launch of bench.py covers), 2 calls
  rot  : 4 x 4K frames, rotated scan at slot 7/32, 2 calls
  pipe : the config-5 pipeline (face -> cluster -> pupils -> 15 landmarks) on 16 class-F 1080p frames, 2 calls,
         then RgbToGrayscale on 16 x 1080p NRGBA frames, 2 calls
Nothing is timed here (numbers printed under a profiler are never bench values)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pigo_b200  # noqa: E402
from pigo_b200 import CascadeParams, ImageParams, pipeline, synth  # noqa: E402

PRM = (20, 1000, 0.2, 1.1)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "scan"
    for kv in filter(None, (sys.argv[2] if len(sys.argv) > 2 else "").split(",")):
        pigo_b200.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    pigo_b200.init(0)
    clf = pigo_b200.NewPigo().Unpack(pigo_b200.load_cascade("facefinder"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    if mode in ("scan", "rot"):
        if mode == "scan":
            R, C, nf, ang = 1080, 1920, 128, 0.0
            base = synth.make_batch(12, R, C, "USF")
            frames = np.concatenate([base] * 11)[:nf]
        else:
            R, C, nf, ang = 2160, 3840, 4, 7 / 32.0
            frames = np.stack([synth.frame_faces(None, R, C, shift=(31 * i, 17 * i), noise_seed=i) for i in range(nf)])
        d = torch.from_numpy(frames).cuda()
        cap = 2048
        d_out = torch.zeros((nf, cap, 4), dtype=torch.int32, device="cuda")
        d_cnt = torch.zeros(nf, dtype=torch.int32, device="cuda")
        for _ in range(2):
            clf.run_cascade_batch_device(d.data_ptr(), nf, R * C, R, C, C, *PRM, ang, d_out.data_ptr(), cap, d_cnt.data_ptr(), st)
            torch.cuda.synchronize()
        print(mode, "detections", int(d_cnt.sum()))
    elif mode == "pipe2":
        # the device-sequenced pipeline (pigo_detect_batch) on 64 class-F 1080p frames, 2 calls
        plc = pigo_b200.NewPuplocCascade().UnpackCascade(pigo_b200.load_cascade("puploc"))
        names = sorted(set(pipeline.EYE_CASCADES + pipeline.MOUTH_CASCADES))
        flp = {n: pigo_b200.NewPuplocCascade().UnpackCascade(pigo_b200.load_cascade("lps/" + n)) for n in names}
        fr = np.stack([synth.frame_faces(None, 1080, 1920, shift=(37 * i, 53 * i), noise_seed=100 + i) for i in range(64)])
        cp = CascadeParams(ImageParams(None, 1080, 1920, 1920), *PRM)
        df = pigo_b200.DeviceFrames(fr)
        for _ in range(2):
            faces, nfaces, points = pipeline.detect_batch_device(clf, plc, flp, df, cp, eye_perturbs=63, raw=True)
        print("pipe2 faces", int((faces["scale"] > 50).sum()))
    elif mode == "gray":
        npx = 16 * 1080 * 1920
        rgba = torch.randint(0, 256, (npx, 4), dtype=torch.uint8, device="cuda")
        gray = torch.empty(npx, dtype=torch.uint8, device="cuda")
        yy = torch.randint(0, 256, (16 * 1080, 1920), dtype=torch.uint8, device="cuda")
        cb = torch.randint(0, 256, (16 * 540, 960), dtype=torch.uint8, device="cuda")
        cr = torch.randint(0, 256, (16 * 540, 960), dtype=torch.uint8, device="cuda")
        out = torch.empty((npx, 4), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            pigo_b200.lib().pigo_rgba_to_gray(rgba.data_ptr(), npx, gray.data_ptr(), 3, st)
            pigo_b200.lib().pigo_ycbcr_to_nrgba(yy.data_ptr(), cb.data_ptr(), cr.data_ptr(), 1920, 960, 2, 0, 0, 1920, 16 * 1080, out.data_ptr(), gray.data_ptr(), 3, st)
            torch.cuda.synchronize()
    elif mode == "pipe":
        plc = pigo_b200.NewPuplocCascade().UnpackCascade(pigo_b200.load_cascade("puploc"))
        names = sorted(set(pipeline.EYE_CASCADES + pipeline.MOUTH_CASCADES))
        flp = {n: pigo_b200.NewPuplocCascade().UnpackCascade(pigo_b200.load_cascade("lps/" + n)) for n in names}
        fr = np.stack([synth.frame_faces(None, 1080, 1920, shift=(37 * i, 53 * i), noise_seed=100 + i) for i in range(16)])
        cp = CascadeParams(ImageParams(None, 1080, 1920, 1920), *PRM)
        for _ in range(2):
            res = pipeline.detect_batch(clf, plc, flp, fr, cp)
        print("pipe faces", sum(1 for f in res for face in f if face.left_eye is not None))
        npx = 16 * 1080 * 1920
        rgba = torch.randint(0, 256, (npx, 4), dtype=torch.uint8, device="cuda")
        gray = torch.empty(npx, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            pigo_b200.lib().pigo_rgba_to_gray(rgba.data_ptr(), npx, gray.data_ptr(), 3, st)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
