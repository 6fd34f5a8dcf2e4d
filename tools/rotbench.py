#!/usr/bin/env python
"""Developer sweep for the rotated scan (configs[3] shape): 8 x 3840x2160 class-F frames, a few table slots, option variants."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pigo_b200  # noqa: E402
from pigo_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opts", default="")
    ap.add_argument("--slots", default="9,25,6,1,20")
    args = ap.parse_args()
    pigo_b200.init(0)
    clf = pigo_b200.NewPigo().Unpack(pigo_b200.load_cascade("facefinder"))
    R, C, nf = 2160, 3840, 8
    frames = np.stack([synth.frame_faces(None, R, C, shift=(31 * i, 17 * i), noise_seed=i) for i in range(nf)])
    d = torch.from_numpy(frames).cuda()
    out = torch.zeros((nf, 4096, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(nf, dtype=torch.int32, device="cuda")
    W = pigo_b200.count_windows(R, C, 20, 1000, 0.2, 1.1)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
    for v in (args.opts.split("/") if args.opts else [""]):
        for kv in filter(None, v.split(",")):
            pigo_b200.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        res = []
        for k in [int(x) for x in args.slots.split(",")]:
            ts = []
            for it in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                clf.run_cascade_batch_device(d.data_ptr(), nf, R * C, R, C, C, 20, 1000, 0.2, 1.1, k / 32.0, out.data_ptr(), 4096, cnt.data_ptr(), st)
                b.record(stream); torch.cuda.synchronize()
                if it >= 2:
                    ts.append(a.elapsed_time(b))
            res.append(f"k={k}: {nf * W / np.median(ts) / 1e6:.2f}")
        print(f"[{v or 'default'}] Gwin/s  " + "  ".join(res), flush=True)


if __name__ == "__main__":
    main()
