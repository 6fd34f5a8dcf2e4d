#!/usr/bin/env python
"""Developer micro-benchmark (not the contract bench): times the device-resident batch scan with CUDA events."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pigo_b200  # noqa: E402
from pigo_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--classes", default="USF")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--opts", default="", help="comma list name=value;... variants separated by '/'")
    ap.add_argument("--rows", type=int, default=1080)
    ap.add_argument("--cols", type=int, default=1920)
    ap.add_argument("--shift", type=float, default=0.2)
    ap.add_argument("--stats", action="store_true", help="also report the tile role's live lanes per walk iteration (walk_stats)")
    ap.add_argument("--host", action="store_true", help="also time the host API (pinned frames in, detections out) per variant")
    args = ap.parse_args()
    pigo_b200.init(0)
    clf = pigo_b200.NewPigo().Unpack(pigo_b200.load_cascade("facefinder"))
    base = synth.make_batch(min(args.frames, 12), args.rows, args.cols, args.classes)
    reps = -(-args.frames // len(base))
    frames = np.concatenate([base] * reps)[:args.frames]
    d_frames = torch.from_numpy(frames).cuda()
    cap = 2048
    d_out = torch.zeros((args.frames, cap, 4), dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(args.frames, dtype=torch.int32, device="cuda")
    W = pigo_b200.count_windows(args.rows, args.cols, 20, 1000, args.shift, 1.1)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    variants = [v for v in args.opts.split("/")] if args.opts else [""]
    for v in variants:
        for kv in filter(None, v.split(",")):
            k, val = kv.split("=")
            pigo_b200.set_option(k, int(val))

        def run():
            clf.run_cascade_batch_device(d_frames.data_ptr(), args.frames, args.rows * args.cols, args.rows, args.cols, args.cols,
                                         20, 1000, args.shift, 1.1, 0.0, d_out.data_ptr(), cap, d_cnt.data_ptr(), st)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        pigo_b200.set_option("timing", 1)
        run(); torch.cuda.synchronize()
        kt = {n: pigo_b200.get_option(f"t_{n}_ns") / 1e6 for n in ("tiled", "gather", "deep", "finalize")}
        pigo_b200.set_option("timing", 0)
        if args.stats:
            pigo_b200.get_option("walk_iters")
            pigo_b200.set_option("walk_stats", 1)
            run(); torch.cuda.synchronize()
            wu0, wi0 = pigo_b200.get_option("walk_useful"), pigo_b200.get_option("walk_iters")
            run(); torch.cuda.synchronize()
            wu, wi = pigo_b200.get_option("walk_useful") - wu0, pigo_b200.get_option("walk_iters") - wi0
            pigo_b200.set_option("walk_stats", 0)
            print(f"    tile role: {wi / 1e6:.2f} M walk iterations, {wu / 1e6:.1f} M useful tree walks, {wu / max(wi, 1):.2f} live lanes per iteration", flush=True)
        host_ms = None
        if args.host:
            import ctypes as C
            if not hasattr(main, "_pin"):
                main._pin = torch.from_numpy(frames).pin_memory()
                main._oh = np.zeros((args.frames, cap), dtype=pigo_b200.DET_DTYPE)
                main._ch = np.zeros(args.frames, dtype=np.int32)
            L = pigo_b200.lib()

            def run_host():
                rc = L.pigo_run_cascade_batch(clf._h, main._pin.data_ptr(), args.frames, args.rows * args.cols, args.rows, args.cols, args.cols, 20, 1000,
                                              args.shift, 1.1, 0.0, main._oh.ctypes.data, cap, main._ch.ctypes.data, 0, None)
                assert rc == 0, L.pigo_last_error()
            run_host(); run_host()
            if not hasattr(main, "_h2d"):
                dst = torch.empty_like(d_frames)
                cs = []
                for _ in range(5):
                    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(main._pin, non_blocking=True); torch.cuda.synchronize()
                    cs.append((time.perf_counter() - t0) * 1e3)
                main._h2d = float(np.median(cs))
                print(f"    pure pinned H2D of the batch: {main._h2d:.3f} ms ({frames.nbytes / main._h2d / 1e6:.1f} GB/s)", flush=True)
                del dst
            hs = []
            for _ in range(max(args.reps, 12)):
                t0 = time.perf_counter(); run_host(); hs.append((time.perf_counter() - t0) * 1e3)
            host_ms = float(np.median(hs))
            assert int(main._ch.sum()) == int(d_cnt.sum()), "host path detections differ"
        if host_ms is not None:
            print(f"    host API: {host_ms:.3f} ms/step (min {min(hs):.3f})  {args.frames * W / host_ms / 1e6:.2f} Gwin/s", flush=True)
        print(f"[{v or 'default'}] frames={args.frames} {args.rows}x{args.cols} shift={args.shift}: {ms:.3f} ms/step  "
              f"{args.frames * W / ms / 1e6:.2f} Gwin/s  dets={int(d_cnt.sum())} (min {min(ts):.3f} max {max(ts):.3f}) kernels ms: "
              + " ".join(f"{k}={v:.3f}" for k, v in kt.items()), flush=True)


if __name__ == "__main__":
    main()
