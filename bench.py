#!/usr/bin/env python
"""bench.py -- candidate windows/s of the PICO cascade scan (RunCascade, core/pigo.go:212-258) on 1080p frames.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun, one rank per GPU) prints ONE JSON
line on rank 0.  A "step" = one pass of the hot path (scan of every (scale,row,col) window + emission-order finalize,
then for N>1 the single NCCL gather of the detection slices) over one batch of synthetic frames.

Workload = BASELINE.json configs[2]: batch of 256 x 1920x1080 uint8 frames PER GPU (weak scaling), facefinder cascade,
reference test parameters MinSize 20 / MaxSize 1000 / ShiftFactor 0.2 / ScaleFactor 1.1 (core/pigo_test.go:44-50),
894,448 windows per frame.  configs[1] (one 1080p frame) is a pure latency case (one frame is ~6 us of issue work);
it is measured too and reported under "single_frame", but the metric -- a throughput -- is quoted on the batch.
Inputs (530 MB per GPU) exceed the 126 MB L2, so no L2 flush is needed between iterations.

  value     : windows/s, frames already resident in HBM, timed with CUDA events on the launching stream, max over ranks
  e2e       : same metric through the public host API (pigo_run_cascade_batch with HOST buffers): pinned H2D of the
              frames and D2H of counts+detections inside the timed region
  roofline  : the scan kernel's ALGORITHMIC bytes (frames*rows*dim + 16*n_det + cascade bytes)
              over its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.  This path is issue/latency bound,
              NOT HBM bound (2.3 B/window): the fraction is reported as it is.
  cpu_baseline / --impl reference : the CPU restatement of the reference (oracle/, kind "port" -- the reference is Go
              and no Go toolchain exists here) on the host cores, frame-parallel, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

ROWS, COLS = 1080, 1920
PARAMS = (20, 1000, 0.2, 1.1)
CASCADE_BYTES = 241488  # codes 119,808 + leaves 119,808 + thresholds 1,872


def make_frames(nframes: int, seed0: int) -> np.ndarray:
    """Deterministic content classes U/S/F: 24 base frames, the rest are circular shifts."""
    from pigo_b200 import synth
    nbase = min(nframes, 24)
    base = synth.make_batch(nbase, ROWS, COLS, "USF", seed0=seed0)
    out = np.empty((nframes, ROWS, COLS), dtype=np.uint8)
    for i in range(nframes):
        b = base[i % nbase]
        k = i // nbase
        out[i] = b if k == 0 else np.roll(b, (17 * k, 29 * k), axis=(0, 1))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(frames: np.ndarray, steps: int, warmup: int, nthreads: int):
    """Times the CPU restatement (oracle) frame-parallel on `nthreads` host threads; returns (windows/s, ms/step)."""
    import oracle_lib as O
    import pigo_b200
    face = O.OracleFace(pigo_b200.load_cascade("facefinder"))
    W = O.count_windows(ROWS, COLS, *PARAMS)
    for _ in range(warmup):
        face.run_cascade_batch(frames, ROWS, COLS, COLS, *PARAMS, 0.0, cap_per_frame=2048, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        face.run_cascade_batch(frames, ROWS, COLS, COLS, *PARAMS, 0.0, cap_per_frame=2048, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return W * frames.shape[0] * steps / dt, dt / steps * 1e3


def _claim_stdout():
    """The contract is ONE JSON line on stdout: everything else any library writes to fd 1 (NCCL prints its version
    banner there at WARN level, torchrun children inherit the fd) is sent to stderr; the line itself goes to the saved fd."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(fd, obj):
    sys.stdout.flush()
    os.write(fd, (json.dumps(obj) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="frames in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opts", default="", help="developer sweeps: library options as name=value,... (default: none)")
    args = ap.parse_args()
    out_fd = _claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    config = {"workload": "configs[2]: batch 256 x 1920x1080 synthetic grayscale frames per GPU (classes U/S/F), facefinder, "
                          "MinSize 20 MaxSize 1000 ShiftFactor 0.2 ScaleFactor 1.1, angle 0",
              "frames_per_gpu": args.frames, "windows_per_frame": 894448, "l2": "inputs (530 MB/GPU) larger than L2, no flush",
              "parallelism": f"frame-sharded dp{world}"}

    # ------------------------------------------------------------------ reference arm: the CPU path on host cores
    if args.impl == "reference":
        if rank != 0:
            return
        nthreads = ncores
        sample = args.cpu_sample_frames or max(8, min(64, 2 * nthreads))
        frames = make_frames(sample, 0)
        v, ms = cpu_reference_run(frames, args.steps, max(args.warmup, 1), nthreads)
        _emit(out_fd, {
            "impl": "reference", "metric": "candidate windows/s on 1080p frames", "value": v, "unit": "windows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "windows/s", "cores": nthreads, "kind": "port",
                             "sample": f"{sample} of the workload's 1080p frames per step, frame-parallel on {nthreads} threads "
                                       "(C restatement of core/pigo.go RunCascade; the Go reference cannot be built here)"},
            "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import pigo_b200
    from pigo_b200 import dist as pdist

    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"   # errors only (they land on stderr, see _claim_stdout)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    pigo_b200.init(dev)
    for kv in filter(None, args.opts.split(",")):
        pigo_b200.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    clf = pigo_b200.NewPigo().Unpack(pigo_b200.load_cascade("facefinder"))
    W = pigo_b200.count_windows(ROWS, COLS, *PARAMS)
    nf = args.frames
    cap = 1024

    frames_host = make_frames(nf, seed0=1000 * rank)
    pinned = torch.from_numpy(frames_host).pin_memory()
    d_frames = pinned.to(f"cuda:{dev}", non_blocking=False)
    d_out = torch.zeros((nf, cap, 4), dtype=torch.int32, device=f"cuda:{dev}")
    d_cnt = torch.zeros(nf, dtype=torch.int32, device=f"cuda:{dev}")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream

    def step_device():
        clf.run_cascade_batch_device(d_frames.data_ptr(), nf, ROWS * COLS, ROWS, COLS, COLS, *PARAMS, 0.0,
                                     d_out.data_ptr(), cap, d_cnt.data_ptr(), st)
        if world > 1:
            pdist.gather_detections(d_out, d_cnt, dst=0)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value")
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()   # nvidia-smi samples every 100 ms: started before the warm-up so the short timed region is covered
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = pigo_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = pigo_b200.launch_count() - launches0
    for _ in range(args.steps):   # keep the GPU under the same load while the sampler collects a few more points
        step_device()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = W * nf * world / (ms_step * 1e-3)
    ndet = int(d_cnt.clamp(max=cap).sum().item())

    # ---- roofline pass: per-kernel CUDA events inside the library (separate pass so the events do not perturb `value`)
    pigo_b200.set_option("timing", 1)
    for _ in range(args.steps):
        step_device()
    torch.cuda.synchronize()
    kt = {}
    for name in ("tiled", "gather", "deep", "finalize"):
        n = pigo_b200.get_option(f"t_{name}_n")
        ns = pigo_b200.get_option(f"t_{name}_ns")
        if n > 0:
            kt[name] = {"launches": int(n), "avg_us": ns / n / 1e3, "total_ms": ns / 1e6}
    pigo_b200.set_option("timing", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    dom = max(kt.items(), key=lambda kv: kv[1]["total_ms"]) if kt else None
    roofline = None
    if dom:
        # One launch of the dominant kernel (the fused scan kernel, "tiled") covers one pipeline group of `sub_batch`
        # frames: algorithmic bytes per launch = its frames once + its share of the detections + the cascade once.
        launches_per_step = dom[1]["launches"] / args.steps
        frames_per_launch = nf / launches_per_step
        alg_bytes = frames_per_launch * ROWS * COLS + 16 * ndet / launches_per_step + CASCADE_BYTES
        scan_ms = sum(v["total_ms"] for k, v in kt.items() if k != "finalize") / args.steps
        launch_ms = dom[1]["avg_us"] / 1e3
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
            if tr.get("kernel") == dom[0]:
                traffic = tr["dram_bytes_per_launch"] * frames_per_launch / tr["frames_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": dom[0], "kernel_launch_ms": launch_ms, "frames_per_launch": frames_per_launch,
                    "scan_kernels_ms_per_step": scan_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                    "traffic_note": "ncu dram bytes of one captured launch (profiles/traffic_r01.json), scaled by frames per launch",
                    "kernels": kt, "note": "path is issue/latency/L2-transaction bound (2.3 algorithmic B/window), not HBM "
                    "bound; see DESIGN.md"}

    # ---- end to end through the host API: pinned H2D of the frames + D2H of counts and detections every step
    out_h = np.zeros((nf, cap), dtype=pigo_b200.DET_DTYPE)
    cnt_h = np.zeros(nf, dtype=np.int32)
    L = pigo_b200.lib()

    def step_host():
        rc = L.pigo_run_cascade_batch(clf._h, pinned.data_ptr(), nf, ROWS * COLS, ROWS, COLS, COLS, PARAMS[0], PARAMS[1],
                                      PARAMS[2], PARAMS[3], 0.0, out_h.ctypes.data, cap, cnt_h.ctypes.data, 0, None)
        if rc != 0:
            raise RuntimeError(L.pigo_last_error().decode())
    e2e_steps = max(2, min(args.steps, 5))
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_value = W * nf * world / (e2e_ms * 1e-3)
    d2h = nf * 4 + int(min(int(cnt_h.max()), cap)) * 16 * nf

    # ---- configs[1]: single 1080p frame (latency case), device resident
    single = None
    if rank == 0:
        one_out = torch.zeros((1, cap, 4), dtype=torch.int32, device=f"cuda:{dev}")
        one_cnt = torch.zeros(1, dtype=torch.int32, device=f"cuda:{dev}")
        lat = {}
        for label, idx in (("U", 0), ("S", 1), ("F", 2)):
            fr = d_frames[idx:idx + 1]
            ts = []
            for it in range(13):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                clf.run_cascade_batch_device(fr.data_ptr(), 1, ROWS * COLS, ROWS, COLS, COLS, *PARAMS, 0.0,
                                             one_out.data_ptr(), cap, one_cnt.data_ptr(), st)
                b.record(stream)
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(a.elapsed_time(b))
            lat[label] = float(np.median(ts))
        med = float(np.mean(list(lat.values())))
        single = {"workload": "configs[1]: one 1920x1080 frame, device resident", "ms_per_frame_by_class": lat,
                  "windows_per_s": W / (med * 1e-3)}

    # ---- CPU baseline (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample_frames or max(8, min(64, 2 * ncores))
        v, ms = cpu_reference_run(frames_host[:sample], 2, 1, ncores)
        v1, ms1 = cpu_reference_run(frames_host[:4], 1, 0, 1)
        cpu_baseline = {"value": v, "unit": "windows/s", "cores": ncores, "kind": "port",
                        "sample": f"first {sample} frames of this workload x 2 passes, frame-parallel on {ncores} threads",
                        "single_thread_value": v1, "note": "C restatement of core/pigo.go (oracle/); the Go reference "
                        "cannot be built in this image (no Go toolchain)"}

    if rank == 0:
        _emit(out_fd, {
            "metric": "candidate windows/s on 1080p frames", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": nf * ROWS * COLS, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "single_frame": single,
            "detections_per_step": ndet})
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
