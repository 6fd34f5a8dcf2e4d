#!/bin/bash
python tools/quickbench.py --frames 256 --reps 3 --host --opts "stream_taper=0/stream_taper=2/stream_taper=0/stream_taper=2,copy_chunk=4/stream_taper=0,copy_chunk=8" 2>&1 | tee gpurun_out/sweep_e2e2.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_b3.json 2> gpurun_out/r02_b3.err; python -c "
import json; d=json.load(open('gpurun_out/r02_b3.json')); print(d['value'], d['ms_per_step'], d['e2e']); print(d['config5']['value'], d['config5']['kernel_ms_rank0'], d['config5']['e2e']); print(d['config4']['rotated_min'], d['config4']['rotated_median']); print(d['single_frame']); print(d['roofline']['tile_role_lanes'], d['roofline']['traffic'], d['roofline']['traffic_note'])"; tail -3 gpurun_out/r02_b3.err
