#!/bin/bash
python tools/quickbench.py --frames 256 --reps 3 --host --opts "stream_taper=1,copy_chunk=8/stream_taper=0/stream_taper=1,copy_chunk=2/stream_taper=1,copy_chunk=32/host_stream=0,sub_batch=64/host_stream=1,sub_batch=0,stream_taper=1,copy_chunk=8" 2>&1 | tee gpurun_out/sweep_e2e.txt
python tools/quickbench.py --frames 1 --reps 20 --opts "gather_limit=0/gather_limit=1000/gather_limit=4/gather_limit=16/gather_limit=8,deep_group=32" 2>&1 | tee gpurun_out/sweep_single2.txt
python tools/quickbench.py --frames 1 --classes F --reps 20 --opts "gather_limit=0/gather_limit=1000" 2>&1 | tee -a gpurun_out/sweep_single2.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -k "sample_image or 1080p or empty_and or strided or concurrent or device_resident or synthetic" 2>&1 | tail -4
