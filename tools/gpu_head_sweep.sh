#!/bin/bash
# dense-head kernel: parity subset, then timing sweep against the classic kernel
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -k "scan_variants or sample_image or 1080p or batch_matches or synthetic_cascades or streamed" 2>&1 | tail -8
python tools/quickbench.py --frames 256 --reps 5 --host --opts "tile_head=0,tile_warps=24/tile_head=2,tile_warps=22/tile_head=2,tile_warps=24/tile_head=1,tile_warps=22/tile_head=3,tile_warps=22/tile_head=2,tile_warps=22,gather_warps=10/tile_head=0,tile_warps=24,gather_warps=8,stream_taper=0/stream_taper=1,copy_chunk=4" 2>&1 | tee gpurun_out/sweep_head2.txt
python tools/quickbench.py --frames 1 --reps 20 --opts "tile_core_cap=0/tile_core_cap=512/tile_core_cap=32/tile_core_cap=16,deep_group=32/tile_core_cap=16,gather_block=8" 2>&1 | tee gpurun_out/sweep_single.txt
