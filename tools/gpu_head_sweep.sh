#!/bin/bash
# dense-head kernel: parity subset, then timing sweep against the classic kernel
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -k "scan_variants or sample_image or 1080p or batch_matches or synthetic_cascades or streamed" 2>&1 | tail -8
python tools/quickbench.py --frames 256 --reps 5 --opts "tile_head=0,tile_warps=24/tile_head=2,tile_warps=22,head_back=12/tile_head=2,tile_warps=24/tile_head=1,tile_warps=22/tile_head=3,tile_warps=22/tile_head=2,tile_warps=22,head_back=6/tile_head=2,tile_warps=22,head_back=16/tile_head=2,tile_warps=23,gather_warps=9/tile_head=2,tile_warps=22,gather_warps=10,tile_ks=40/tile_head=0,tile_warps=24,gather_warps=8,tile_ks=48" 2>&1 | tee gpurun_out/sweep_head.txt
