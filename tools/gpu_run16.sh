#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "scan_variants or sample_image or 1080p or batch_matches or synthetic_cascades" 2>&1 | tail -4
python tools/quickbench.py --frames 256 --reps 7 --opts "tile_head=0,tile_warps=24/tile_head=1,tile_warps=22/tile_head=0,tile_warps=24,tile_prefetch=1/tile_prefetch=0" 2>&1 | tee gpurun_out/sweep_lds.txt
