#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "scan_variants or sample_image or 1080p or batch_matches or synthetic_cascades or streamed or strided or empty_and or concurrent or device_resident" 2>&1 | tail -6
python tools/quickbench.py --frames 256 --reps 5 --opts "tile_tmap=1/tile_tmap=0/tile_tmap=1,tile_warps=25,gather_warps=7/tile_warps=24,gather_warps=8,tile_head=1,tile_warps=22/tile_head=0,tile_warps=24" 2>&1 | tee gpurun_out/sweep_tmap.txt
python tools/quickbench.py --frames 1 --reps 20 --opts "tile_tmap=1/tile_tmap=0" 2>&1 | tee -a gpurun_out/sweep_tmap.txt
