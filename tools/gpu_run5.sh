python tools/quickbench.py --frames 256 --reps 5 --host --opts "deep_flat=0/deep_flat=1/deep_flat=0,tile_ks=40/tile_ks=48,gather_ks=48/gather_ks=32,deep_group=16/deep_group=8" 2>&1 | tee gpurun_out/sweep3.txt
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'puploc_pair|seed_kernel' -s 2 -c 4 -o gpurun_out/r02b_pipe -f python tools/profile_run.py pipe2 > gpurun_out/r02b_pipe.log 2>&1
$NCU -k regex:'scan_gather2|deep_kernel|rot_table' -s 3 -c 3 -o gpurun_out/r02b_rot -f python tools/profile_run.py rot > gpurun_out/r02b_rot.log 2>&1
tail -3 gpurun_out/r02b_pipe.log gpurun_out/r02b_rot.log
