#!/bin/bash
# Runs under gpurun: ncu captures of every kernel at the default configuration + launch list of the bench command.
# usage: tools/profile_gpu.sh <tag>      (outputs land in gpurun_out/<tag>_*)
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
# one 128-frame group of the bench workload: fused, gather-v2, deep (second call = warm)
$NCU -k regex:'scan_|deep_' -s 3 -c 3 -o $O/${TAG}_scan -f python tools/profile_run.py scan > $O/${TAG}_scan.log 2>&1
# 4K rotated (table-driven block kernel + deep kernel + node table)
$NCU -k regex:'scan_|deep_|rot_table' -s 3 -c 3 -o $O/${TAG}_rot -f python tools/profile_run.py rot > $O/${TAG}_rot.log 2>&1
# device pipeline on 64 frames: finalize, cluster, eye seeds, pupil kernel (eyes), landmark seeds, pupil kernel (landmarks)
$NCU -k regex:'finalize|cluster|seed_kernel|puploc' -s 6 -c 6 -o $O/${TAG}_pipe -f python tools/profile_run.py pipe2 > $O/${TAG}_pipe.log 2>&1
$NCU -k regex:'gray|ycbcr' -s 1 -c 2 -o $O/${TAG}_gray -f python tools/profile_run.py gray > $O/${TAG}_gray.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 160 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1
echo profile done
