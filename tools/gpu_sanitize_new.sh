#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (tools/sanitize_run.py --new)
O=gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck initcheck racecheck; do
  timeout 700 $S --tool $tool --error-exitcode 3 python tools/sanitize_run.py --new > $O/${tool}_r02new.txt 2>&1; echo "$tool rc=$?" | tee -a $O/${tool}_r02new.txt
done
for f in memcheck synccheck initcheck racecheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run counts|rc=" $O/${f}_r02new.txt | tail -4; done
