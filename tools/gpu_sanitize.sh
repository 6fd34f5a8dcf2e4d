#!/bin/bash
O=gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 3 python tools/sanitize_run.py > $O/memcheck_r02.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/memcheck_r02.txt
timeout 900 $S --tool synccheck --error-exitcode 3 python tools/sanitize_run.py > $O/synccheck_r02.txt 2>&1; echo "synccheck rc=$?" | tee -a $O/synccheck_r02.txt
timeout 900 $S --tool initcheck --error-exitcode 3 python tools/sanitize_run.py > $O/initcheck_r02.txt 2>&1; echo "initcheck rc=$?" | tee -a $O/initcheck_r02.txt
timeout 1500 $S --tool racecheck --error-exitcode 3 python tools/sanitize_run.py > $O/racecheck_r02.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/racecheck_r02.txt
for f in memcheck synccheck initcheck racecheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run counts|rc=" $O/${f}_r02.txt | tail -4; done
