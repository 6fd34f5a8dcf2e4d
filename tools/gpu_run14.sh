#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "puploc or pipeline or landmark" 2>&1 | tail -6
for st in 1 2; do
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opts puploc_stage=$st > gpurun_out/r02_pl$st.json 2> gpurun_out/r02_pl$st.err
python -c "
import json; d=json.load(open('gpurun_out/r02_pl$st.json')); print('stage $st', d['config5']['value'], d['config5']['kernel_ms_rank0'], d['config5']['e2e']['value']); print(d['single_frame'])"
done
