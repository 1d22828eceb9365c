#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 400 --warmup 20 --no-cpu --no-e2e --no-roofline"
for d in 4 8 32 64; do
  timeout 120 $B --depth $d > gpurun_out/c3_d$d.json 2> gpurun_out/c3_d$d.err
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/c3_d$d.json').read().strip().splitlines()[-1]); print('depth $d value', round(d['value'],1))
except Exception as e: print('depth $d FAILED', e, open('gpurun_out/c3_d$d.err').read()[-300:])
"
done
DEFER_TIMELINE=/tmp/tl_te.txt timeout 120 $B > gpurun_out/c3_tl_te.json 2> gpurun_out/c3_tl_te.err
python tools/timeline_stats.py /tmp/tl_te.txt > gpurun_out/c3_tl_te_stats.txt 2>&1
head -7 gpurun_out/c3_tl_te_stats.txt
