#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 400 --warmup 20 --no-cpu --no-e2e --no-roofline --depth 32"
run() { local name=$1; shift
  env "$@" timeout 40 $B > gpurun_out/c5_$name.json 2> gpurun_out/c5_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/c5_$name.json').read().strip().splitlines()[-1]); print('$name value', round(d['value'],1))
except Exception as e: print('$name FAILED', e)
"
}
run d32 DEFER_X=1
run d32_ew8 DEFER_UMMA_EPI_WARPS=8
run d32_st2 DEFER_UMMA_STAGES=2
