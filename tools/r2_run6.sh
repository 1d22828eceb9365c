#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q --timeout 300 -k "single_stage or batch4 or batch8 or pipeline_same_gpu or coalesced" > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/r2k_pytest.log
run() { # name, env..., --, args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e "$@" > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2k_$name.json').read().strip().splitlines()[-1])
    ops=d.get('ops',[])
    print('value %.0f ms/step %.3f sum_hot %.0f stem %s %.1f/%.1f' % (d['value'], d['ms_per_step'], sum(o['us_hot'] for o in ops), ops[0]['kernel'], ops[0]['us_hot'], ops[0]['us_cold']))
except Exception as e:
    print('no line', e)
PY
)"; tail -n 2 gpurun_out/r2k_$name.err
}
run base X=1 --
run nofuse DEFER_STEM_FUSED=0 --
run g32 X=1 -- --coalesce 32
run bf16 X=1 -- --dtype bfloat16
