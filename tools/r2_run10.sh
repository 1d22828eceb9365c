#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -x -q --timeout 300 -k "single_stage or batch8 or stream" > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2o_pytest.log
run() { # name, env..., --, args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e "$@" > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2o_$name.json').read().strip().splitlines()[-1])
    ops=d.get('ops',[])
    print('value %.0f ms/step %.3f sum_hot %.0f stem %.1f op5 %.1f op8 %.1f op45 %.1f' % (d['value'], d['ms_per_step'], sum(o['us_hot'] for o in ops), ops[0]['us_hot'], ops[5]['us_hot'], ops[8]['us_hot'], ops[45]['us_hot']))
except Exception as e:
    print('no line', e)
PY
)"; tail -n 2 gpurun_out/r2o_$name.err
}
run base X=1 --
run cluster DEFER_UMMA_CLUSTER=1 DEFER_UMMA_CSPLIT_MAX_CTAS=160 --
run light_u5s1 DEFER_STREAM_LIGHT_UNITS=5 DEFER_STREAM_LIGHT_STAGES=1 --
run light_u4 DEFER_STREAM_LIGHT_UNITS=4 --
run light_u2 DEFER_STREAM_LIGHT_UNITS=2 --
run kheavy3 DEFER_STREAM_KHEAVY=3 --
run kheavy12 DEFER_STREAM_KHEAVY=12 --
run g32 X=1 -- --coalesce 32
