#!/bin/bash
# round 2, GPU call 2 (1 GPU): split-K policy variants of the per-op kernel under the coalesced workload
mkdir -p gpurun_out
run() { # name, env..., --, args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e "$@" > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2c_$name.json').read().strip().splitlines()[-1])
    ops=d.get('ops',[])
    print('value %.0f ms/step %.3f sum_hot %.0f' % (d['value'], d['ms_per_step'], sum(o['us_hot'] for o in ops)))
except Exception as e:
    print('no line', e)
PY
)"
}
run base X=1 --
run t64 DEFER_UMMA_TARGET_CTAS=64 --
run t128 DEFER_UMMA_TARGET_CTAS=128 --
run t148_cl DEFER_UMMA_CLUSTER=1 DEFER_UMMA_CSPLIT_MAX_CTAS=160 --
run nopersist DEFER_PERSIST_MIN_TILES=0 --
run nopersist_t128 DEFER_PERSIST_MIN_TILES=0 DEFER_UMMA_TARGET_CTAS=128 --
run persist_all DEFER_PERSIST_MIN_TILES=1 --
run g32_t128 DEFER_UMMA_TARGET_CTAS=128 -- --coalesce 32
run g8_t128 DEFER_UMMA_TARGET_CTAS=128 -- --coalesce 8
run bf16 X=1 -- --dtype bfloat16
run bf16_t128 DEFER_UMMA_TARGET_CTAS=128 -- --dtype bfloat16
