#!/bin/bash
# round 2, GPU call 5 (1 GPU): two-instruction MMA scheme + new im2col: full parity suite, trace, bench variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r2i_pytest.log 2>&1
echo "pytest(all) rc=$?"; tail -n 8 gpurun_out/r2i_pytest.log
rm -f gpurun_out/r2i_trace.txt
export DEFER_UMMA_TRACE=gpurun_out/r2i_trace.txt
python tools/run_one_conv.py bf16x2 4 16 56 56 64 64 3 1 1 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 5 16 28 28 128 128 3 1 1 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 5 16 14 14 1024 256 1 1 0 1 > /dev/null 2>&1
unset DEFER_UMMA_TRACE
cat gpurun_out/r2i_trace.txt | cut -c1-400
run() { # name, env..., --, args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e "$@" > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2i_$name.json').read().strip().splitlines()[-1])
    ops=d.get('ops',[])
    print('value %.0f ms/step %.3f sum_hot %.0f sum_cold %.0f frac %.3f stem %.1f' % (d['value'], d['ms_per_step'], sum(o['us_hot'] for o in ops), sum(o['us_cold'] for o in ops), d.get('roofline',{}).get('frac',0), ops[0]['us_hot'] if ops else 0))
except Exception as e:
    print('no line', e)
PY
)"
}
run base X=1 --
run min192 DEFER_STREAM_MIN_TILES=192 --
run min32 DEFER_STREAM_MIN_TILES=32 --
run g32 X=1 -- --coalesce 32
run g32_d2 X=1 -- --coalesce 32 --depth 2
run g8 X=1 -- --coalesce 8
run bf16 X=1 -- --dtype bfloat16
run bf16_g32 X=1 -- --dtype bfloat16 --coalesce 32
run g1 X=1 -- --coalesce 1 --depth 20
