#!/bin/bash
# round 2, GPU call 3 (1 GPU): conv_stream_kernel parity, then throughput variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 300 -k "stream" > gpurun_out/r2d_pytest_stream.log 2>&1
echo "pytest(stream) rc=$?"; tail -n 6 gpurun_out/r2d_pytest_stream.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest(all) rc=$?"; tail -n 8 gpurun_out/r2d_pytest.log
run() { # name, env..., --, args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e "$@" > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2d_$name.json').read().strip().splitlines()[-1])
    ops=d.get('ops',[])
    print('value %.0f ms/step %.3f sum_hot %.0f sum_cold %.0f frac %.3f' % (d['value'], d['ms_per_step'], sum(o['us_hot'] for o in ops), sum(o['us_cold'] for o in ops), d.get('roofline',{}).get('frac',0)))
except Exception as e:
    print('no line', e)
PY
)"
}
run base X=1 --
run old DEFER_STREAM=0 --
run min32 DEFER_STREAM_MIN_TILES=32 --
run min192 DEFER_STREAM_MIN_TILES=192 --
run bn64 DEFER_STREAM_BN=64 --
run bn128all DEFER_STREAM_BN128_TILES=1 --
run u2 DEFER_STREAM_UNITS=2 --
run kheavy3 DEFER_STREAM_KHEAVY=3 --
run g32 X=1 -- --coalesce 32
run g8 X=1 -- --coalesce 8
run g16d2 X=1 -- --depth 2
run bf16 X=1 -- --dtype bfloat16
