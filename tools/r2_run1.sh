#!/bin/bash
# round 2, GPU call 1 (1 GPU): full GPU test suite, then a coalesce x depth sweep of the new bench protocol
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -n 12 gpurun_out/r2a_pytest.log
run() { # name, args...
  name=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-roofline "$@" > gpurun_out/r2a_$name.json 2> gpurun_out/r2a_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2a_$name.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f parity %s' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d.get('parity_rel_err')))
except Exception as e:
    print('no line', e)
PY
)"
}
run g1_d32 --coalesce 1 --depth 32
run g4_d8 --coalesce 4 --depth 8
run g8_d2 --coalesce 8 --depth 2
run g8_d4 --coalesce 8 --depth 4
run g16_d2 --coalesce 16 --depth 2
run g16_d4 --coalesce 16 --depth 4
run g32_d2 --coalesce 32 --depth 2
run g32_d4 --coalesce 32 --depth 4
run g64_d2 --coalesce 64 --depth 2
# long-run agreement check of the protocol (400 steps vs 20 steps)
run g16_d4_long --coalesce 16 --depth 4 --steps 400 --warmup 20
# full default line (roofline + cpu baseline)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_default.json 2> gpurun_out/r2a_default.err
echo "default rc=$?"; head -c 600 gpurun_out/r2a_default.json
