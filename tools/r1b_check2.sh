#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x --timeout 150 -k "dense or epilogue_variants or resnet50_single_stage or bitwise or vgg16_single" > gpurun_out/c4_pytest.log 2>&1; tail -n 5 gpurun_out/c4_pytest.log
grep -q " passed" gpurun_out/c4_pytest.log && ! grep -q "failed\|error" gpurun_out/c4_pytest.log || { grep -E "Error|error|assert|rel=" gpurun_out/c4_pytest.log | head -30; exit 1; }
B="python bench.py --steps 400 --warmup 20 --no-cpu --no-e2e --no-roofline"
run() { local name=$1; shift
  env "$@" timeout 120 $B > gpurun_out/c4_$name.json 2> gpurun_out/c4_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/c4_{n}.json").read().strip().splitlines()[-1]); print(n, "value", round(d["value"],1))
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/c4_{n}.err").read()[-600:])
PY
}
run all
run nostem DEFER_TC_STEM=0
run nodense DEFER_DENSE_FUSED=0
run all_d32 DEFER_X=1
timeout 120 $B --depth 32 > gpurun_out/c4_d32.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/c4_d32.json').read().strip().splitlines()[-1]); print('depth32', round(d['value'],1))"
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e --batched-roofline 0 > gpurun_out/c4_ops.json 2> gpurun_out/c4_ops.err
