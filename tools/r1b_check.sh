#!/bin/bash
# cluster split-K bring-up: kernel parity tests, then bench A/B over the split heuristics
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "tcgen05" > gpurun_out/c1_pytest.log 2>&1; tail -n 5 gpurun_out/c1_pytest.log
grep -q " passed" gpurun_out/c1_pytest.log && ! grep -q "failed\|error" gpurun_out/c1_pytest.log || { grep -E "Error|error|assert" gpurun_out/c1_pytest.log | head -20; exit 1; }
B="python bench.py --steps 400 --warmup 20 --no-cpu --no-e2e --no-roofline"
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 120 $B > gpurun_out/c1_$name.json 2> gpurun_out/c1_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/c1_{n}.json").read().strip().splitlines()[-1]); print(n, "value", round(d["value"],1), "probs_sum", d.get("probs_sum"))
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/c1_{n}.err").read()[-600:])
PY
}
run base DEFER_UMMA_TMA_EPI=0
run te
run te_res128 DEFER_UMMA_TE_RES_BN64=0
run te_cl4 DEFER_UMMA_CLUSTER=1 DEFER_UMMA_CSPLIT_KB=4
run te_cl8 DEFER_UMMA_CLUSTER=1 DEFER_UMMA_CSPLIT_KB=8
run te_bn64 DEFER_UMMA_BN=64
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e --batched-roofline 0 > gpurun_out/c1_te_ops.json 2> gpurun_out/c1_te_ops.err
