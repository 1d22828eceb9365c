#!/bin/bash
# first GPU call of the next round: validate the opt-in executors, then A/B them against the default
mkdir -p gpurun_out
DEFER_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "steal or fast_flags" > gpurun_out/s1_pytest.log 2>&1; tail -n 15 gpurun_out/s1_pytest.log
B="python bench.py --steps 400 --warmup 20 --no-cpu --no-e2e --no-roofline"
run() { local name=$1; shift
  env "$@" timeout 90 $B > gpurun_out/s1_$name.json 2> gpurun_out/s1_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/s1_{n}.json").read().strip().splitlines()[-1]); print(n, "value", round(d["value"],1))
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/s1_{n}.err").read()[-400:])
PY
}
run default
run steal4 DEFER_STEAL=1 DEFER_STEAL_CTAS=4
run steal6 DEFER_STEAL=1 DEFER_STEAL_CTAS=6
run steal9 DEFER_STEAL=1 DEFER_STEAL_CTAS=9
run mega8 DEFER_MEGA=1 DEFER_MEGA_CLUSTER=8
run fast1 DEFER_UMMA_FAST=1
run fast3 DEFER_UMMA_FAST=3
run bn64_st2 DEFER_UMMA_BN=64 DEFER_UMMA_STAGES=2
# per-tile timeline of the steal executor (claim -> epilogue done -> stored -> published), same summary tool
DEFER_STEAL=1 DEFER_TIMELINE=/tmp/tl_steal.txt timeout 120 $B > gpurun_out/s1_tl_steal.json 2> gpurun_out/s1_tl_steal.err
python tools/timeline_stats.py /tmp/tl_steal.txt > gpurun_out/s1_tl_steal_stats.txt 2>&1; head -8 gpurun_out/s1_tl_steal_stats.txt
