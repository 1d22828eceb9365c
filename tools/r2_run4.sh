#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the real cross-GPU hop (parity through Node.run + CUDA-IPC), then bench at N = 2 and N = 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q --timeout 600 -s > gpurun_out/r2e_pytest_dist.log 2>&1
echo "pytest(dist) rc=$?"; tail -n 12 gpurun_out/r2e_pytest_dist.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q --timeout 300 -k "defer_api or coalesced or poison" > gpurun_out/r2e_pytest_model.log 2>&1
echo "pytest(model subset, 2 GPUs visible) rc=$?"; tail -n 5 gpurun_out/r2e_pytest_model.log
brun() { # name N args...
  name=$1; N=$2; shift; shift
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu "$@" > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err
  fi
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2e_$name.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f parity %s' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d.get('parity_rel_err')))
except Exception as e:
    print('no line', e)
PY
)"; tail -n 3 gpurun_out/r2e_$name.err
}
brun n1 1 --no-roofline
brun n2 2
brun n2_long 2 --steps 200 --warmup 20
brun n2_ref 2 --impl reference
