#!/bin/bash
# 2 GPUs: N=2 depth / steps matrix of the device-timed window vs e2e
mkdir -p gpurun_out
brun() { # name N args...
  name=$1; N=$2; shift; shift
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --no-cpu --no-roofline "$@" > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  fi
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2l_$name.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f wall/step %.3f' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d['wall_ms_per_step_incl_fill_drain']))
except Exception as e:
    print('no line', e)
PY
)"
}
brun n1_k20 1 --steps 20 --warmup 5
brun n1_k200 1 --steps 200 --warmup 20
brun n2_d2_k20 2 --steps 20 --warmup 5 --depth 2
brun n2_d2_k200 2 --steps 200 --warmup 20 --depth 2
brun n2_d4_k20 2 --steps 20 --warmup 5 --depth 4
brun n2_d4_k200 2 --steps 200 --warmup 20 --depth 4
brun n2_d5_k200 2 --steps 200 --warmup 20 --depth 5
brun n2_g32_d4_k200 2 --steps 200 --warmup 20 --depth 4 --coalesce 32
brun n2_g32_d2_k200 2 --steps 200 --warmup 20 --depth 2 --coalesce 32
