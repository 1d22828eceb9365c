#!/bin/bash
# 8 GPUs with the copy-engine hop: headline pipeline at G = 16 / 32, short and long windows, 4 stages
mkdir -p gpurun_out
brun() { # name N args...
  name=$1; N=$2; shift; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2p_$name.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f wall/step %.3f parity %s' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d['wall_ms_per_step_incl_fill_drain'], d.get('parity_rel_err')))
except Exception as e:
    print('no line', e)
PY
)"; grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/r2p_$name.err | tail -n 3
}
brun n8_g16_k20 8 --steps 20 --warmup 5 --coalesce 16
brun n8_g16_k200 8 --steps 200 --warmup 20 --coalesce 16
brun n8_g32_k20 8 --steps 20 --warmup 5 --coalesce 32
brun n4_g16_k20 4 --steps 20 --warmup 5 --coalesce 16
brun n8_bf16_g16_k20 8 --steps 20 --warmup 5 --dtype bfloat16 --coalesce 16
