#!/bin/bash
# final 8-GPU lines: headline pipeline (default config), balanced cuts, ResNet152 bf16 (config 5), VGG16 on 4 GPUs (config 4)
mkdir -p gpurun_out
brun() { # name N args...
  name=$1; N=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r2v_$name.json 2> gpurun_out/r2v_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2v_$name.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f parity %s cuts %s' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d.get('parity_rel_err'), d['engine']['cuts'].get('cuts')))
except Exception as e:
    print('no line', e)
PY
)"; grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/r2v_$name.err | tail -n 3
}
brun r50_n8 8 --steps 20 --warmup 5
brun r50_n8_balanced 8 --steps 20 --warmup 5 --cuts balanced
brun r152_bf16_n8 8 --steps 20 --warmup 5 --model resnet152 --dtype bfloat16
brun vgg16_n4 4 --steps 20 --warmup 5 --model vgg16
