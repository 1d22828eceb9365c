#!/bin/bash
# 8-GPU box sweep: scaling of the headline config + the other BASELINE configs. Writes gpurun_out/sweep_*.json
mkdir -p gpurun_out
run() { # name n extra-args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 240 python bench.py --gpus 1 --steps 600 --warmup 30 --no-cpu --no-roofline "$@" > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 600 --warmup 30 "$@" > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_$name.json").read().strip().splitlines()[-1])
    print("$name", "N=%d"%d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],4), d["config"].get("cuts",{}).get("cuts"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/sweep_$name.err").read()[-1500:])
PY
}
run r50_f32_n1 1
run r50_f32_n2 2
run r50_f32_n4 4
run r50_f32_n8 8
run r50_f32_n8_bal 8 --cuts balanced
run r50_f32_n4_bal 4 --cuts balanced
run r50_bf16_n8 8 --dtype bfloat16
run r50_bf16_n8_bal 8 --dtype bfloat16 --cuts balanced
run r152_bf16_n8 8 --model resnet152 --dtype bfloat16
run vgg16_f32_n4 4 --model vgg16
