#!/bin/bash
# N = 2 and N = 4 with the shipped defaults (companion of evidence_n8.sh); run with --gpus 4
mkdir -p gpurun_out
for N in 2 4; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2z_n$N.json 2> gpurun_out/r2z_n$N.err
  echo "n$N rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2z_n$N.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f parity %s' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d.get('parity_rel_err')))
except Exception as e:
    print('no line', e)
PY
)"
done
