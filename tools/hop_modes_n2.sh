#!/bin/bash
# 2 GPUs: hop modes (copy engine | TMA store to peer | per-thread peer stores): parity, then N=2 throughput
mkdir -p gpurun_out
for hop in copy tma direct; do
  DEFER_HOP=$hop timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q --timeout 600 > gpurun_out/r2n_pytest_$hop.log 2>&1
  echo "pytest(dist, hop=$hop) rc=$?"; tail -n 3 gpurun_out/r2n_pytest_$hop.log
done
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q --timeout 300 -k "pipeline or defer_api or coalesced or poison or unfused or vgg16_4 or resnet152" > gpurun_out/r2n_pytest_model.log 2>&1
echo "pytest(model pipelines, hop=copy) rc=$?"; tail -n 4 gpurun_out/r2n_pytest_model.log
brun() { # name hop N args...
  name=$1; hop=$2; N=$3; shift; shift; shift
  DEFER_HOP=$hop timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2n_$name.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f parity %s' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d.get('parity_rel_err')))
except Exception as e:
    print('no line', e)
PY
)"; grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/r2n_$name.err | tail -n 3
}
brun n2_copy copy 2 --steps 20 --warmup 5
brun n2_copy_k200 copy 2 --steps 200 --warmup 20
brun n2_tma tma 2 --steps 20 --warmup 5
brun n2_tma_k200 tma 2 --steps 200 --warmup 20
brun n2_direct direct 2 --steps 20 --warmup 5
