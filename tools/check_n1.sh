#!/bin/bash
# final 1-GPU check of the tree: full GPU suite, smoke(), PDL on/off, relu_planes capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2u_pytest.log 2>&1
echo "pytest(all) rc=$?"; tail -n 5 gpurun_out/r2u_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/r2u_smoke.log
run() { # name, env..., --, args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-roofline "$@" > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2u_$name.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f ms/step %.3f parity %s' % (d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], d.get('parity_rel_err')))
except Exception as e:
    print('no line', e)
PY
)"; tail -n 2 gpurun_out/r2u_$name.err
}
run base X=1 --
run pdl DEFER_PDL=1 --
run pdl_g16 DEFER_PDL=1 -- --coalesce 16
run base_g16 X=1 -- --coalesce 16
run vgg16 X=1 -- --model vgg16
run vgg16_tcstem DEFER_TC_STEM=2 -- --model vgg16
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"relu_planes|pad_kernel|copy" -f -o /tmp/full_relu \
   python tools/run_stage_once.py resnet50 float32 8 add_2,add_4 > gpurun_out/r2u_ncu_relu.log 2>&1
echo "ncu relu rc=$?"; tail -n 1 gpurun_out/r2u_ncu_relu.log
ncu -i /tmp/full_relu.ncu-rep --page raw --csv > gpurun_out/r2u_full_relu_raw.csv 2>/dev/null
du -sh gpurun_out
