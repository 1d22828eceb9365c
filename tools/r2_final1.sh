#!/bin/bash
# round 2 evidence run (1 GPU): full GPU suite, default bench line, exact-fp32 SIMT line, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2q_pytest.log 2>&1
echo "pytest(all) rc=$?"; tail -n 5 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err
echo "bench default rc=$?"; head -c 400 gpurun_out/r2q_bench_n1.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --dtype bfloat16 --no-cpu > gpurun_out/r2q_bench_n1_bf16.json 2> gpurun_out/r2q_bench_n1_bf16.err
echo "bench bf16 rc=$?"; head -c 300 gpurun_out/r2q_bench_n1_bf16.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --dtype float32_simt --coalesce 8 --no-cpu --no-roofline > gpurun_out/r2q_bench_n1_simt.json 2> gpurun_out/r2q_bench_n1_simt.err
echo "bench exact-fp32 SIMT rc=$?"; head -c 300 gpurun_out/r2q_bench_n1_simt.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --coalesce 1 --depth 20 --no-cpu > gpurun_out/r2q_bench_n1_g1.json 2> gpurun_out/r2q_bench_n1_g1.err
echo "bench G=1 rc=$?"; head -c 300 gpurun_out/r2q_bench_n1_g1.json; echo
# ncu: one launch of every kernel of a 32-image step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2q_launches_g32.csv \
   python tools/run_stage_once.py resnet50 float32 32 > gpurun_out/r2q_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -n 1 gpurun_out/r2q_ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2q_full_step_g32 \
   python tools/run_stage_once.py resnet50 float32 32 > gpurun_out/r2q_ncu_full.log 2>&1
echo "ncu full step rc=$?"; tail -n 1 gpurun_out/r2q_ncu_full.log
# the standalone element-wise kernels + hop flags only appear at odd cut points
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"eltwise|relu_planes|pad_kernel|copy|flag" -f -o gpurun_out/r2q_full_oddcuts_g8 \
   python tools/run_stage_once.py resnet50 float32 8 conv1,activation_9,avg_pool > gpurun_out/r2q_ncu_oddcuts.log 2>&1
echo "ncu odd cuts rc=$?"; tail -n 1 gpurun_out/r2q_ncu_oddcuts.log
ls -la gpurun_out/r2q_*.ncu-rep
