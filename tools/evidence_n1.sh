#!/bin/bash
# round 2 evidence run (1 GPU), second part: bench lines + ncu captures exported to CSV on the box (reports are too big to bring back)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 300 -k "megakernel" > gpurun_out/r2r_pytest_mega.log 2>&1
echo "pytest(mega) rc=$?"; tail -n 3 gpurun_out/r2r_pytest_mega.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err
echo "bench default rc=$?"; head -c 200 gpurun_out/r2r_bench_n1.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --dtype bfloat16 --no-cpu > gpurun_out/r2r_bench_n1_bf16.json 2> gpurun_out/r2r_bench_n1_bf16.err
timeout 600 python bench.py --steps 20 --warmup 5 --dtype float32_simt --coalesce 8 --no-cpu --no-roofline > gpurun_out/r2r_bench_n1_simt.json 2> gpurun_out/r2r_bench_n1_simt.err
timeout 600 python bench.py --steps 20 --warmup 5 --coalesce 1 --depth 20 --no-cpu > gpurun_out/r2r_bench_n1_g1.json 2> gpurun_out/r2r_bench_n1_g1.err
timeout 600 python bench.py --steps 20 --warmup 5 --coalesce 16 --no-cpu > gpurun_out/r2r_bench_n1_g16.json 2> gpurun_out/r2r_bench_n1_g16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2r_launches_g32.csv \
   python tools/run_stage_once.py resnet50 float32 32 > gpurun_out/r2r_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -n 1 gpurun_out/r2r_ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/full_step_g32 \
   python tools/run_stage_once.py resnet50 float32 32 > gpurun_out/r2r_ncu_full.log 2>&1
echo "ncu full step rc=$?"; tail -n 1 gpurun_out/r2r_ncu_full.log
ncu -i /tmp/full_step_g32.ncu-rep --page raw --csv > gpurun_out/r2r_full_step_g32_raw.csv 2>/dev/null
# source-level stall samples of the dominant kernel (first conv_stream<2,128> launch of the step)
ncu -i /tmp/full_step_g32.ncu-rep --page source --csv -k regex:conv_stream --launch-skip 3 --launch-count 1 > gpurun_out/r2r_full_stream_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"eltwise|relu_planes|pad_kernel|copy|flag" -f -o /tmp/full_oddcuts_g8 \
   python tools/run_stage_once.py resnet50 float32 8 conv1,activation_9,avg_pool > gpurun_out/r2r_ncu_oddcuts.log 2>&1
echo "ncu odd cuts rc=$?"; tail -n 1 gpurun_out/r2r_ncu_oddcuts.log
ncu -i /tmp/full_oddcuts_g8.ncu-rep --page raw --csv > gpurun_out/r2r_full_oddcuts_g8_raw.csv 2>/dev/null
du -sh gpurun_out
