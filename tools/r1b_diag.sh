#!/bin/bash
mkdir -p gpurun_out
DEFER_UMMA_CSPLIT_KB=4 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e --batched-roofline 0 > gpurun_out/c2_kb4_ops.json 2> gpurun_out/c2_kb4_ops.err
tail -c 300 gpurun_out/c2_kb4_ops.err
for v in base kb4; do
  if [ $v = base ]; then E="DEFER_UMMA_CLUSTER=0"; else E="DEFER_UMMA_CSPLIT_KB=4"; fi
  env $E DEFER_TIMELINE=/tmp/tl_$v.txt timeout 120 python bench.py --steps 300 --warmup 20 --no-cpu --no-e2e --no-roofline > gpurun_out/c2_tl_$v.json 2> gpurun_out/c2_tl_$v.err
  python tools/timeline_stats.py /tmp/tl_$v.txt > gpurun_out/c2_tl_${v}_stats.txt 2>&1
  head -8 gpurun_out/c2_tl_${v}_stats.txt
done
