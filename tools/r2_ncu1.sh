#!/bin/bash
# ncu --set full of conv_stream_kernel on four representative ResNet50 convs at batch 16 (one launch each, 2nd launch captured)
mkdir -p gpurun_out
cap() { # name backend shape...
  name=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_stream -s 1 -c 1 -f -o gpurun_out/r2g_$name \
      python tools/run_one_conv.py bf16x2 "$@" 2 > gpurun_out/r2g_$name.log 2>&1
  echo "$name rc=$?"; tail -n 2 gpurun_out/r2g_$name.log
}
cap c3x3_56_bn64 4 16 56 56 64 64 3 1 1
cap c3x3_28_bn128 5 16 28 28 128 128 3 1 1
cap c1x1_56_res_bn128 5 16 56 56 64 256 1 1 0
cap c1x1_14_k1024_bn64 4 16 14 14 1024 256 1 1 0
ls -la gpurun_out/*.ncu-rep
