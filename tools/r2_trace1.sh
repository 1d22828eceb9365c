#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2h_trace.txt
export DEFER_UMMA_TRACE=gpurun_out/r2h_trace.txt
python tools/run_one_conv.py bf16x2 4 16 56 56 64 64 3 1 1 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 5 16 28 28 128 128 3 1 1 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 5 16 56 56 64 256 1 1 0 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 4 16 14 14 1024 256 1 1 0 1 > /dev/null 2>&1
cat gpurun_out/r2h_trace.txt
