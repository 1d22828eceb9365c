#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/trace_p.txt
export DEFER_UMMA_TRACE=gpurun_out/trace_p.txt
python tools/run_one_conv.py bf16x2 3 8 56 56 64 256 1 1 0 1
python tools/run_one_conv.py bf16x2 3 8 56 56 64 64 3 1 1 1
cat gpurun_out/trace_p.txt
