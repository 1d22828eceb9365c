#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2h_trace.txt
export DEFER_UMMA_TRACE=gpurun_out/r2h_trace.txt
python tools/run_one_conv.py bf16x2 4 16 56 56 64 64 3 1 1 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 5 16 28 28 128 128 3 1 1 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 5 16 56 56 64 256 1 1 0 1 > /dev/null 2>&1
python tools/run_one_conv.py bf16x2 4 16 14 14 1024 256 1 1 0 1 > /dev/null 2>&1
cat gpurun_out/r2h_trace.txt
mkdir -p gpurun_out; rm -f gpurun_out/r2s_stem_trace.txt
DEFER_STEM_TRACE=gpurun_out/r2s_stem_trace.txt python tools/run_stage_once.py resnet50 float32 16 max_pooling2d > gpurun_out/r2s_log.txt 2>&1
tail -n 2 gpurun_out/r2s_log.txt; cat gpurun_out/r2s_stem_trace.txt | cut -c1-600
