#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2s_stem_trace.txt
DEFER_STEM_TRACE=gpurun_out/r2s_stem_trace.txt python tools/run_stage_once.py resnet50 float32 16 max_pooling2d > gpurun_out/r2s_log.txt 2>&1
tail -n 2 gpurun_out/r2s_log.txt; cat gpurun_out/r2s_stem_trace.txt | cut -c1-600
