#!/bin/bash
# per-CTA phase trace of the tcgen05 conv kernel for a few ResNet50 shapes (writes gpurun_out/trace.txt)
mkdir -p gpurun_out; rm -f gpurun_out/trace.txt
export DEFER_UMMA_TRACE=gpurun_out/trace.txt
for shape in "1 56 56 64 64 1 1 0" "1 56 56 64 64 3 1 1" "1 56 56 64 256 1 1 0" "1 28 28 128 128 3 1 1" "1 14 14 1024 256 1 1 0" "1 7 7 512 512 3 1 1" "1 7 7 512 2048 1 1 0"; do
  python tools/run_one_conv.py bf16x2 2 $shape 1 > /dev/null 2>&1
done
python - <<'PY'
import numpy as np
cur=None; rows=[]
def flush():
    if cur and rows:
        a=np.array(rows,dtype=float)
        t0=a[:,1]-a[:,1].min()
        print(cur)
        print("   start_skew_us max %.2f | cycles(med): setup %.0f first_full %.0f mma_issued %.0f acc_ready %.0f epi_done %.0f end %.0f  (max end %.0f)" % (
            t0.max()/1e3, *np.median(a[:,2:8],axis=0), a[:,7].max()))
for line in open("gpurun_out/trace.txt"):
    if line.startswith("#"):
        flush(); cur=line.strip(); rows=[]
    else:
        rows.append([float(x) for x in line.split()])
flush()
PY
