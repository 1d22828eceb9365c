#!/bin/bash
# round-1 (session 2) evidence: bench line, ncu launch list, GPU tests, one full ncu capture
mkdir -p gpurun_out
timeout 170 python bench.py --cpu-seconds 6 > gpurun_out/f2_bench_n1.json 2> gpurun_out/f2_bench_n1.err; tail -c 200 gpurun_out/f2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/f2_bench_n1.json").read().strip().splitlines()[-1])
    print("bench value", round(d["value"],1), "e2e", round(d.get("e2e",{}).get("value",0),1), "roof", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("bench FAILED", e)
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f2_launches.csv python bench.py --steps 3 --warmup 3 --depth 1 --no-e2e --no-cpu --no-roofline > /dev/null 2> gpurun_out/f2_ncu1.err
timeout 260 python -m pytest tests -m gpu -q --timeout 100 > gpurun_out/f2_pytest.log 2>&1; tail -n 12 gpurun_out/f2_pytest.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 40 -c 3 -o gpurun_out/f2_prof_conv python bench.py --steps 3 --warmup 3 --depth 1 --no-e2e --no-cpu --no-roofline > /dev/null 2> gpurun_out/f2_ncu2.err
ls -la gpurun_out | grep f2_
