#!/bin/bash
# round-end single-GPU evidence run: tests, official bench lines, batched roofline, ncu launch list + full capture
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/final_pytest.log 2>&1; tail -n 3 gpurun_out/final_pytest.log
timeout 300 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 300 gpurun_out/final_bench_n1.err
timeout 200 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err
timeout 200 python bench.py --dtype bfloat16 --no-cpu > gpurun_out/final_bench_n1_bf16.json 2> gpurun_out/final_bench_n1_bf16.err
timeout 200 python bench.py --batch 32 --steps 30 --warmup 5 --depth 4 --no-cpu --no-e2e > gpurun_out/final_bench_n1_b32.json 2> gpurun_out/final_bench_n1_b32.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --depth 1 --no-e2e --no-cpu --no-roofline > /dev/null 2> gpurun_out/final_ncu1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 60 -c 2 -o gpurun_out/final_prof_conv python bench.py --steps 3 --warmup 3 --depth 1 --no-e2e --no-cpu --no-roofline > /dev/null 2> gpurun_out/final_ncu2.err
python - <<'PY'
import json
for f in ["final_bench_n1","final_ref_n1","final_bench_n1_bf16","final_bench_n1_b32"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "e2e", round(d.get("e2e",{}).get("value",0),1), "roof", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
