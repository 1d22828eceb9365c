#!/bin/bash
# ncu launch list (device time per launch, cold + serialised) of one bench run at the default config; depth sweep
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-roofline --depth 1 > gpurun_out/r2j_ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/r2j_ncu_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2j_launches.csv')) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split('(')[0][-60:]
    v = float(r[iv].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/1e3:10.1f} us {100*v/tot:5.1f}%  x{n:4d}  {k}")
PY
run() { # name, args...
  name=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-roofline "$@" > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2j_$name.json').read().strip().splitlines()[-1])
    print('value %.0f ms/step %.3f' % (d['value'], d['ms_per_step']))
except Exception as e:
    print('no line', e)
PY
)"
}
run g16_d4 --coalesce 16 --depth 4
run g16_d5 --coalesce 16 --depth 5
run g16_d10 --coalesce 16 --depth 10
run g32_d4 --coalesce 32 --depth 4
run g32_d5 --coalesce 32 --depth 5
run g32_d10 --coalesce 32 --depth 10
run g64_d4 --coalesce 64 --depth 4
DEFER_STREAM_EVEN_GRID=0 run g16_d4_noeven --coalesce 16 --depth 4
