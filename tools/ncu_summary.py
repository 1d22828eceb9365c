"""Summarise an `ncu --page raw --csv` export: one line per kernel (launch count, mean duration, DRAM bytes, DRAM / L2 /
tensor-pipe utilisation), and optionally update profiles/ncu_traffic.json with the DRAM traffic per launch of the dominant
conv kernel (the `roofline.traffic` field of bench.py).

usage: ncu_summary.py raw.csv [--traffic-key resnet50_float32_b32 --kernel conv_stream --out profiles/ncu_traffic.json]"""
import argparse
import collections
import csv
import json
import re
from pathlib import Path

WANT = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "smem_dyn": "launch__shared_mem_per_block_dynamic",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
         "msecond": 1e3, "Kbyte/block": 1e3, "byte/block": 1.0, "Mbyte/block": 1e6}


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("defer::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()[-56:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic-key")
    ap.add_argument("--kernel", default="conv_stream")
    ap.add_argument("--out", default="profiles/ncu_traffic.json")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(v) for k, v in WANT.items() if v in hdr}
    ik = hdr.index("Kernel Name")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= ik:
            continue
        d = agg.setdefault(short(r[ik]), collections.defaultdict(list))
        for k, i in col.items():
            try:
                v = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            except ValueError:
                continue
            d[k].append(v)
    print(f"{'kernel':58s} {'n':>4s} {'us':>8s} {'grid':>5s} {'regs':>4s} {'smemKB':>7s} {'dramMB':>8s} {'dram%':>6s} {'L2%':>5s} {'sm%':>5s} {'tensor%':>7s}")
    mean = lambda x: sum(x) / len(x) if x else float("nan")   # noqa: E731
    for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1]["dur_us"])):
        print(f"{k:58s} {len(d['dur_us']):4d} {mean(d['dur_us']):8.1f} {mean(d['grid']):5.0f} {mean(d['regs']):4.0f} "
              f"{mean(d['smem_dyn']) / 1e3:7.1f} {(mean(d['dram_rd']) + mean(d['dram_wr'])) / 1e6:8.2f} {mean(d['dram_pct']):6.1f} "
              f"{mean(d['lts_pct']):5.1f} {mean(d['sm_pct']):5.1f} {mean(d['tensor_pct']):7.1f}")
    if a.traffic_key:
        sel = [d for k, d in agg.items() if a.kernel in k]
        n = sum(len(d["dur_us"]) for d in sel)
        tot = sum(sum(d["dram_rd"]) + sum(d["dram_wr"]) for d in sel)
        out = Path(a.out)
        js = json.loads(out.read_text()) if out.exists() else {}
        js[a.traffic_key] = {"traffic_bytes_per_launch": tot / max(n, 1), "launches": n,
                             "source": f"profiles/{Path(a.csv).name} (one `ncu --set full` capture of every launch of a step)",
                             "kernel": a.kernel}
        out.write_text(json.dumps(js, indent=1))
        print(f"{a.out}: {a.traffic_key} -> {tot / max(n, 1) / 1e6:.2f} MB per launch over {n} launches")


if __name__ == "__main__":
    main()
