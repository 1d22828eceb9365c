"""Shared-memory fill model of the tcgen05 conv kernels (no GPU needed).

A conv launch moves tiles x k_blocks x (A stage + B stage) bytes from L2 into shared memory; on B200 that path is
capped per SM (~43-64 B/clk) and chip-wide (LTS), so fill bytes - not algorithmic bytes - bound the K loops.
Prints, per conv of ResNet50 at batch G: tiles, CTAs, fill MB for a given (BM=128, BN) tiling and the resulting
lower bounds, next to the tensor time of the issued MMAs (x3 for the bf16x3 fp32-parity path)."""
import sys
sys.path.insert(0, ".")
from defer_b200 import applications
from defer_b200.planner import plan_stage
from defer_b200 import _cabi as A


def convs(model):
    plan = plan_stage(model, is_first=True, is_last=True)
    out = []
    for o in plan.ops:
        if o.kind != A.OP_CONV:
            continue
        hi, wi, ci, _ = plan.bufs[o.in0]
        ho, wo, co, _ = plan.bufs[o.out]
        out.append(dict(name=o.layers[0], h=hi, w=wi, cin=ci, ho=ho, wo=wo, cout=co, k=o.kh, s=o.sh, res=bool(o.flags & 2)))
    return out


def m_tiles(c, G):
    if c["k"] == 1 and c["s"] == 1:
        return -(-G * c["ho"] * c["wo"] // 128), 1.0
    wo, ho = c["wo"], c["ho"]
    parts_w = -(-wo // 128)
    tw = -(-wo // parts_w)
    tiles_w = -(-wo // tw)
    th = min(128 // tw, ho)
    tiles_h = -(-ho // th)
    th = -(-ho // tiles_h)
    tn = 1
    if th == ho and tiles_w == 1:
        tn = max(1, min(G, 128 // (th * tw)))
    tiles_n = -(-G // tn)
    eff = G * ho * wo / (tiles_n * tiles_h * tiles_w * 128)
    return tiles_n * tiles_h * tiles_w, eff


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    planes = 2
    m = applications.ResNet50()
    tot = {}
    print(f"{'layer':18s} {'M':>6s} {'K':>5s} {'N':>5s} mt  eff | " + " | ".join(f"BN{bn}: ctas fillMB" for bn in (64, 128, 256)))
    for c in convs(m):
        if c["cin"] < 64:
            continue
        mt, eff = m_tiles(c, G)
        K = c["k"] * c["k"] * c["cin"]
        kb = K // 64
        row = f"{c['name']:18s} {G*c['ho']*c['wo']:6d} {K:5d} {c['cout']:5d} {mt:3d} {eff:4.2f} |"
        for bn in (64, 128, 256):
            b = min(bn, c["cout"])
            nt = c["cout"] // b
            fill = mt * nt * kb * planes * (128 + b) * 128 / 1e6
            out = G * c["ho"] * c["wo"] * c["cout"] * 4 * (2 if c["res"] else 1) / 1e6
            tot[bn] = tot.get(bn, 0) + fill + out
            row += f" {mt*nt:5d} {fill:7.1f} |"
        mma_us = 3 * 2 * G * c["ho"] * c["wo"] * c["cout"] * K / 2.25e15 * 1e6 / eff
        row += f" mma {mma_us:5.1f}us"
        print(row)
    for bn, v in tot.items():
        print(f"BN={bn}: total smem fill + epilogue traffic {v:8.1f} MB per step of {G} images -> {v/6.5e6*1e3:7.1f} us at 6.5 TB/s, "
              f"{v/12e6*1e3:7.1f} us at 12 TB/s")


if __name__ == "__main__":
    main()
