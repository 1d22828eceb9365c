"""Summarise a DEFER_TIMELINE log: SM busy fraction, CTA lifetimes per op, concurrency."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.int64)
a = a[a[:, 0] > 0]
t0, t1, sm, tag = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
T0, T1 = np.percentile(t0, 60), np.percentile(t1, 95)   # steady-state window
m = (t0 >= T0) & (t1 <= T1)
win = (T1 - T0) / 1e3
print(f"entries {len(a)}  window {win:.0f} us  CTAs in window {m.sum()}  -> {m.sum()/win:.2f} CTAs/us")
life = (t1 - t0)[m] / 1e3
print(f"CTA lifetime us: median {np.median(life):.2f} mean {life.mean():.2f} p90 {np.percentile(life,90):.2f} p99 {np.percentile(life,99):.2f}")
print(f"sum of CTA lifetimes / (148 SMs x window) = {life.sum()/(148*win):.3f} (avg resident conv CTAs per SM)")
# per-SM busy fraction (union of intervals)
busy = []
for s in np.unique(sm[m]):
    iv = np.stack([t0[m & (sm == s)], t1[m & (sm == s)]], 1)
    iv = iv[np.argsort(iv[:, 0])]
    tot, cur_s, cur_e = 0, iv[0, 0], iv[0, 1]
    for b, e in iv[1:]:
        if b > cur_e:
            tot += cur_e - cur_s; cur_s, cur_e = b, e
        else:
            cur_e = max(cur_e, e)
    tot += cur_e - cur_s
    busy.append(tot / (T1 - T0))
print(f"SMs seen {len(busy)}; fraction of time an SM hosts >=1 conv CTA: mean {np.mean(busy):.3f} min {np.min(busy):.3f} max {np.max(busy):.3f}")
if a.shape[1] >= 8:
    ph = a[m][:, 4:8] - a[m][:, 0:1]
    ok = (a[m][:, 4:8] > 0).all(axis=1)
    ph = ph[ok] / 1e3
    tot_l = life[ok]
    print("phase medians (us since CTA start): setup %.2f | first operands %.2f | accumulator ready %.2f | epilogue done %.2f | end %.2f" %
          (np.median(ph[:, 0]), np.median(ph[:, 1]), np.median(ph[:, 2]), np.median(ph[:, 3]), np.median(tot_l)))
    print("phase means                        : setup %.2f | first operands %.2f | accumulator ready %.2f | epilogue done %.2f | end %.2f" %
          (ph[:, 0].mean(), ph[:, 1].mean(), ph[:, 2].mean(), ph[:, 3].mean(), tot_l.mean()))
print("per-op (ho, k, cout): n CTAs, median / mean lifetime us, share of CTA-time")
tot = life.sum()
for tg in np.unique(tag[m]):
    l = (t1 - t0)[m & (tag == tg)] / 1e3
    sel = m & (tag == tg)
    extra = ""
    if a.shape[1] >= 8:
        php = (a[sel][:, 4:8] - a[sel][:, 0:1]) / 1e3
        okp = (a[sel][:, 4:8] > 0).all(axis=1)
        if okp.any():
            md = np.median(php[okp], axis=0)
            extra = f"  phases med: setup {md[0]:5.2f} ops {md[1]:5.2f} acc {md[2]:6.2f} epi {md[3]:6.2f}"
    print(f"  ho={tg>>20:3d} k={(tg>>16)&15} cout={tg&0xffff:4d}  n={len(l):6d}  med {np.median(l):6.2f}  mean {l.mean():6.2f}  share {l.sum()/tot:5.3f}{extra}")
