"""Aggregates `ncu --page source --csv --print-source cuda,sass` per CUDA source line (argv[1] = csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
cur = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 10 or r[0] == 'Line No' or r[2] != '-': continue
    try: n = int(r[7]); t = int(r[8]); s = int(r[6])
    except ValueError: continue
    agg[(cur, int(r[0]))] = (n, t, s, r[1])
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print('total warp-inst', tot, 'samples', ts)
for k in sorted(agg):
    n, t, s, src = agg[k]
    if n > tot * thr or s > ts * thr:
        print('%s:%4d inst %5.2f%% lanes %5.1f samp %5.2f%% | %s' % (k[0][:8], k[1], 100 * n / tot, t / max(n, 1), 100 * s / ts, src.strip()[:95]))
