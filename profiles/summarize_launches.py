"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [first_id last_id]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0], ns))
    if len(sys.argv) > 3:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
        rows = [r for r in rows if lo <= r[0] <= hi]
    tot = sum(r[2] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for _, k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    print("launches %d total %.3f ms" % (len(rows), tot / 1e6))
    print("%-60s %8s %12s %8s" % ("kernel", "count", "total_us", "share"))
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.1f %7.1f%%" % (k[:60], c, ns / 1e3, 100 * ns / tot))


if __name__ == "__main__":
    main()
