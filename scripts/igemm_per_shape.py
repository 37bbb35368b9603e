"""Formats DRB_PROFILE_DUMP=1 output (one line per tensor-core GEMM launch, CUDA events around it) into the
per-shape table of profiles/r02_igemm_per_shape.txt.
    DRB_PROFILE_DUMP=1 python bench.py --stage register --steps 1 --warmup 3 --no-cpu-baseline 2> dump_fp32.txt
    DRB_PROFILE_DUMP=1 python bench.py --stage register --precision bf16 ...                    2> dump_bf16.txt
    python scripts/igemm_per_shape.py dump_fp32.txt:3 dump_bf16.txt:1 [peak TFLOP/s]
(file:N = MMAs the tensor pipe executes per product in that mode)."""
import collections
import re
import sys

PEAK = 1388.0


def table(path, mma_per_product):
    rows = collections.OrderedDict()
    for line in open(path):
        m = re.search(r"igemm M=\s*(-?\d+) Cin=\s*(\d+) Cout=\s*(\d+) k=(\d+)( \(tile list\))?: ([\d.]+) us, ([\d.]+) TFLOP/s", line)
        if not m or int(m[1]) < 0:
            continue
        key = (int(m[1]), int(m[2]), int(m[3]), int(m[4]), bool(m[5]))
        r = rows.setdefault(key, [0, 0.0, 0.0])
        r[0] += 1
        r[1] += float(m[6])
        r[2] += float(m[6]) * float(m[7])          # us x TFLOP/s = executed MFLOP
    n = sum(r[0] for r in rows.values())
    total = sum(r[1] for r in rows.values())
    flop = sum(r[2] for r in rows.values())
    print("%s (%d MMA per product): %d launches, %.2f ms per pair, %.0f TFLOP/s executed overall = %.0f %% of the MMA peak"
          % (path.split("/")[-1], mma_per_product, n, total / 1e3, flop / total, 100.0 * mma_per_product * flop / total / PEAK))
    print("  shape (M = voxels or tokens; L = tile list)    n   total us   us each   TFLOP/s  MMA-rate of peak")
    for (M, ci, co, k, lst), (cnt, us, mf) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        tf = mf / us
        print("  M=%7d Cin=%5d Cout=%5d k=%d %s       %3d  %9.1f %9.1f %9.0f %9.0f   %4.0f%%"
              % (M, ci, co, k, "L" if lst else " ", cnt, us, us / cnt, tf, tf * mma_per_product, 100.0 * tf * mma_per_product / PEAK))
    print()


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if ":" in a]
    rest = [a for a in sys.argv[1:] if ":" not in a]
    if rest:
        PEAK = float(rest[0])
    for a in args:
        p, n = a.rsplit(":", 1)
        table(p, int(n))
