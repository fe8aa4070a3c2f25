"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one bench iteration.

    python profiles/summarise_launches.py gpurun_out/launches_rNN.csv > profiles/launches_rNN_summary.txt
An iteration is delimited by consecutive launches of vldd::row_normalise_kernel (first kernel of vldd_unrolled_match).
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "row_normalise_kernel" in n and "bwd" not in n]
    a, b = starts[-2], starts[-1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[a:b]:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: one bench iteration = launches [{a},{b}) = {b - a} launches, {tot:.1f} us serialised (cold-cache, under ncu)")
    print(f"# {'us':>10} {'share':>6} {'count':>5}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:12.1f} {100 * t / tot:5.1f}% {c:5d}  {n[:120]}")


if __name__ == "__main__":
    main(sys.argv[1])
