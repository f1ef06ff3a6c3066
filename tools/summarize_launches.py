"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.
usage: python tools/summarize_launches.py launches.csv [first_launch [last_launch]]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    rows = [r for r in rows if lo <= r[0] < hi]
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, ns, _, _ in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(anonymous namespace\)::", "", short)
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"launches {len(rows)}  total {total / 1e6:.3f} ms  (ids {lo}..{min(hi, rows[-1][0] + 1) if rows else lo})")
    print(f"{'kernel':70s} {'n':>6s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:6d} {ns / 1e6:10.3f} {100 * ns / total:6.1f}% {ns / n / 1e3:10.1f}")


if __name__ == "__main__":
    main()
