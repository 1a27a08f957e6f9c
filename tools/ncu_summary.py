#!/usr/bin/env python
"""Condense an `ncu --csv` launch list (one row per launch and metric) into one line per kernel name:
launch count, mean duration, mean DRAM read / write bytes and L2 sectors.  Usage: ncu_summary.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ix = {n: i for i, n in enumerate(rows[h])}
    per = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) < len(ix):
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0][-60:], r[ix.get("Grid Size", 0)])
        per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    agg = collections.OrderedDict()
    for (_, name, grid), m in per.items():
        a = agg.setdefault((name, grid), collections.defaultdict(float))
        a["n"] += 1
        for k, v in m.items():
            a[k] += v
    print(f"{'kernel':60s} {'grid':>18s} {'n':>4s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'L2 MB':>9s}")
    for (name, grid), a in agg.items():
        n = a["n"]
        print(f"{name:60s} {grid:>18s} {int(n):4d} {a['gpu__time_duration.sum'] / n / 1e3:9.1f} "
              f"{a['dram__bytes_read.sum'] / n / 1e6:9.1f} {a['dram__bytes_write.sum'] / n / 1e6:9.1f} "
              f"{a['lts__t_sectors.sum'] * 32 / n / 1e6:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
