#!/usr/bin/env python3
"""Per-kernel summary of the LAST decode in an ncu launch list of profiles/run_parse.py
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("unnamed>::", "")
        agg.setdefault((row["ID"], k), {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    items = list(agg.items())
    starts = [i for i, (key, _) in enumerate(items) if "tile_exit" in key[1]]
    summ = collections.OrderedDict()
    for (_, k), m in items[starts[-1]:]:
        s = summ.setdefault(k, [0, 0.0, 0.0, 0.0])
        s[0] += 1
        s[1] += m.get("gpu__time_duration.sum", 0) / 1e3
        s[2] += m.get("dram__bytes_read.sum", 0)
        s[3] += m.get("dram__bytes_write.sum", 0)
    total = sum(v[1] for v in summ.values())
    print(f"{'kernel':44} {'n':>3} {'us':>9} {'share':>6} {'dram rd MB':>11} {'dram wr MB':>11}")
    for k, (n, t, rd, wr) in summ.items():
        print(f"{k[:44]:44} {n:3d} {t:9.1f} {t / total:6.1%} {rd / 1e6:11.1f} {wr / 1e6:11.1f}")
    print(f"{'total':44} {'':3} {total:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
