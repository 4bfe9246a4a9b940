#!/usr/bin/env python3
"""Per-kernel table of ONE block decode from an ncu launch list of `python profiles/run_parse.py <repeats>`:
  python profiles/summarize_block.py <launches.csv> <repeats> [title]
The launches are split evenly over the repeats and the last repeat (warm) is tabulated.  ncu serialises the kernels and
runs them cold-cache: read shares, not absolutes."""
import collections
import csv
import sys


def main():
    path, reps = sys.argv[1], int(sys.argv[2])
    title = sys.argv[3] if len(sys.argv) > 3 else "one block"
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        e = L.setdefault(d["ID"], {"kernel": d["Kernel Name"].split("(")[0].replace("void ", "").replace("ppd::", "").replace("<unnamed>::", ""), "grid": d["Grid Size"]})
        try:
            e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    items = list(L.values())
    per = len(items) // reps
    items = items[-per:]
    agg = collections.OrderedDict()
    for o in items:
        a = agg.setdefault(o["kernel"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += o.get("gpu__time_duration.sum", 0.0)
        a[2] += o.get("dram__bytes_read.sum", 0.0)
        a[3] += o.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values()) or 1.0
    print(f"{title}: {per} launches, {tot / 1e3:.1f} us of kernel time (serialised by ncu)")
    print(f"{'kernel':44s} {'launches':>8s} {'time us':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
    for k, a in agg.items():
        print(f"{k[:44]:44s} {a[0]:8d} {a[1] / 1e3:10.1f} {100 * a[1] / tot:6.1f}% {a[2] / 1e6:11.1f} {a[3] / 1e6:11.1f}")


if __name__ == "__main__":
    main()
