#!/usr/bin/env python3
"""Average the per-block host phase timings ([ppd] lines printed under PPD_TIMING=1) of the LAST batch
in a bench.py stderr log:  python profiles/phase_times.py gpurun_out/bench.err <blocks per step>"""
import collections
import sys

lines = [l.split() for l in open(sys.argv[1]) if l.startswith("[ppd]") and l.split()[-1] == "ms"]
n = int(sys.argv[2])
per = collections.defaultdict(list)
for l in lines:
    per[l[1]].append(float(l[-2]))
# the last measurement in the log is the three single-block latency calls: skip them
out = {}
for k, v in per.items():
    v = v[:-3] if len(v) > 3 else v
    out[k] = sum(v[-n:]) / max(1, len(v[-n:]))
print(" ".join(f"{k}={v:.1f}ms" for k, v in out.items()), " sum=%.1fms" % sum(out.values()))
