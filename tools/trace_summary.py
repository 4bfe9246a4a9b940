"""Summary of a PPD_TRACE file (lane, block, label, device_ms, host_ms): per stage, how long it lasts on the device
when the lanes run together (interval from the previous mark of the same block), and where the time of a block goes."""
import csv, sys
from collections import defaultdict


def main(path, last_blocks=None):
    rows = list(csv.DictReader(open(path)))
    blocks = defaultdict(list)
    for r in rows:
        blocks[int(r["block"])].append((r["label"], float(r["device_ms"]), float(r["host_ms"]), int(r["lane"])))
    ids = sorted(blocks)
    if last_blocks:
        ids = ids[-last_blocks:]
    dur = defaultdict(list)
    order = []
    t_begin, t_end = [], []
    for b in ids:
        marks = blocks[b]
        t_begin.append(marks[0][1]), t_end.append(marks[-1][1])
        for (l0, d0, h0, _), (l1, d1, h1, _) in zip(marks, marks[1:]):
            if l1 not in order:
                order.append(l1)
            dur[l1].append(d1 - d0)
    span = max(t_end) - min(t_begin)
    print(f"{len(ids)} blocks, makespan {span:.1f} ms -> {1e3 * len(ids) / span:.1f} blocks/s; mean block latency {sum(e - s for s, e in zip(t_begin, t_end)) / len(ids):.1f} ms")
    print(f"{'stage (ends at mark)':24s} {'mean ms':>9s} {'median':>9s} {'max':>9s} {'sum/makespan':>13s}")
    for l in order:
        d = sorted(dur[l])
        print(f"{l:24s} {sum(d) / len(d):9.3f} {d[len(d) // 2]:9.3f} {d[-1]:9.3f} {sum(d) / span:13.2f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
