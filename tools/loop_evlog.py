"""Reads the txn loop's per-thread event log of one txn (development build: make -C csrc prof, PPD_LIB=..._prof.so,
PPD_LOOP_EVLOG=<txn>:<file>): how long each phase lasts, how long its busiest thread is busy, and that thread's events."""
import csv, sys
from collections import defaultdict

NAMES = {1: "walk>", 2: "walk<", 3: "fill>", 4: "fill<", 5: "announce>", 6: "announce<", 10: "climb1>", 11: "terminal", 12: "last-report", 13: "assembled",
         14: "climb1<", 20: "climb2>", 24: "climb2<", 30: "record>", 31: "record<"}


def main(path):
    ev = sorted((int(r["clock"]), int(r["thread"]), int(r["event"])) for r in csv.DictReader(open(path)))
    t0 = ev[0][0]
    phases = [(c - t0, e - 100) for c, t, e in ev if e >= 100]
    print("phase ends (cycles since the first event):", ", ".join(f"{k}:{c}" for c, k in phases))
    per_thread = defaultdict(list)
    for c, t, e in ev:
        if e < 100:
            per_thread[t].append((c - t0, e))
    for a, b, name in ((1, 2, "walk"), (3, 4, "fill"), (5, 6, "announce"), (10, 14, "climb1"), (20, 24, "climb2"), (30, 31, "record")):
        best = None
        n = 0
        starts, ends = [], []
        for t, evs in per_thread.items():
            s = [c for c, e in evs if e == a]
            f = [c for c, e in evs if e == b]
            if not s or not f:
                continue
            n += len(s)
            starts.append(min(s)), ends.append(max(f))
            busy = sum(y - x for x, y in zip(s, f))
            if best is None or busy > best[0]:
                best = (busy, t)
        if best is None:
            continue
        print(f"{name:9s} {n:4d} items, first start {min(starts)}, last end {max(ends)} (span {max(ends) - min(starts)}), busiest thread {best[1]}: {best[0]} cycles")
        evs = [(c, e) for c, e in per_thread[best[1]] if min(starts) <= c <= max(ends)]
        print("          " + " ".join(f"{NAMES.get(e, e)}@{c}" for c, e in evs[:40]))
        # the thread that ends last
        last_t = max(per_thread, key=lambda t: max([c for c, e in per_thread[t] if e == b] or [0]))
        evs = [(c, e) for c, e in per_thread[last_t] if min(starts) <= c <= max(ends)]
        print(f"          last to end: thread {last_t}: " + " ".join(f"{NAMES.get(e, e)}@{c}" for c, e in evs[:40]))


if __name__ == "__main__":
    main(sys.argv[1])
