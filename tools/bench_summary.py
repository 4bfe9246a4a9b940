"""Prints the headline numbers of a bench.py JSON line (development helper)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    e2e = d.get("e2e", {})
    lim = d.get("limiter", {})
    print(path)
    print("  value %.3g nodes/s  ms_per_step %.2f  roofline.frac %.3f" % (d["value"], d["ms_per_step"], d.get("roofline", {}).get("frac", 0)))
    print("  device ms/step:", {k: round(v, 2) for k, v in d.get("device_ms_per_step", {}).items() if isinstance(v, (int, float))})
    print("  device-resident blocks/s %.1f   e2e blocks/s %.1f  (ms/step %.1f, single block %.1f ms)" % (d.get("blocks_per_sec", 0), e2e.get("blocks_per_sec", 0), e2e.get("ms_per_step", 0), e2e.get("single_block_latency_ms", 0)))
    print("  h2d/d2h MB per block: %.1f / %.1f   host busy ms per block %.2f" % (lim.get("pcie_bytes_per_block", {}).get("h2d", 0) / 1e6, lim.get("pcie_bytes_per_block", {}).get("d2h", 0) / 1e6, lim.get("host_busy_ms_per_block", 0)))
    print("  limiter:", lim.get("name"), {k: (round(v, 1) if v else v) for k, v in lim.get("bounds_blocks_per_sec", {}).items()}, "pcie GB/s", round(lim.get("pcie_gbs_per_direction", 0), 1))
    print("  loops on device per step:", d["config"].get("txn_loops_on_device_per_step"), " threads", d["config"].get("host_threads_per_gpu"), " cores", d["config"].get("host_cores"))
