"""Development helper: N full-size C2 blocks through ppd_blocks_decode_batch a few times with PPD_TRACE on; prints the
blocks/s of every pass and a summary of the stage timeline (tools/trace_summary.py prints the same from a file)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
trace = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "trace.csv")
distinct = int(os.environ.get("DISTINCT", "8"))
os.environ["PPD_TRACE"] = trace
os.environ.setdefault("PPD_HOST_THREADS", str(n))
import bench
flats = bench.c2_blocks([2 + j for j in range(distinct)], 1.0, os.cpu_count() or 1)
from proof_protocol_decoder_b200.lib import Context
ctx = Context(0)
flats = [ctx.pinned_copy(f) for f in flats]
batch = [flats[i % distinct] for i in range(n)]
for it in range(5):
    t0 = time.perf_counter()
    outs = ctx.blocks_decode_batch_view(batch)
    for o in outs:
        if isinstance(o, Exception):
            raise o
        o.close()
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"pass {it}: {n / dt:.1f} blocks/s ({dt * 1e3:.1f} ms) loops={st['txn_loops_on_gpu']} busy={st['host_busy_ms'] / n:.2f} wait={st['host_wait_ms'] / n:.2f}", flush=True)
ctx.close()
import trace_summary
trace_summary.main(trace, last_blocks=n)
