"""Development helper: one full-size C2 block decoded a few times alone (PPD_TIMING=1 prints the phases)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from proof_protocol_decoder_b200.lib import Context

flat = bench.c2_block(2, 1.0)
ctx = Context(0)
f = ctx.pinned_copy(flat)
for i in range(4):
    t0 = time.perf_counter()
    with ctx.block_decode_view(f) as v:
        n = v.nbytes
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"decode {i}: {dt*1e3:.2f} ms out={n} loops={st['txn_loops_on_gpu']} txn_ms={st['txn_gpu_ms']:.2f} hash_ms={st['gpu_ms']:.2f} parse_ms={st['parse_gpu_ms']:.2f} dump_ms={st['dump_gpu_ms']:.2f} busy={st['host_busy_ms']:.2f} wait={st['host_wait_ms']:.2f} launches={st['kernel_launches']}", flush=True)
