"""Development helper: end-to-end blocks/s of ppd_blocks_decode_stream under different pipeline knobs.

  python tools/knob_sweep.py                      # the sweep: one child process per setting
  python tools/knob_sweep.py --child              # one measurement with the environment as it is

Every child reads DISTINCT (default 16) full-size C2 blocks from bench.py's cache under /tmp (a bench.py run of the
same box makes them), decodes 512 blocks per call through the stream entry point from page-locked buffers, and
prints the best of three timed calls.  Knobs are read when the library creates its context, hence the processes."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETTINGS = [
    {},
    {"PPD_PARSE_SLOTS": "12"},
    {"PPD_PARSE_SLOTS": "16"},
    {"PPD_PARSE_SLOTS": "32"},
    {"PPD_MAX_LANES": "96", "PPD_HOST_THREADS": "96", "PPD_PARSE_SLOTS": "12"},
    {"PPD_STREAM_POOL": "24", "PPD_LOOP_STREAMS": "6"},
    {"PPD_STREAM_POOL": "26", "PPD_LOOP_STREAMS": "4"},
    {"PPD_LOOP_CHUNK": "64"},
    {"PPD_PARSE_V1": "31"},
    {},
]


def child():
    from proof_protocol_decoder_b200.lib import Context

    distinct = int(os.environ.get("DISTINCT", "16"))
    per_call = int(os.environ.get("PER_CALL", "512"))
    flats = []
    for seed in range(2, 2 + distinct):
        p = f"/tmp/ppd_c2r2_seed{seed}_scale1.0.flat"
        if os.path.exists(p):
            flats.append(open(p, "rb").read())
    if not flats:
        print(json.dumps({"error": "no cached blocks under /tmp (run bench.py first)"}))
        return
    ctx = Context(0)
    pinned = [ctx.pinned_copy(f) for f in flats]
    batch = [pinned[i % len(pinned)] for i in range(per_call)]
    rates = []
    for it in range(5):
        got = []

        def on_done(i, o):
            if isinstance(o, Exception):
                raise o
            got.append(o.nbytes + o.view[0] + o.view[o.nbytes - 1])
            o.close()

        t0 = time.perf_counter()
        ctx.blocks_decode_stream(batch, on_done)
        dt = time.perf_counter() - t0
        assert len(got) == per_call
        if it >= 2:
            rates.append(per_call / dt)
    st = ctx.stats()
    print(json.dumps({"blocks_per_sec_best": max(rates), "all": [round(r, 1) for r in rates], "distinct": len(flats),
                      "loops_on_gpu": st["txn_loops_on_gpu"], "host_busy_ms_per_block": st["host_busy_ms"] / per_call,
                      "host_wait_ms_per_block": st["host_wait_ms"] / per_call}))
    ctx.close()


def main():
    if "--child" in sys.argv:
        child()
        return
    for s in SETTINGS:
        env = dict(os.environ)
        env.setdefault("PPD_HOST_THREADS", "64")
        env.update(s)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=300)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-300:]
        print(json.dumps(s), "->", line, flush=True)


if __name__ == "__main__":
    main()
