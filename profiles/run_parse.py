#!/usr/bin/env python3
"""The GPU witness parse / pre-image arena build (ppd_parse.cu) on one config-2 block.

  python profiles/run_parse.py [repeats]

Decodes the block `repeats` times on one lane (PPD_HOST_THREADS=1) and prints one JSON line: witness bytes,
instructions, device time of the parse kernels (CUDA events around the three phases) and the byte rate that
gives against the measured HBM bandwidth; plus pinned host<->device copy rates of this box."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PPD_HOST_THREADS", "1")


def main():
    import bench
    import torch

    from proof_protocol_decoder_b200.lib import Context

    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    flat = bench.c2_blocks([2], 1.0, 1)[0]
    ctx = Context(0)
    pinned = ctx.pinned_copy(flat)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        with ctx.block_decode_view(pinned) as v:
            _ = v.view[0]
        wall = time.perf_counter() - t0
        st = ctx.stats()
        st["wall_ms"] = wall * 1e3
        if best is None or st["parse_gpu_ms"] < best["parse_gpu_ms"]:
            best = st
    # pinned copy rates
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    rates = {}
    for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        rates[name + "_gbs"] = 4 * n / (e0.elapsed_time(e1) / 1e3) / 1e9
    peaks = bench.load_peaks()
    sec = best["parse_gpu_ms"] / 1e3
    print(json.dumps({
        "workload": "one config-2 block, one lane",
        "witness_bytes": best["witness_bytes"], "witness_instructions": best["witness_instructions"],
        "witnesses_on_gpu": best["witnesses_on_gpu"], "parse_gpu_ms": best["parse_gpu_ms"], "hash_gpu_ms": best["gpu_ms"],
        "parse_gbs": best["witness_bytes"] / sec / 1e9 if sec else None,
        "parse_hbm_frac": best["witness_bytes"] / sec / 1e9 / peaks["hbm_gbs"] if sec else None,
        "instructions_per_sec": best["witness_instructions"] / sec if sec else None,
        "block_wall_ms": best["wall_ms"], "kernel_launches": best["kernel_launches"],
        "h2d_bytes": best["h2d_bytes"], "d2h_bytes": best["d2h_bytes"], **rates,
    }))


if __name__ == "__main__":
    main()
