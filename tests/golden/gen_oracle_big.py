#!/usr/bin/env python3
"""Pins the BIG configs of BASELINE.json to the CPU oracle (oracle/ppd_oracle.cpp), once, here, so that the GPU tests
can compare the CUDA path at full size without running the oracle for minutes on the GPU box:

  C3  one storage-heavy block (4 contracts x 1 000 000 slots + 1 000 accounts, one txn writing 10 000 slots per contract),
      and its 4 x 100 000 variant: SHA-256 of the IrDump
  C4  1 024 C1-shaped blocks (seeds 1000..2023): SHA-256 of every IrDump
  C5  root of the trie over 1 000 000 and 10 000 000 sorted leaves (100 000 000 does not fit the time the oracle has
      here: about 30 000 node hashes per second on one core)

  python tests/golden/gen_oracle_big.py [c3|c3small|c4|c5 ...]      writes tests/golden/oracle_big_roots.json

The inputs come from proof_protocol_decoder_b200/synth.py (numpy Generator streams keyed by the seeds below), so the
GPU box regenerates byte-identical inputs."""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "oracle_big_roots.json")

C3_FULL = dict(seed=3, n_contracts=4, slots=1_000_000, n_plain=1000, writes_per_contract=10_000)
C3_SMALL = dict(seed=3, n_contracts=4, slots=100_000, n_plain=1000, writes_per_contract=10_000)
C4_SEEDS = list(range(1000, 2024))
C5_SIZES = [1_000_000, 10_000_000]


def c4_block(seed):
    from proof_protocol_decoder_b200 import synth

    return synth.gen_block(seed, n_accounts=1000, n_txns=10, n_withdrawals=seed % 3).flat


def _c4_job(seed):
    import ppd_oracle_lib

    return seed, hashlib.sha256(ppd_oracle_lib.load().block_decode(c4_block(seed))).hexdigest()


def _c3_job(params):
    import ppd_oracle_lib
    from proof_protocol_decoder_b200 import synth

    t0 = time.time()
    blk = synth.gen_c3_block(**params)
    fb = blk.flat
    t1 = time.time()
    o = ppd_oracle_lib.load()
    ir = o.block_decode(fb)
    st = o.last_stats()
    return {"params": params, "flat_sha256": hashlib.sha256(fb).hexdigest(), "flat_bytes": len(fb), "ir_sha256": hashlib.sha256(ir).hexdigest(),
            "ir_bytes": len(ir), "nodes_hashed": st["nodes_hashed"], "generate_s": t1 - t0, "oracle_s": time.time() - t1}


def _c5_job(n):
    import ppd_oracle_lib
    from proof_protocol_decoder_b200 import synth

    keys, val_off, vals = synth.gen_sorted_leaves(n, seed=5)
    t0 = time.time()
    o = ppd_oracle_lib.load()
    root = o.trie_root_from_leaves(keys, val_off, vals)
    return {"leaves": n, "seed": 5, "root": root.hex(), "nodes_hashed": o.last_stats()["nodes_hashed"], "oracle_s": time.time() - t0}


def main():
    what = sys.argv[1:] or ["c3small", "c4", "c5", "c3"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    if "c4" in what:
        with mp.get_context("fork").Pool(min(6, os.cpu_count() or 1)) as pool:
            res["c4"] = {"blocks": "synth.gen_block(seed, n_accounts=1000, n_txns=10, n_withdrawals=seed % 3)", "ir_sha256": {str(s): h for s, h in pool.map(_c4_job, C4_SEEDS)}}
    if "c3small" in what:
        res["c3_small"] = _c3_job(C3_SMALL)
    if "c5" in what:
        res["c5"] = [_c5_job(n) for n in C5_SIZES]
    if "c3" in what:
        res["c3"] = _c3_job(C3_FULL)
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps({k: (v if k != "c4" else len(v["ir_sha256"])) for k, v in res.items()})[:2000])


if __name__ == "__main__":
    main()
