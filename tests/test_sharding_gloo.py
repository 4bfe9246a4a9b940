"""The N>1 host logic (block sharding, trie bin-packing, the all-gather of 32-byte roots) on CPU:
two processes over gloo, with the oracle standing in for the GPU (it is only the checker here:
what is tested is shard.py)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import ppd_oracle_lib
    from proof_protocol_decoder_b200 import flat, shard, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = ppd_oracle_lib.load()
    # config 3 in miniature: five storage tries of very different sizes
    sizes = [3000, 10, 1200, 800, 1]
    tries = [synth.gen_sorted_leaves(n, seed=50 + i, val_lo=1, val_hi=33) for i, n in enumerate(sizes)]
    calls = []

    def root_fn(i):
        calls.append(i)
        return oracle.trie_root_from_leaves(*tries[i])

    roots = shard.sharded_trie_roots(root_fn, sizes, dist)
    # config 4 in miniature: five blocks, block i on rank i mod 2
    blocks = [synth.gen_block(900 + i, n_accounts=60, n_txns=2) for i in range(5)]
    decoded = []

    def decode_fn(i):
        decoded.append(i)
        irs = flat.parse_ir_dump(oracle.block_decode(blocks[i].flat))
        r = irs[-1]["trie_roots_after"]
        return [r["state_root"], r["transactions_root"], r["receipts_root"]]

    broots = shard.sharded_block_roots(decode_fn, len(blocks), dist)
    q.put((rank, calls, [r.hex() for r in roots], decoded, [[x.hex() for x in b] for b in broots]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather_roots(oracle):
    import torch.multiprocessing as mp

    from proof_protocol_decoder_b200 import flat, shard, synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, calls0, roots0, dec0, b0), (r1, calls1, roots1, dec1, b1) = res
    # every trie hashed exactly once, by the rank the packing names; both ranks hold all roots
    assert sorted(calls0 + calls1) == [0, 1, 2, 3, 4]
    assert [calls0, calls1] == shard.assign_tries([3000, 10, 1200, 800, 1], 2)
    assert roots0 == roots1
    sizes = [3000, 10, 1200, 800, 1]
    want = [oracle.trie_root_from_leaves(*synth.gen_sorted_leaves(n, seed=50 + i, val_lo=1, val_hi=33)).hex() for i, n in enumerate(sizes)]
    assert roots0 == want
    # blocks: i mod 2
    assert dec0 == [0, 2, 4] and dec1 == [1, 3]
    assert b0 == b1
    for i in range(5):
        irs = flat.parse_ir_dump(oracle.block_decode(synth.gen_block(900 + i, n_accounts=60, n_txns=2).flat))
        r = irs[-1]["trie_roots_after"]
        assert b0[i] == [r["state_root"].hex(), r["transactions_root"].hex(), r["receipts_root"].hex()]


def test_assign_tries_balances():
    from proof_protocol_decoder_b200 import shard

    sizes = [1_000_000] * 4 + [100] * 1000
    for world in (1, 2, 4, 8):
        a = shard.assign_tries(sizes, world)
        assert sorted(i for lst in a for i in lst) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in lst) for lst in a]
        assert max(loads) <= max(max(sizes), -(-sum(sizes) // world) + max(sizes))
    assert shard.shard_blocks(10, 1, 4) == [1, 5, 9]
