"""BASELINE.json's configs at FULL size on the GPU, against results the CPU oracle produced once in the build
container (tests/golden/gen_oracle_big.py -> tests/golden/oracle_big_roots.json; the oracle needs minutes for these, so
it is not run on the GPU box), plus size-independent properties where the oracle cannot reach (100 M leaves).

  C3  storage-heavy block: 4 contracts x 1 000 000 slots + 1 000 accounts, one txn writing 10 000 slots per contract
  C4  1 024 C1-shaped blocks through ppd_blocks_decode_batch
  C5  state-trie rehash over sorted leaves: 1 M and 10 M pinned; 100 M by the split-equals-whole property

Inputs are regenerated here from the same seeds (numpy Generator streams are reproducible); the goldens also pin the
SHA-256 of the regenerated FlatBlocks, so a generator drift is reported as such."""
import hashlib
import json
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def big():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_big_roots.json")))


@pytest.fixture(scope="module")
def ctx():
    from proof_protocol_decoder_b200.lib import Context

    c = Context(0)
    yield c
    c.close()


def _c4_flat(seed):
    from proof_protocol_decoder_b200 import synth

    return synth.gen_block(seed, n_accounts=1000, n_txns=10, n_withdrawals=seed % 3).flat


def test_c4_1024_blocks_match_pinned_oracle(ctx, big):
    seeds = sorted(int(s) for s in big["c4"]["ir_sha256"])
    assert len(seeds) == 1024
    with mp.get_context("fork").Pool(min(32, os.cpu_count() or 1)) as pool:
        flats = pool.map(_c4_flat, seeds)
    outs = ctx.blocks_decode_batch(flats)
    bad = [s for s, o in zip(seeds, outs) if isinstance(o, Exception) or hashlib.sha256(o).hexdigest() != big["c4"]["ir_sha256"][str(s)]]
    assert not bad, f"{len(bad)} of 1024 blocks differ from the oracle's IrDump, first seeds: {bad[:5]}"


@pytest.mark.parametrize("which", [0, 1], ids=["1M", "10M"])
def test_c5_sorted_leaves_root_matches_pinned_oracle(ctx, big, which):
    from proof_protocol_decoder_b200 import synth

    g = big["c5"][which]
    keys, val_off, vals = synth.gen_sorted_leaves(g["leaves"], seed=g["seed"])
    assert ctx.trie_root_sorted_leaves(keys, val_off, vals).hex() == g["root"]
    assert ctx.stats()["nodes_hashed"] == g["nodes_hashed"]


def _device_leaves(torch, n, seed, val_lo=70, val_hi=80):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    keys = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    k64 = torch.zeros(n, dtype=torch.int64, device="cuda")
    for b in range(8):
        k64 = (k64 << 8) | keys[:, b].to(torch.int64)
    order = torch.argsort((k64 >> 1) & 0x7FFFFFFFFFFFFFFF)  # top 63 bits, unsigned order
    keys = keys[order].contiguous()
    del k64, order
    lens = torch.randint(val_lo, val_hi + 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    val_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    val_off[1:] = torch.cumsum(lens, 0)
    vals = torch.randint(0, 256, (int(val_off[-1].item()),), dtype=torch.uint8, device="cuda", generator=g)
    return keys, val_off, vals


@pytest.mark.parametrize("n", [200_000, 100_000_000], ids=["200k", "100M"])
def test_c5_split_by_top_nibble_equals_whole_trie(ctx, n):
    """one huge trie split at its top nibble into 16 sub-tries that are hashed apart (on other GPUs in
    bench.py --gpus N), 16 refs gathered, the top branch hashed last.  Must give the whole trie's root (the 100 M-leaf
    case is config 5 at full size)."""
    import torch

    from proof_protocol_decoder_b200 import shard

    keys, val_off, vals = _device_leaves(torch, n, 5)
    whole = ctx.trie_root_sorted_leaves_dev(keys.data_ptr(), val_off.data_ptr(), vals.data_ptr(), n, int(val_off[-1].item()))
    refs, mask, _ = shard.split_trie_refs(ctx, torch, keys, val_off, vals)
    assert mask == 0xFFFF
    assert ctx.trie_root_from_children(b"".join(refs), mask) == whole
    del keys, val_off, vals
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["c3_small", "c3"], ids=["4x100k_slots", "4x1M_slots"])
def test_c3_storage_heavy_block_matches_pinned_oracle(ctx, big, name):
    from proof_protocol_decoder_b200 import synth

    g = big[name]
    blk = synth.gen_c3_block(**g["params"])
    fb = blk.flat
    assert hashlib.sha256(fb).hexdigest() == g["flat_sha256"], "the generator no longer reproduces the pinned input"
    with ctx.block_decode_view(fb) as v:
        got = hashlib.sha256(v.view).hexdigest()
        n = v.nbytes
    assert n == g["ir_bytes"] and got == g["ir_sha256"]
