"""The txn loop on the GPU (csrc/ppd_txn.cu, gpu_txn.cu) against the oracle, and against the library's host txn loop.

Every block goes through the C ABI (ppd_block_decode).  With PPD_HOST_TXN unset the whole block stays on the device
(witness parse, by-root storage join, txn loop, sweeps, IR dump); with PPD_HOST_TXN=1 the host shapes the tries and
the device hashes them.  Both must give the oracle's bytes."""
import os

import numpy as np
import pytest

from witness_shapes import HDR, OP_ACCOUNT, branch, hashnode, leaf, nibs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from proof_protocol_decoder_b200.lib import Context

    c = Context(0)
    yield c
    c.close()


def _decode(ctx, flat_bytes, host_txn):
    from proof_protocol_decoder_b200.lib import PpdError

    saved = os.environ.pop("PPD_HOST_TXN", None)
    if host_txn:
        os.environ["PPD_HOST_TXN"] = "1"
    try:
        try:
            return ctx.block_decode(flat_bytes), ctx.stats()
        except PpdError as e:
            return e.code, ctx.stats()
    finally:
        os.environ.pop("PPD_HOST_TXN", None)
        if saved is not None:
            os.environ["PPD_HOST_TXN"] = saved


def _oracle(oracle, flat_bytes):
    from ppd_oracle_lib import OracleError

    try:
        return oracle.block_decode(flat_bytes)
    except OracleError as e:
        return e.code


def test_device_loop_takes_the_block_and_matches_oracle_and_host_loop(ctx, oracle):
    from proof_protocol_decoder_b200 import synth

    rng = np.random.default_rng(21)
    for i in range(24):
        vd = 0 if i % 3 else int(rng.integers(1, 4))
        blk = synth.gen_block(
            6000 + i, n_accounts=int(rng.integers(2, 300)), n_txns=int(rng.integers(0, 20)), contract_frac=float(rng.uniform(0.2, 1.0)),
            slots_hi=int(rng.integers(1, 40)), accounts_per_txn=(1, int(rng.integers(2, 30))), slot_reads=(0, int(rng.integers(1, 10))),
            slot_writes=(0, int(rng.integers(1, 16))), zero_write_frac=float(rng.uniform(0, 0.7)), virtual_depth=vd, allow_new_accounts=(vd == 0),
            n_withdrawals=int(rng.integers(0, 3)) if i % 2 else 0,
        )
        want = _oracle(oracle, blk.flat)
        got, st = _decode(ctx, blk.flat, host_txn=False)
        assert got == want, f"block {i}: the device txn loop differs from the oracle"
        if not isinstance(want, int):
            assert st["txn_loops_on_gpu"] == 1, f"block {i} did not take the device txn loop"
        got_host, st_host = _decode(ctx, blk.flat, host_txn=True)
        assert st_host["txn_loops_on_gpu"] == 0
        assert got_host == want, f"block {i}: the host txn loop differs from the oracle"


def test_shortened_slot_keys_and_read_write_overlap(ctx, oracle):
    """decoding.rs:235 hashes a written slot key without its leading zero bytes, while the subset is cut with the hash
    of the full key (processed_block_trace.rs:234); and a slot may be both read and written by one txn."""
    from proof_protocol_decoder_b200 import synth

    rng = np.random.default_rng(13)
    for i in range(6):
        blk = synth.gen_block(7000 + i, n_accounts=40, n_txns=6, contract_frac=0.8, slots_hi=12, accounts_per_txn=(3, 12), slot_reads=(1, 6), slot_writes=(1, 8))
        for tx in blk.txns:
            for _, tr in tx["traces"]:
                if tr.get("storage_written"):
                    w = list(tr["storage_written"])
                    for _ in range(int(rng.integers(1, 3))):
                        z = int(rng.integers(1, 30))
                        w.append((bytes(z) + bytes([1 + int(rng.integers(0, 255))]) + rng.bytes(31 - z), int(rng.integers(1, 1 << 60))))
                    tr["storage_written"] = w
                    tr["storage_read"] = list(tr.get("storage_read") or []) + [k for k, _ in w[:2]]
        want = _oracle(oracle, blk.flat)
        got, st = _decode(ctx, blk.flat, host_txn=False)
        assert got == want and st["txn_loops_on_gpu"] == 1
        got_host, _ = _decode(ctx, blk.flat, host_txn=True)
        assert got_host == want


def _account(key_nibbles, balance, storage_stream=None):
    from proof_protocol_decoder_b200.synth import cbor_bytes, compact_key

    flags = 8 | (2 if storage_stream is not None else 0)
    return (storage_stream or b"") + bytes([OP_ACCOUNT]) + cbor_bytes(compact_key(key_nibbles)) + bytes([flags]) + cbor_bytes(bytes([balance]))


@pytest.mark.parametrize("hashed_last", [True, False], ids=["hashed_form_witnessed_last", "expanded_form_witnessed_last"])
@pytest.mark.parametrize("host_txn", [False, True], ids=["device_loop", "host_loop"])
def test_storage_tries_are_joined_by_root_hash(ctx, oracle, hashed_last, host_txn):
    """compact_to_partial_trie.rs:167-190: storage tries are kept by ROOT HASH while the witness is processed (a later
    trie with the same root replaces an earlier one) and every account takes the trie stored under its storage root.
    Two accounts with the same storage, one witnessed expanded and one as a bare hash: both get whichever form came
    last, so a slot read of the account witnessed by hash succeeds or fails with the stream order."""
    from ppd_oracle_lib import parse_pre_image_dump
    from proof_protocol_decoder_b200 import flat, synth

    rnd = np.random.default_rng(5)
    # two accounts whose hashed addresses start with different nibbles (and not with a zero byte)
    addrs = []
    while len(addrs) < 2:
        a = rnd.bytes(20)
        h = synth.keccak256(a)
        if h[0] >= 0x10 and all(h[0] >> 4 != synth.keccak256(x)[0] >> 4 for x in addrs):
            addrs.append(a)
    addrs.sort(key=lambda a: synth.keccak256(a))
    h0, h1 = (synth.keccak256(a).hex() for a in addrs)
    # one storage trie of two slots under a branch
    slots = []
    while len(slots) < 2:
        k = bytes([1 + len(slots)]) + rnd.bytes(31)
        hk = synth.keccak256(k)
        if all(hk[0] >> 4 != synth.keccak256(x)[0] >> 4 for x in slots):
            slots.append(k)
    slots.sort(key=lambda k: synth.keccak256(k))
    s0, s1 = (synth.keccak256(k).hex() for k in slots)
    expanded = leaf(nibs(s0[1:]), b"\x2a") + leaf(nibs(s1[1:]), b"\x2b") + branch((1 << int(s0[0], 16)) | (1 << int(s1[0], 16)))
    probe = HDR + _account(nibs(h0), 1, expanded)
    root = list(parse_pre_image_dump(oracle.compact_decode(probe))["storage"].values())[0]
    forms = [hashnode(root), expanded] if not hashed_last else [expanded, hashnode(root)]
    witness = (HDR + _account(nibs(h0[1:]), 1, forms[0]) + _account(nibs(h1[1:]), 2, forms[1])
               + branch((1 << int(h0[0], 16)) | (1 << int(h1[0], 16))))
    # txn 0 reads a slot of each account; txn 1 writes one
    receipt = synth.legacy_receipt(1, 21000, 0, rnd)
    txns = [
        {"traces": [(addrs[0], {"storage_read": [slots[0]]}), (addrs[1], {"storage_read": [slots[1]], "balance": 9})],
         "byte_code": b"\x01" * 120, "new_receipt_trie_node_byte": receipt, "gas_used": 21000},
        {"traces": [(addrs[1], {"storage_written": [(slots[0], 77)], "nonce": 1})],
         "byte_code": b"\x02" * 120, "new_receipt_trie_node_byte": receipt, "gas_used": 21000},
    ]
    fb = flat.encode_flat_block(witness, txns, [], [], bytes(32), b"meta", b"hashes")
    want = _oracle(oracle, fb)
    # (with the hashed form last both accounts hold a hashed-out trie and the reads fail; the other way round both succeed)
    assert isinstance(want, int) == hashed_last
    got, st = _decode(ctx, fb, host_txn)
    assert got == want
