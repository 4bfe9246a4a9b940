"""The device txn loop (csrc/txn_core.h: the code ppd_txn.cu's txn_loop_kernel instantiates) checked on the CPU.

tests/cpp/txn_core_check.cpp runs it as a thread block of one thread on host memory and compares, txn by txn and
structurally, with the host form of the same loop (csrc/host_txn.cu), which the GPU parity tests pin to the oracle:
the tries every subset is cut from, every storage trie in hashed-address order, the set of nodes the marking walks
touch, the tries after the txn, and the dummy / withdrawal entries.  Test infrastructure only: the binary is built with
-DPPD_HOSTPROF (no GPU, no node hashing) and is not part of libppd_b200.so."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "proof_protocol_decoder_b200", "csrc")


@pytest.fixture(scope="module")
def txncheck():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc is needed to build the check harness")
    res = subprocess.run(["make", "-C", CSRC, "txncheck"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return os.path.join(CSRC, "build", "txncheck")


def _run(binary, tmp_path, blocks):
    paths = []
    for i, blk in enumerate(blocks):
        p = tmp_path / f"b{i}.flat"
        p.write_bytes(blk.flat)
        paths.append(str(p))
    res = subprocess.run([binary] + paths, capture_output=True, text=True)
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert res.returncode == 0, "\n".join(lines[-20:]) + res.stderr[-2000:]
    assert len(lines) == len(blocks) and all(ln.endswith("identical") for ln in lines), "\n".join(lines[-20:])


def test_device_txn_loop_matches_host_loop_on_random_blocks(txncheck, tmp_path):
    from proof_protocol_decoder_b200 import synth

    rng = np.random.default_rng(11)
    blocks = []
    for i in range(40):
        vd = 0 if i % 3 else int(rng.integers(1, 4))  # a virtual state (hashed-out siblings) every third block
        blocks.append(
            synth.gen_block(
                2000 + i,
                n_accounts=int(rng.integers(2, 60)),
                n_txns=int(rng.integers(2, 25)),
                contract_frac=float(rng.uniform(0.2, 1.0)),
                slots_hi=int(rng.integers(1, 30)),
                accounts_per_txn=(1, int(rng.integers(2, 20))),
                slot_reads=(0, int(rng.integers(1, 10))),
                slot_writes=(0, int(rng.integers(1, 16))),
                zero_write_frac=float(rng.uniform(0, 0.7)),  # deletes: branches collapse, extensions merge
                virtual_depth=vd,
                allow_new_accounts=(vd == 0),
            )
        )
    _run(txncheck, tmp_path, blocks)


def test_device_txn_loop_dummies_and_withdrawals(txncheck, tmp_path):
    """pad_gen_inputs_with_dummy_inputs_if_needed / add_withdrawals_to_txns (decoding.rs:304-402): every combination
    of 0 / 1 / 2 / 5 txns with 0 / 1 / 3 withdrawals."""
    from proof_protocol_decoder_b200 import synth

    rng = np.random.default_rng(12)
    blocks = []
    for n_txns in (0, 1, 2, 5):
        for wd in (0, 1, 3):
            for _ in range(2):
                blocks.append(
                    synth.gen_block(
                        3000 + len(blocks),
                        n_accounts=int(rng.integers(2, 60)),
                        n_txns=n_txns,
                        contract_frac=float(rng.uniform(0.2, 1.0)),
                        slots_hi=int(rng.integers(1, 30)),
                        accounts_per_txn=(1, int(rng.integers(2, 20))),
                        slot_writes=(0, int(rng.integers(1, 16))),
                        zero_write_frac=float(rng.uniform(0, 0.7)),
                        n_withdrawals=wd,
                    )
                )
    _run(txncheck, tmp_path, blocks)


def test_device_txn_loop_mainnet_shaped(txncheck, tmp_path):
    from proof_protocol_decoder_b200 import synth

    blk = synth.gen_block(4, n_accounts=2000, n_txns=20, contract_frac=0.15, slots_hi=256, virtual_depth=7, accounts_per_txn=(30, 60),
                          slot_reads=(0, 3), slot_writes=(0, 3), allow_new_accounts=False, allow_self_destruct=False, inline_code_frac=0.02)
    _run(txncheck, tmp_path, [blk, synth.gen_block(1, n_accounts=1000, n_txns=10, n_withdrawals=2)])


def test_device_txn_loop_shortened_slot_keys_and_read_write_overlap(txncheck, tmp_path):
    """decoding.rs:235 hashes Nibbles::bytes_be() of a written slot key, which drops leading zero bytes, while the
    subset is cut with the hash of the full key (processed_block_trace.rs:234): such a write walks two keys.  And a
    slot may be both read and written by one txn: the same key twice in one batch."""
    from proof_protocol_decoder_b200 import synth

    rng = np.random.default_rng(13)
    blocks = []
    for i in range(6):
        blk = synth.gen_block(7000 + i, n_accounts=40, n_txns=6, contract_frac=0.8, slots_hi=12, accounts_per_txn=(3, 12), slot_reads=(1, 6), slot_writes=(1, 8))
        for tx in blk.txns:
            for _, tr in tx["traces"]:
                if tr.get("storage_written"):
                    w = list(tr["storage_written"])
                    # new slots whose raw key starts with zero bytes
                    for _ in range(int(rng.integers(1, 3))):
                        z = int(rng.integers(1, 30))
                        key = bytes(z) + bytes([1 + int(rng.integers(0, 255))]) + rng.bytes(31 - z)
                        w.append((key, int(rng.integers(1, 1 << 60))))
                    tr["storage_written"] = w
                    # read what is written, too
                    tr["storage_read"] = list(tr.get("storage_read") or []) + [k for k, _ in tr["storage_written"][:2]]
        blocks.append(blk)
    _run(txncheck, tmp_path, blocks)


def _error_blocks():
    """Blocks whose decoding ends in a TraceParsingError variant that carries a payload (decoding.rs:31-49)."""
    import copy

    from proof_protocol_decoder_b200 import synth

    def base(**kw):
        return synth.gen_block(55, n_accounts=80, n_txns=3, n_withdrawals=1, **kw)

    cases = []
    b = base()
    b.withdrawals = [(bytes(range(20)), 5)]  # MissingWithdrawalAccount(addr, hashed addr, amount)
    cases.append(("withdrawal_to_missing_account", 25, b.flat))
    b = base(virtual_depth=3, virtual_accounts_log16=3)
    b.txns = copy.deepcopy(b.txns)
    b.txns[0]["traces"].append((bytes([7] * 20), {"balance": 1}))  # MissingKeysCreatingSubPartialTrie(State)
    cases.append(("touched_account_behind_hash_node", 24, b.flat))
    # MissingKeysCreatingSubPartialTrie(Storage): the storage trie of an account a txn reads a slot of, with the node on
    # that slot's path hashed out (the tries as direct nodes from the oracle, re-spelled as a witness by the product's
    # host-only ppd_direct_to_compact)
    import ppd_oracle_lib
    from proof_protocol_decoder_b200 import flat, lib

    orc = ppd_oracle_lib.load()
    b = synth.gen_block(55, n_accounts=80, n_txns=3, n_withdrawals=1, inline_code_frac=0.0, contract_frac=0.5, slots_hi=40)
    state, storage = flat.parse_direct_pre_image(orc.compact_to_direct(flat.pre_image_of(b.flat)[1]))
    addr, tr = next((a, t) for a, t in b.txns[1]["traces"] if t.get("storage_read"))
    haddr, hslot = orc.keccak256(addr), orc.keccak256(tr["storage_read"][0])
    root = storage[haddr]
    if root[0] == "branch":
        ch = list(root[1])
        ch[hslot[0] >> 4] = ("hash", bytes([0x5A]) * 32)
        storage[haddr] = ("branch", ch, root[2])
    else:
        storage[haddr] = ("hash", bytes([0x5A]) * 32)
    wit = lib.load_library().direct_to_compact(flat.encode_direct_pre_image(state, storage))
    cases.append(("storage_slot_behind_hash_node", 24, flat.with_pre_image(b.flat, flat.PRE_IMAGE_COMBINED, wit)))
    # an address whose Keccak begins with a zero byte: H256::from_slice(&bytes_be()) panics (decoding.rs:202) after the
    # state / txn / receipt subsets and before any storage subset; no payload
    short = next(a for a in (i.to_bytes(20, "big") for i in range(1, 100000)) if orc.keccak256(a)[0] == 0)
    b = base()
    b.txns = copy.deepcopy(b.txns)
    b.txns[1]["traces"].append((short, {"balance": 1}))
    cases.append(("hashed_address_with_a_leading_zero_byte", 42, b.flat))
    # two faults: txn 0 fails in the loop (24), txn 2's receipt does not decode (43).  Every TxnInfo is processed before the
    # loop starts (processed_block_trace.rs:58-66), so the receipt panic is what the reference reports
    b = base(virtual_depth=3, virtual_accounts_log16=3)
    b.txns = copy.deepcopy(b.txns)
    b.txns[0]["traces"].append((bytes([7] * 20), {"balance": 1}))
    b.txns[2]["new_receipt_trie_node_byte"] = b"\xc1\x80"
    cases.append(("loop_error_in_txn_0_and_bad_receipt_in_txn_2", 43, b.flat))
    # likewise a code hash nobody resolves in txn 1 (this library's own status 61: the resolver callback of the reference)
    b = base(virtual_depth=3, virtual_accounts_log16=3)
    b.txns = copy.deepcopy(b.txns)
    b.txns[0]["traces"].append((bytes([7] * 20), {"balance": 1}))
    addr, tr = b.txns[1]["traces"][0]
    tr = dict(tr)
    tr.pop("code_write", None)
    tr["code_read"] = bytes([0xAB] * 32)
    b.txns[1]["traces"][0] = (addr, tr)
    cases.append(("loop_error_in_txn_0_and_unresolvable_code_in_txn_1", 61, b.flat))
    return cases


@pytest.mark.parametrize("name,code,flat_block", _error_blocks(), ids=[c[0] for c in _error_blocks()])
def test_host_path_error_payloads_equal_the_oracles(txncheck, oracle, tmp_path, name, code, flat_block):
    """ppd_last_error's wording for the TraceParsingError statuses carries the variant's payload (csrc/err_detail.h) so
    that the Rust shim can rebuild the error value: the host path of the product (which reports every block error) and
    the oracle must spell it the same way."""
    import re

    from ppd_oracle_lib import OracleError

    with pytest.raises(OracleError) as eo:
        oracle.block_decode(flat_block)
    assert eo.value.code == code
    p = tmp_path / "b.flat"
    p.write_bytes(flat_block)
    res = subprocess.run([txncheck, str(p)], capture_output=True, text=True)
    m = re.search(r": status (\d+) \((.*)\) from the host path; the device path (declines|raises flag)", res.stdout)
    assert m, res.stdout[-2000:] + res.stderr[-2000:]  # (a MISMATCH line: the device loop finished a failing block)
    assert int(m.group(1)) == code
    assert m.group(2) == eo.value.msg
    if code not in (24, 25):
        assert "; " not in m.group(2)
        return
    sentence, payload = m.group(2).split("; ")
    words = dict(w.split("=") for w in payload.split(" "))
    if code == 25:
        assert words["addr"] == bytes(range(20)).hex() and int(words["amount"], 16) == 5
        assert bytes.fromhex(words["hashed_addr"]) == oracle.keccak256(bytes(range(20)))
    else:
        assert words == {"trie_type": "Storage" if name.startswith("storage_slot") else "State"}


def test_host_path_reports_the_fault_the_oracle_reports_when_a_block_has_several(txncheck, oracle, tmp_path):
    """Blocks with one to three injected faults of different kinds at random txns: which of them is reported is a matter
    of the reference's ORDER of work (every TxnInfo processed first, then per txn the state / txn / receipt subsets, the
    H256::from_slice of every trace, the storage subsets, the deltas; withdrawals last).  The product's host path (which
    reports every block error) and the oracle must pick the same one, with the same words."""
    import copy
    import re

    from ppd_oracle_lib import OracleError
    from proof_protocol_decoder_b200 import synth

    short = next(a for a in (i.to_bytes(20, "big") for i in range(1, 100000)) if oracle.keccak256(a)[0] == 0)
    rng = np.random.default_rng(5)
    kinds = ["behind_hash", "bad_receipt", "unresolvable_code", "short_haddr", "missing_withdrawal"]
    paths, want = [], []
    for case in range(int(os.environ.get("PPD_FAULT_CASES", "36"))):
        n_txns = int(rng.integers(2, 6))
        b = synth.gen_block(700 + case, n_accounts=60, n_txns=n_txns, n_withdrawals=1, virtual_depth=3, virtual_accounts_log16=3, allow_new_accounts=False)
        b.txns = copy.deepcopy(b.txns)
        faults = rng.choice(kinds, size=int(rng.integers(1, 4)), replace=False)
        for k in faults:
            t = int(rng.integers(0, n_txns))
            if k == "behind_hash":
                b.txns[t]["traces"].append((bytes([7] * 20), {"balance": 1}))
            elif k == "bad_receipt":
                b.txns[t]["new_receipt_trie_node_byte"] = b"\xc1\x80"
            elif k == "unresolvable_code":
                b.txns[t]["traces"].append((bytes([9] * 19 + [t]), {"code_read": bytes([0xAB] * 32)}))
            elif k == "short_haddr":
                b.txns[t]["traces"].append((short, {"balance": 1}))
            else:
                b.withdrawals = [(bytes(range(20)), 5)]
        with pytest.raises(OracleError) as eo:
            oracle.block_decode(b.flat)
        p = tmp_path / f"f{case}.flat"
        p.write_bytes(b.flat)
        paths.append(str(p))
        want.append((eo.value.code, eo.value.msg, sorted(faults)))
    res = subprocess.run([txncheck] + paths, capture_output=True, text=True)
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == len(paths), res.stdout[-2000:] + res.stderr[-2000:]
    seen = set()
    for ln, (code, msg, faults) in zip(lines, want):
        # "... from the host path; the device path declines ... / raises flag ...": the device side of the harness ran too
        # and handed the block over (a block it finished would print MISMATCH)
        m = re.search(r": status (\d+) \((.*)\) from the host path; the device path (declines|raises flag)", ln)
        assert m, (ln, faults)
        assert (int(m.group(1)), m.group(2)) == (code, msg), (ln, code, msg, faults)
        seen.add(code)
    assert len(seen) >= 4, seen  # the draw must exercise most kinds as the winning fault


def _witness_like_storage(oracle, blk, rng):
    """The block with every storage trie cut down to what a real witness carries: the nodes on the paths of the slots
    the block's txns read or write, every other subtree hashed out (the generator witnesses storage tries in full).
    The tries are taken as direct nodes from the oracle, pruned, and re-spelled as a compact witness by the product's
    host-only ppd_direct_to_compact.  Returns (FlatBlock, number of subtrees hashed out)."""
    from proof_protocol_decoder_b200 import flat, lib

    def nibs_of(h):
        return [x for byte in h for x in (byte >> 4, byte & 15)]

    hashed = [0]

    def prune(node, keys):
        if not keys:
            if node[0] in ("empty", "hash"):
                return node
            hashed[0] += 1
            return ("hash", bytes(rng.bytes(32)))
        if node[0] == "branch":
            return ("branch", [prune(c, [k[1:] for k in keys if k and k[0] == i]) for i, c in enumerate(node[1])], node[2])
        if node[0] == "extension":
            n = len(node[1])
            return ("extension", node[1], prune(node[2], [k[n:] for k in keys if k[:n] == list(node[1])]))
        return node

    state, storage = flat.parse_direct_pre_image(oracle.compact_to_direct(flat.pre_image_of(blk.flat)[1]))
    accessed = {}
    for t in blk.txns:
        for addr, tr in t["traces"]:
            ks = accessed.setdefault(oracle.keccak256(addr), [])
            ks += [nibs_of(oracle.keccak256(s)) for s in tr.get("storage_read") or []]
            sw = tr.get("storage_written") or {}
            ks += [nibs_of(oracle.keccak256(s)) for s in (sw.keys() if isinstance(sw, dict) else [x[0] for x in sw])]
    for h, trie in list(storage.items()):
        if accessed.get(h):
            storage[h] = prune(trie, accessed[h])
    wit = lib.load_library().direct_to_compact(flat.encode_direct_pre_image(state, storage))
    return flat.with_pre_image(blk.flat, flat.PRE_IMAGE_COMBINED, wit), hashed[0]


def test_device_txn_loop_on_witness_like_storage_tries(txncheck, oracle, tmp_path):
    """Storage tries as mainnet witnesses carry them — hashed siblings next to every touched path — with 60 % of the
    writes zero (deletes: a branch that loses a child next to hashed-out siblings, extensions that merge): the oracle
    decodes the block, and the device loop builds the same tries as the host loop, txn by txn."""
    from proof_protocol_decoder_b200 import synth

    rng = np.random.default_rng(1)
    paths, total = [], 0
    for i in range(int(os.environ.get("PPD_WITNESS_LIKE_CASES", "16"))):
        b = synth.gen_block(900 + i, n_accounts=40, n_txns=int(rng.integers(2, 8)), inline_code_frac=0.0, contract_frac=0.7, slots_lo=4, slots_hi=60,
                            slot_reads=(0, 4), slot_writes=(0, 6), zero_write_frac=0.6, accounts_per_txn=(2, 6))
        fb, n_hashed = _witness_like_storage(oracle, b, rng)
        total += n_hashed
        oracle.block_decode(fb)  # no error: every touched path is witnessed
        p = tmp_path / f"w{i}.flat"
        p.write_bytes(fb)
        paths.append(str(p))
    assert total > 20 * len(paths), total  # the pruning did hash subtrees out
    res = subprocess.run([txncheck] + paths, capture_output=True, text=True)
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert res.returncode == 0 and len(lines) == len(paths) and all(ln.endswith("identical") for ln in lines), "\n".join(lines[-10:]) + res.stderr[-2000:]
