"""Separate{Direct} pre-images (FlatBlock pre_image_kind 2; csrc/host_direct.cu, oracle process_direct_pre_image).

The reference takes a Direct state trie as it is (processed_block_trace.rs:143-148) and leaves the storage side as
todo!() (:164-168); kind 2 completes it with a Direct trie per hashed address.  There is no reference behaviour to
pin this to, so parity is anchored on the Combined path: the tries a compact witness decodes to, sent as a direct
pre-image, must give the same IRs as the witness itself (no inline code in the witness: a Separate pre-image carries
no code mappings, processed_block_trace.rs:139).  The product re-spells the direct tries as a witness on the host
and takes the witness path; the oracle builds its tries from the direct form without that detour."""
import struct

import pytest

import ppd_oracle_lib
from proof_protocol_decoder_b200 import flat, synth
from proof_protocol_decoder_b200.trace_protocol import BlockTrace, OtherBlockData, ProcessingMeta


@pytest.fixture(scope="module")
def oracle():
    return ppd_oracle_lib.load()


def _direct_block(oracle, blk_flat):
    kind, compact = flat.pre_image_of(blk_flat)
    assert kind == flat.PRE_IMAGE_COMBINED
    return flat.with_pre_image(blk_flat, flat.PRE_IMAGE_DIRECT, oracle.compact_to_direct(compact))


def _blocks():
    yield "c1", synth.gen_block(11, n_accounts=300, n_txns=6, inline_code_frac=0.0)
    yield "withdrawals", synth.gen_block(12, n_accounts=120, n_txns=3, n_withdrawals=2, inline_code_frac=0.0)
    yield "one_txn_dummy", synth.gen_block(13, n_accounts=80, n_txns=1, inline_code_frac=0.0)
    yield "no_txn", synth.gen_block(14, n_accounts=40, n_txns=0, n_withdrawals=1, inline_code_frac=0.0)


def test_oracle_direct_equals_combined(oracle):
    for name, blk in _blocks():
        want = oracle.block_decode(blk.flat)
        got = oracle.block_decode(_direct_block(oracle, blk.flat))
        assert got == want, name


def test_direct_payload_round_trips_through_the_python_codec(oracle):
    blk = synth.gen_block(15, n_accounts=60, n_txns=2, inline_code_frac=0.0)
    payload = oracle.compact_to_direct(flat.pre_image_of(blk.flat)[1])
    state, storage = flat.parse_direct_pre_image(payload)
    assert flat.encode_direct_pre_image(state, storage) == payload
    # the mirror's BlockTrace takes the tries as node tuples
    kind, compact = flat.pre_image_of(blk.flat)
    bt = BlockTrace(trie_pre_images={"separate": {"state": {"direct": state}, "storage": {"multiple_tries": {h: {"direct": t} for h, t in storage.items()}}}}, txn_info=[])
    f = bt.to_flat(ProcessingMeta(lambda h: None), OtherBlockData())
    assert flat.pre_image_of(f) == (flat.PRE_IMAGE_DIRECT, payload)


def test_other_separate_forms_stay_unimplemented(oracle):
    from proof_protocol_decoder_b200.lib import PpdError

    blk = synth.gen_block(16, n_accounts=20, n_txns=1, inline_code_frac=0.0)
    for k in (1, 3):
        f = bytearray(blk.flat)
        struct.pack_into("<I", f, 8, k)
        with pytest.raises(ppd_oracle_lib.OracleError) as e:
            oracle.block_decode(bytes(f))
        assert e.value.code == 45
    for pre in ({"separate": {"state": {"uncompressed": {}}, "storage": {"single_trie": {}}}}, {"separate": {"state": {"direct": ("empty",)}, "storage": {"single_trie": {}}}}):
        with pytest.raises(PpdError) as e:
            BlockTrace(trie_pre_images=pre, txn_info=[]).to_flat(ProcessingMeta(lambda h: None), OtherBlockData())
        assert e.value.code == 45


def test_malformed_direct_payloads_are_rejected_by_the_oracle(oracle):
    blk = synth.gen_block(17, n_accounts=30, n_txns=1, inline_code_frac=0.0)
    good = oracle.compact_to_direct(flat.pre_image_of(blk.flat)[1])
    for bad in (good[:-1], good + b"\x00", b"\x09", b""):
        with pytest.raises(ppd_oracle_lib.OracleError) as e:
            oracle.block_decode(flat.with_pre_image(blk.flat, flat.PRE_IMAGE_DIRECT, bad))
        assert e.value.code == 60, bad[:4]  # PPD_ERR_BAD_FLAT_INPUT


def test_product_transcoder_round_trips_on_the_cpu(oracle):
    """ppd_direct_to_compact is host code: direct tries -> witness (the product) -> tries (the oracle) is the identity,
    on the reference's golden witnesses and on drawn blocks."""
    import json
    import os

    from proof_protocol_decoder_b200.lib import load_library

    lib = load_library()
    root = os.path.dirname(os.path.abspath(__file__))
    goldens = json.load(open(os.path.join(root, "golden", "reference_goldens.json")))["compact_goldens"]
    witnesses = [bytes.fromhex(g["witness_hex"]) for g in goldens]
    witnesses += [flat.pre_image_of(blk.flat)[1] for _, blk in _blocks()]
    witnesses.append(flat.pre_image_of(synth.gen_block(21, n_accounts=400, n_txns=3, inline_code_frac=0.0).flat)[1])
    for k, w in enumerate(witnesses):
        direct = oracle.compact_to_direct(w)
        w2 = lib.direct_to_compact(direct)
        assert w2[0] == 1
        assert oracle.compact_to_direct(w2) == direct, k
        a, b = ppd_oracle_lib.parse_pre_image_dump(oracle.compact_decode(w)), ppd_oracle_lib.parse_pre_image_dump(oracle.compact_decode(w2))
        assert a["state_root"] == b["state_root"] and a["storage"] == b["storage"], k


def test_product_transcoder_rejects_what_it_cannot_spell():
    from proof_protocol_decoder_b200.lib import PpdError, load_library

    lib = load_library()
    leaf = lambda nib, v: ("leaf", nib, v)  # noqa: E731
    cases = {
        "state leaf that is not an account": (flat.encode_direct_pre_image(leaf([1] * 64, b"\x01"), {}), 44),
        "branch with a value": (flat.encode_direct_pre_image(("branch", [("hash", bytes(32))] * 16, b"\x01"), {}), 60),
        "storage trie of an unknown account": (flat.encode_direct_pre_image(("hash", bytes(32)), {bytes(32): ("empty",)}), 60),
        "truncated": (b"\x02", 60),
        "trailing bytes": (flat.encode_direct_pre_image(("empty",), {}) + b"\x00", 60),
    }
    for name, (payload, code) in cases.items():
        with pytest.raises(PpdError) as e:
            lib.direct_to_compact(payload)
        assert e.value.code == code, name
    assert lib.direct_to_compact(flat.encode_direct_pre_image(("empty",), {})) == bytes([1, 6])  # header, EMPTY_ROOT


# ---------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx():
    from proof_protocol_decoder_b200.lib import Context, PpdError

    try:
        c = Context(0)
    except PpdError as e:
        pytest.skip(f"no usable CUDA device: {e}")
    yield c
    c.close()


@pytest.mark.gpu
def test_product_direct_equals_oracle_and_combined(ctx, oracle, monkeypatch):
    monkeypatch.setenv("PPD_GPU_PARSE_MIN_BYTES", "0")
    monkeypatch.setenv("PPD_GPU_DUMP_MIN_TOUCHED", "0")
    for name, blk in _blocks():
        want = oracle.block_decode(blk.flat)
        direct = _direct_block(oracle, blk.flat)
        assert oracle.block_decode(direct) == want, name
        assert ctx.block_decode(direct) == want, name
        assert ctx.block_decode(blk.flat) == want, name
    # through the batch entry point, mixed with Combined blocks
    blks = [b for _, b in _blocks()]
    flats = [blks[0].flat, _direct_block(oracle, blks[1].flat), _direct_block(oracle, blks[0].flat), blks[1].flat]
    outs = ctx.blocks_decode_batch(flats)
    assert outs[0] == outs[2] == oracle.block_decode(blks[0].flat)
    assert outs[1] == outs[3] == oracle.block_decode(blks[1].flat)


@pytest.mark.gpu
def test_product_direct_host_builder_and_account_without_its_trie(ctx, oracle, monkeypatch):
    """The witness the direct tries are re-spelled as also goes through the host builder (PPD_HOST_PARSE); and an
    account whose storage trie is NOT sent keeps its root but has no trie (a slot access then fails as in the oracle)."""
    blk = synth.gen_block(18, n_accounts=150, n_txns=4, inline_code_frac=0.0)
    want = oracle.block_decode(blk.flat)
    direct = _direct_block(oracle, blk.flat)
    monkeypatch.setenv("PPD_HOST_PARSE", "1")
    assert ctx.block_decode(direct) == want
    monkeypatch.delenv("PPD_HOST_PARSE")
    monkeypatch.setenv("PPD_GPU_PARSE_MIN_BYTES", "0")
    # drop every storage trie: the product and the oracle must agree on the outcome, whatever it is
    state, storage = flat.parse_direct_pre_image(flat.pre_image_of(direct)[1])
    stripped = flat.with_pre_image(blk.flat, flat.PRE_IMAGE_DIRECT, flat.encode_direct_pre_image(state, {}))
    from proof_protocol_decoder_b200.lib import PpdError

    try:
        o = ("ok", oracle.block_decode(stripped))
    except ppd_oracle_lib.OracleError as e:
        o = ("err", e.code)
    try:
        p = ("ok", ctx.block_decode(stripped))
    except PpdError as e:
        p = ("err", e.code)
    assert p == o
    # half of them
    keep = dict(list(sorted(storage.items()))[::2])
    half = flat.with_pre_image(blk.flat, flat.PRE_IMAGE_DIRECT, flat.encode_direct_pre_image(state, keep))
    try:
        o = ("ok", oracle.block_decode(half))
    except ppd_oracle_lib.OracleError as e:
        o = ("err", e.code)
    try:
        p = ("ok", ctx.block_decode(half))
    except PpdError as e:
        p = ("err", e.code)
    assert p == o


# ---------------------------------------------------------------------------------------------- drawn tries (CPU)
def _rlp_str(b: bytes) -> bytes:
    if len(b) == 1 and b[0] < 0x80:
        return b
    if len(b) < 56:
        return bytes([0x80 + len(b)]) + b
    ln = len(b).to_bytes((len(b).bit_length() + 7) // 8, "big")
    return bytes([0xb7 + len(ln)]) + ln + b


def _rlp_list(items) -> bytes:
    body = b"".join(items)
    if len(body) < 56:
        return bytes([0xc0 + len(body)]) + body
    ln = len(body).to_bytes((len(body).bit_length() + 7) // 8, "big")
    return bytes([0xf7 + len(ln)]) + ln + body


def _int_be(v: int) -> bytes:
    return v.to_bytes((v.bit_length() + 7) // 8, "big")


EMPTY_TRIE_HASH = bytes.fromhex("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421")
EMPTY_CODE_HASH = bytes.fromhex("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470")


def _canonical_trie(items, depth=0):
    """The canonical trie of prefix-free (nibble list, payload) items that agree on their first `depth` nibbles; a payload
    is ("leaf", value bytes) or ("hash", h32) (a hashed-out subtree at that path)."""
    if not items:
        return ("empty",)
    if len(items) == 1:
        key, (kind, v) = items[0]
        rest = key[depth:]
        if kind == "hash":
            return ("extension", rest, ("hash", v)) if rest else ("hash", v)
        return ("leaf", rest, v)
    first = items[0][0]
    common = min(len(k) for k, _ in items) - depth
    for k, _ in items[1:]:
        c = 0
        while c < common and k[depth + c] == first[depth + c]:
            c += 1
        common = c
    if common:
        return ("extension", first[depth : depth + common], _canonical_trie(items, depth + common))
    ch = []
    for nib in range(16):
        ch.append(_canonical_trie([(k, p) for k, p in items if k[depth] == nib], depth + 1))
    return ("branch", ch, b"")


def test_drawn_direct_tries_round_trip_through_the_product_transcoder(oracle):
    """hypothesis: canonical state tries with accounts of every shape (no code / code, no storage / storage sent /
    storage withheld / storage sent as an empty trie, hashed-out siblings at every depth, short and long values) go
    direct -> witness (product) -> direct (oracle) unchanged, with every account's storage_root the root of its trie."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from proof_protocol_decoder_b200.lib import load_library

    lib = load_library()
    nib = st.integers(min_value=0, max_value=15)
    h32 = st.binary(min_size=32, max_size=32)

    def storage_items(draw):
        n = draw(st.integers(min_value=0, max_value=6))
        items, seen = [], set()
        for _ in range(n):
            key = tuple(draw(st.lists(nib, min_size=64, max_size=64)))
            cut = draw(st.integers(min_value=1, max_value=64))
            hashed = draw(st.booleans()) and cut < 64
            k = key[:cut] if hashed else key
            if any(k[: len(o)] == o or o[: len(k)] == k for o in seen):
                continue
            seen.add(k)
            val = _rlp_str(draw(st.binary(min_size=1, max_size=40).filter(lambda b: b[0] != 0 or len(b) > 1)))
            items.append((list(k), ("hash", draw(h32)) if hashed else ("leaf", val)))
        return sorted(items)

    @st.composite
    def pre_image(draw):
        n_acct = draw(st.integers(min_value=0, max_value=7))
        state_items, storage, seen = [], {}, set()
        for _ in range(n_acct):
            key = tuple(draw(st.lists(nib, min_size=64, max_size=64)))
            cut = draw(st.integers(min_value=1, max_value=64))
            hashed = draw(st.integers(min_value=0, max_value=3)) == 0 and cut < 64
            k = key[:cut] if hashed else key
            if any(k[: len(o)] == o or o[: len(k)] == k for o in seen):
                continue
            seen.add(k)
            if hashed:
                state_items.append((list(k), ("hash", draw(h32))))
                continue
            haddr = bytes((key[2 * i] << 4) | key[2 * i + 1] for i in range(32))
            mode = draw(st.sampled_from(["none", "sent", "withheld", "sent_empty"]))
            trie = None
            if mode == "sent":
                trie = _canonical_trie(storage_items(draw))
            elif mode == "sent_empty":
                trie = ("empty",)
            if trie is not None:
                storage[haddr] = trie
            nonce = draw(st.one_of(st.just(0), st.integers(min_value=1, max_value=(1 << 64) - 1)))
            balance = draw(st.one_of(st.just(0), st.integers(min_value=1, max_value=(1 << 256) - 1)))
            code_hash = draw(st.one_of(st.just(EMPTY_CODE_HASH), h32))
            state_items.append((list(key), ("acct", (nonce, balance, code_hash, mode, draw(h32)))))
        return sorted(state_items), storage

    @settings(max_examples=120, deadline=None)
    @given(pre_image())
    def check(pi):
        state_items, storage = pi
        # storage roots: the root of the trie sent (through the oracle), the drawn hash when withheld, else empty
        leaves = []
        for key, (kind, v) in state_items:
            if kind == "hash":
                leaves.append((key, (kind, v)))
                continue
            nonce, balance, code_hash, mode, withheld_root = v
            haddr = bytes((key[2 * i] << 4) | key[2 * i + 1] for i in range(32))
            if mode in ("sent", "sent_empty") and storage[haddr] != ("empty",):
                # root of that trie: a one-account state around it, through the product transcoder and the oracle
                probe = flat.encode_direct_pre_image(("leaf", key, _rlp_list([_rlp_str(b""), _rlp_str(b""), _rlp_str(bytes(32)), _rlp_str(EMPTY_CODE_HASH)])), {haddr: storage[haddr]})
                root = ppd_oracle_lib.parse_pre_image_dump(oracle.compact_decode(lib.direct_to_compact(probe)))["storage"].get(haddr)
                sroot = root if root is not None else EMPTY_TRIE_HASH
            elif mode == "withheld":
                sroot = withheld_root
            else:
                sroot = EMPTY_TRIE_HASH
            leaves.append((key, ("leaf", _rlp_list([_rlp_str(_int_be(nonce)), _rlp_str(_int_be(balance)), _rlp_str(sroot), _rlp_str(code_hash)]))))
        state = _canonical_trie(leaves)
        payload = flat.encode_direct_pre_image(state, storage)
        w = lib.direct_to_compact(payload)
        back_state, back_storage = flat.parse_direct_pre_image(oracle.compact_to_direct(w))
        assert back_state == state
        # the witness path joins by root: every sent trie comes back under its address (a withheld one as its bare hash,
        # which the product drops again after the pre-image is built: direct_filter_storage)
        for h, t in storage.items():
            if t != ("empty",):
                assert back_storage.get(h) == t

    check()


def test_direct_block_through_the_json_wire_form(oracle):
    """A Combined block's tries as Separate{Direct} JSON (wire.py) -> BlockTrace -> kind-2 FlatBlock: the oracle's IRs equal
    those of the Combined block (the txn part of the FlatBlock is carried over as it is)."""
    from proof_protocol_decoder_b200 import wire

    blk = synth.gen_block(31, n_accounts=90, n_txns=3, inline_code_frac=0.0)
    want = oracle.block_decode(blk.flat)
    state, storage = flat.parse_direct_pre_image(oracle.compact_to_direct(flat.pre_image_of(blk.flat)[1]))
    js = wire.pre_images_to_json({"separate": {"state": {"direct": state}, "storage": {"multiple_tries": {h: {"direct": t} for h, t in storage.items()}}}})
    import json

    pre = wire.pre_images_from_json(json.loads(json.dumps(js)))
    bt = BlockTrace(trie_pre_images=pre, txn_info=[])
    kind, payload = flat.pre_image_of(bt.to_flat(ProcessingMeta(lambda h: None), OtherBlockData()))
    assert kind == flat.PRE_IMAGE_DIRECT
    assert oracle.block_decode(flat.with_pre_image(blk.flat, kind, payload)) == want
