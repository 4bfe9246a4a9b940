"""The JSON wire format of BlockTrace (proof_protocol_decoder_b200/wire.py) against the reference's serde schema
(protocol_decoder/src/trace_protocol.rs:40-205, deserializers.rs:8-79).  CPU only."""
import json

import pytest

from proof_protocol_decoder_b200 import synth, wire
from proof_protocol_decoder_b200.trace_protocol import BlockTrace, ContractCodeUsage, TxnInfo, TxnMeta, TxnTrace


def test_byte_string_forms():
    # deserializers.rs:31-38: the prefix is optional and may be upper case; output is always 0x + lower case
    assert wire.byte_string_from_json("0xdeadBEEF") == bytes.fromhex("deadbeef")
    assert wire.byte_string_from_json("0Xdeadbeef") == bytes.fromhex("deadbeef")
    assert wire.byte_string_from_json("deadbeef") == bytes.fromhex("deadbeef")
    assert wire.byte_string_from_json("0x") == b""
    assert wire.byte_string_to_json(bytes.fromhex("00ff10")) == "0x00ff10"
    with pytest.raises(wire.WireFormatError, match="Odd number"):
        wire.byte_string_from_json("0xabc")
    with pytest.raises(wire.WireFormatError, match="Invalid character"):
        wire.byte_string_from_json("0xzz")
    # the reference slices data[..2] before looking at it: a string shorter than two bytes panics
    for s in ("", "a"):
        with pytest.raises(wire.WireFormatPanic):
            wire.byte_string_from_json(s)
    with pytest.raises(wire.WireFormatError):
        wire.byte_string_from_json(17)


def test_fixed_hashes_and_u256():
    assert wire.address_from_json("0x" + "ab" * 20) == bytes.fromhex("ab" * 20)
    for bad in ("ab" * 20, "0x" + "ab" * 19, "0x" + "ab" * 21, "0x" + "zz" * 20, 5):
        with pytest.raises(wire.WireFormatError):
            wire.address_from_json(bad)
    assert wire.h256_from_json("0x" + "01" * 32) == bytes.fromhex("01" * 32)
    assert wire.u256_from_json("0x0") == 0 and wire.u256_from_json("0x") == 0
    assert wire.u256_from_json("0x1f") == 31 and wire.u256_from_json("0x001f") == 31 and wire.u256_from_json("0xf" * 1 + "f" * 63) == (1 << 256) - 1
    assert wire.u256_to_json(0) == "0x0" and wire.u256_to_json(31) == "0x1f"
    for bad in ("1f", "0x" + "f" * 65, "0xg", 31):
        with pytest.raises(wire.WireFormatError):
            wire.u256_from_json(bad)


def test_schema_of_a_hand_written_trace():
    text = json.dumps({
        "trie_pre_images": {"combined": {"compact": "0x01"}},
        "txn_info": [{
            "traces": {
                "0x" + "11" * 20: {"balance": "0xde0b6b3a7640000", "nonce": "0x1"},
                "0x" + "22" * 20: {
                    "storage_read": ["0x" + "00" * 31 + "05"],
                    "storage_written": {"0x" + "00" * 31 + "06": "0x2a"},
                    "code_usage": {"read": "0x" + "cc" * 32},
                },
                "0x" + "33" * 20: {"code_usage": {"write": "0x6080"}, "self_destructed": True},
                "0x" + "44" * 20: {},
            },
            "meta": {"byte_code": "0xf86c", "new_txn_trie_node_byte": "f86c", "new_receipt_trie_node_byte": "0xf901", "gas_used": 21000},
        }],
    })
    bt = BlockTrace.from_json(text)
    assert bt.trie_pre_images == {"combined": {"compact": b"\x01"}}
    tr = bt.txn_info[0].traces
    assert tr[bytes.fromhex("11" * 20)] == TxnTrace(balance=10**18, nonce=1)
    assert tr[bytes.fromhex("22" * 20)] == TxnTrace(storage_read=[bytes(31) + b"\x05"], storage_written={bytes(31) + b"\x06": 42},
                                                     code_usage=ContractCodeUsage(read=bytes.fromhex("cc" * 32)))
    assert tr[bytes.fromhex("33" * 20)] == TxnTrace(code_usage=ContractCodeUsage(write=bytes.fromhex("6080")), self_destructed=True)
    assert tr[bytes.fromhex("44" * 20)] == TxnTrace()
    assert bt.txn_info[0].meta == TxnMeta(byte_code=bytes.fromhex("f86c"), new_txn_trie_node_byte=bytes.fromhex("f86c"),
                                          new_receipt_trie_node_byte=bytes.fromhex("f901"), gas_used=21000)
    # Option::is_none fields are skipped on output; byte strings come back 0x-prefixed
    out = json.loads(bt.to_json())
    assert out["txn_info"][0]["traces"]["0x" + "44" * 20] == {}
    assert out["txn_info"][0]["meta"]["new_txn_trie_node_byte"] == "0xf86c"
    assert BlockTrace.from_json(out) == bt


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_generated_blocks_round_trip_to_the_same_flat_block(seed):
    """BlockTrace -> JSON -> BlockTrace -> FlatBlock gives the bytes the generator's own FlatBlock has: the ingest step
    feeds the C ABI exactly what the direct path does."""
    blk = synth.gen_block(seed, n_accounts=300, n_txns=5, n_withdrawals=seed % 2)
    bt, meta, other = blk.to_block_trace()
    again = BlockTrace.from_json(bt.to_json())
    assert again == bt
    assert again.to_flat(meta, other) == bt.to_flat(meta, other)


def test_enum_tags_and_missing_fields():
    ok = {"trie_pre_images": {"combined": {"compact": "0x01"}}, "txn_info": []}
    assert BlockTrace.from_json(ok).txn_info == []
    for bad in (
        {"txn_info": []},                                                              # missing field
        {"trie_pre_images": {"Combined": {"compact": "0x01"}}, "txn_info": []},       # tags are snake_case
        {"trie_pre_images": {"combined": {}}, "txn_info": []},
        {"trie_pre_images": {"combined": {"compact": "0x01"}, "separate": {}}, "txn_info": []},
        {"trie_pre_images": {"combined": {"compact": "0x01"}}, "txn_info": {}},
        {"trie_pre_images": {"combined": {"compact": "0x01"}}, "txn_info": [{"traces": {}, "meta": {"byte_code": "0x", "new_txn_trie_node_byte": "0x",
                                                                                                      "new_receipt_trie_node_byte": "0x"}}]},
        {"trie_pre_images": {"combined": {"compact": "0x01"}}, "txn_info": [{"traces": {"0x" + "11" * 20: {"code_usage": {"exec": "0x"}}},
                                                                                  "meta": {"byte_code": "0x", "new_txn_trie_node_byte": "0x", "new_receipt_trie_node_byte": "0x", "gas_used": 1}}]},
    ):
        with pytest.raises(wire.WireFormatError):
            BlockTrace.from_json(bad)
    with pytest.raises(wire.WireFormatError, match="invalid JSON"):
        BlockTrace.from_json("{")
    # the `separate` variant parses (the schema has it) but has no decode path in the reference (todo!())
    sep = BlockTrace.from_json({"trie_pre_images": {"separate": {"state": {"uncompressed": {}}, "storage": {"single_trie": {}}}}, "txn_info": []})
    assert "separate" in sep.trie_pre_images
    from proof_protocol_decoder_b200.lib import PpdError
    from proof_protocol_decoder_b200.trace_protocol import OtherBlockData, ProcessingMeta

    with pytest.raises(PpdError) as ei:
        sep.to_flat(ProcessingMeta(lambda h: b""), OtherBlockData())
    assert ei.value.code == 45


def test_random_traces_round_trip():
    """hypothesis: any TxnTrace / TxnMeta the dataclasses can hold survives to_json -> from_json unchanged."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    h256 = st.binary(min_size=32, max_size=32)
    u256 = st.integers(min_value=0, max_value=(1 << 256) - 1)
    usage = st.one_of(st.builds(lambda h: ContractCodeUsage(read=h), h256), st.builds(lambda b: ContractCodeUsage(write=b), st.binary(max_size=64)))
    trace = st.builds(
        TxnTrace, balance=st.none() | u256, nonce=st.none() | u256, storage_read=st.none() | st.lists(h256, max_size=4),
        storage_written=st.none() | st.dictionaries(h256, u256, max_size=4), code_usage=st.none() | usage, self_destructed=st.none() | st.booleans(),
    )
    meta = st.builds(TxnMeta, byte_code=st.binary(max_size=80), new_txn_trie_node_byte=st.binary(max_size=80),
                     new_receipt_trie_node_byte=st.binary(max_size=80), gas_used=st.integers(min_value=0, max_value=(1 << 64) - 1))
    info = st.builds(TxnInfo, traces=st.dictionaries(st.binary(min_size=20, max_size=20), trace, max_size=4), meta=meta)

    @settings(max_examples=150, deadline=None)
    @given(st.binary(min_size=1, max_size=200), st.lists(info, max_size=3))
    def check(compact, infos):
        bt = BlockTrace(trie_pre_images={"combined": {"compact": compact}}, txn_info=infos)
        assert BlockTrace.from_json(bt.to_json()) == bt

    check()


def test_direct_pre_image_json_round_trip_unpinned():
    """Separate{Direct, MultipleTries{Direct}} through the JSON form (the serde shape of HashedPartialTrie is recalled,
    not pinned: eth_trie_utils is not under /root/reference) into a kind-2 FlatBlock."""
    from proof_protocol_decoder_b200 import flat
    from proof_protocol_decoder_b200.trace_protocol import OtherBlockData, ProcessingMeta

    leaf = ("leaf", [1, 2, 3], b"\x05")
    state = ("branch", [("hash", bytes([7]) * 32), ("extension", [0xa, 0xb], ("branch", [leaf] + [("empty",)] * 14 + [leaf], b""))] + [("empty",)] * 14, b"")
    bt = BlockTrace(trie_pre_images={"separate": {"state": {"direct": state}, "storage": {"multiple_tries": {bytes([9]) * 32: {"direct": leaf}}}}}, txn_info=[])
    again = BlockTrace.from_json(bt.to_json())
    assert again == bt
    f = again.to_flat(ProcessingMeta(lambda h: None), OtherBlockData())
    kind, payload = flat.pre_image_of(f)
    assert kind == flat.PRE_IMAGE_DIRECT and flat.parse_direct_pre_image(payload) == (state, {bytes([9]) * 32: leaf})
    assert wire.nibbles_from_json({"count": 3, "packed": "0x123"}) == [1, 2, 3]
    assert wire.nibbles_from_json({"count": 2, "packed": "0x3"}) == [0, 3]
    for bad in ({"count": 1, "packed": "0x12"}, {"count": 65, "packed": "0x0"}, {"count": 1}):
        with pytest.raises(wire.WireFormatError):
            wire.nibbles_from_json(bad)
    # a bare Node and hex byte vectors are accepted on input
    assert wire.direct_trie_from_json({"Leaf": {"nibbles": {"count": 1, "packed": "0xf"}, "value": "0x0102"}}) == ("leaf", [15], b"\x01\x02")
    assert wire.direct_trie_from_json("Empty") == ("empty",)
