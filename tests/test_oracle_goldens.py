"""Pins the CPU oracle against every fixture the reference's own tests hold for this path
(SURVEY.md 8c): six witness -> state-root goldens, the instruction-list KAT, three constants."""
import struct

import pytest

from ppd_oracle_lib import OracleError, parse_pre_image_dump


def test_constants_are_keccak_kats(oracle, goldens):
    c = goldens["constants"]
    assert oracle.keccak256(b"").hex() == c["EMPTY_CODE_HASH"]  # types.rs:24-28
    assert oracle.keccak256(b"\x80").hex() == c["EMPTY_TRIE_HASH"]  # types.rs:30-34


def test_empty_account_rlp_constant(oracle, goldens):
    # types.rs:36-41: rlp(AccountRlp{0, 0, EMPTY_TRIE_HASH, EMPTY_CODE_HASH})
    import ctypes

    c = goldens["constants"]
    out = ctypes.create_string_buffer(128)
    n = ctypes.c_size_t()
    z = bytes(32)
    oracle.L.oracle_rlp_account(z, z, bytes.fromhex(c["EMPTY_TRIE_HASH"]), bytes.fromhex(c["EMPTY_CODE_HASH"]), out, ctypes.byref(n))
    assert out.raw[: n.value].hex() == c["EMPTY_ACCOUNT_BYTES_RLPED"]
    assert n.value == 70


@pytest.mark.parametrize("idx", range(6))
def test_golden_state_roots(oracle, goldens, idx):
    g = goldens["compact_goldens"][idx]
    d = parse_pre_image_dump(oracle.compact_decode(bytes.fromhex(g["witness_hex"])))
    assert d["version"] == 1  # complex_test_payloads.rs:67
    assert d["state_root"].hex() == g["state_root"]  # complex_test_payloads.rs:68


def test_golden_4_storage_root(oracle, goldens):
    # the one-slot storage trie of golden 4
    g = goldens["compact_goldens"][3]
    d = parse_pre_image_dump(oracle.compact_decode(bytes.fromhex(g["witness_hex"])))
    assert [v.hex() for v in d["storage"].values()] == ["768c3c9e7d4393a36e3198da611dda885ea29b9b0f044fefa307f7853c4cd1dc"]


def _parse_instr_dump(b):
    ver, n = struct.unpack_from("<BI", b, 0)
    pos = 5
    out = []
    for _ in range(n):
        op = b[pos]
        pos += 1
        if op == 0:
            k = b[pos]
            nib = list(b[pos + 1 : pos + 1 + k])
            pos += 1 + k
            (ln,) = struct.unpack_from("<I", b, pos)
            out.append(("leaf", nib, b[pos + 4 : pos + 4 + ln]))
            pos += 4 + ln
        elif op == 1:
            k = b[pos]
            out.append(("extension", list(b[pos + 1 : pos + 1 + k])))
            pos += 1 + k
        elif op == 2:
            out.append(("branch", struct.unpack_from("<I", b, pos)[0]))
            pos += 4
        elif op == 3:
            out.append(("hash", b[pos : pos + 32]))
            pos += 32
        elif op == 4:
            (ln,) = struct.unpack_from("<I", b, pos)
            out.append(("code", b[pos + 4 : pos + 4 + ln]))
            pos += 4 + ln
        elif op == 5:
            k = b[pos]
            nib = list(b[pos + 1 : pos + 1 + k])
            pos += 1 + k
            out.append(("account_leaf", nib, b[pos : pos + 32], b[pos + 32 : pos + 64], b[pos + 64], b[pos + 65]))
            pos += 66
        elif op == 6:
            out.append(("empty_root",))
    assert pos == len(b)
    return ver, out


def test_simple_payload_instruction_kat(oracle, goldens):
    # compact_prestate_processing.rs:1471-1497
    sp = goldens["simple_payload"]
    ver, ins = _parse_instr_dump(oracle.compact_instructions(bytes.fromhex(sp["witness_hex"])))
    assert ver == 1
    assert len(ins) == len(sp["instructions"])
    for got, want in zip(ins, sp["instructions"]):
        assert got[0] == want["op"]
        if "key_bytes" in want:
            assert got[1] == oracle.key_bytes_to_nibbles(bytes.fromhex(want["key_bytes"]))
        if "value" in want:
            assert got[2].hex() == want["value"]
        if "mask" in want:
            assert got[1] == want["mask"]
    # key decoding spot checks
    assert oracle.key_bytes_to_nibbles(bytes.fromhex("10")) == [0]
    assert oracle.key_bytes_to_nibbles(bytes.fromhex("0350")) == [5]
    assert len(oracle.key_bytes_to_nibbles(bytes.fromhex("00" + "00" * 30 + "12"))) == 62


@pytest.mark.parametrize(
    "witness,code",
    [
        (b"", 1),  # MissingHeader
        (b"\x01\x07", 2),  # InvalidOperator
        (b"\x01\x05\x41\x10", 3),  # account leaf: flags byte missing -> UnexpectedEndOfStream
        (b"\x01\x00\x58", 4),  # truncated CBOR byte string -> InvalidByteVector
        (b"\x01\x03\x00", 5),  # short raw hash -> InvalidBytesForType
        (b"\x01\x06\x06", 7),  # two entries left -> NonSingleEntryAfterProcessing
        (b"\x01\x06\x02\x03", 8),  # branch wants 2 nodes, 1 precedes
        (b"\x01\x06\x02\x1a\x00\x01\x00\x00", 9),  # mask bit 16 set
    ],
)
def test_compact_error_variants(oracle, witness, code):
    with pytest.raises(OracleError) as e:
        oracle.compact_decode(witness)
    assert e.value.code == code


def test_header_only_witness_is_empty_state(oracle, goldens):
    # compact_prestate_processing.rs:342: nothing but the header -> default (empty) output
    d = parse_pre_image_dump(oracle.compact_decode(b"\x01"))
    assert d["state_root"].hex() == goldens["constants"]["EMPTY_TRIE_HASH"]
    assert d["storage"] == {} and d["code"] == {}
