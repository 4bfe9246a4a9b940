"""The JSON wire format of `BlockTrace` (the step immediately before the hot path).

Mirrors the serde schema of the reference, field for field:

  protocol_decoder/src/trace_protocol.rs:40-48     BlockTrace { trie_pre_images, txn_info }
  trace_protocol.rs:50-108                         BlockTraceTriePreImages: snake_case enum tags "separate" / "combined";
                                                   CombinedPreImages { compact: TrieCompact(ByteString) }
  trace_protocol.rs:112-145                        TxnInfo { traces: HashMap<Address, TxnTrace>, meta: TxnMeta }
  trace_protocol.rs:152-183                        TxnTrace: every field optional and skipped when None
  trace_protocol.rs:189-196                        ContractCodeUsage: {"read": H256} | {"write": ByteString}
  protocol_decoder/src/deserializers.rs:8-79       ByteString: hex string, "0x" / "0X" prefix optional on input,
                                                   always "0x" + lower-case hex on output

Scalars follow the crates the reference uses (RECALLED, not under /root/reference): `ethereum-types` H160 / H256 are
"0x" + exactly 40 / 64 hex digits (the prefix is required), U256 is "0x" + the minimal hex digits ("0x0" for zero;
leading zeros and odd digit counts are accepted on input), `u64` is a JSON number.

Only parsing and formatting live here; the result is the dataclasses of `trace_protocol.py`, whose `to_flat` produces
the FlatBlock the C ABI takes.  Error behaviour: everything serde would reject raises `WireFormatError`; the one place
where the reference PANICS instead (`remove_hex_prefix_if_present` slices `data[..2]` of a string shorter than two
bytes, deserializers.rs:31-38) raises `WireFormatPanic`, a subclass, so that a caller can tell the two apart.
"""
import json
from typing import Any, Dict

from .trace_protocol import BlockTrace, ContractCodeUsage, TxnInfo, TxnMeta, TxnTrace


class WireFormatError(ValueError):
    """What serde reports as a deserialisation error."""


class WireFormatPanic(WireFormatError):
    """Inputs on which the reference's deserialiser panics (deserializers.rs:31-38)."""


_HEX = set("0123456789abcdefABCDEF")


# ---- scalars -----------------------------------------------------------------------------------------------------
def byte_string_from_json(v: Any) -> bytes:
    """deserializers.rs:41-66: a string of hex digits with an optional 0x / 0X prefix."""
    if not isinstance(v, str):
        raise WireFormatError("a hex encoded string with a prefix")  # the visitor's `expecting` text
    if len(v.encode()) < 2:
        raise WireFormatPanic(f"byte index 2 is out of bounds of `{v}`")  # &data[..2]
    body = v[2:] if v[:2] in ("0x", "0X") else v
    if len(body) % 2:
        raise WireFormatError("Odd number of digits")  # hex::FromHexError::OddLength
    for i, ch in enumerate(body):
        if ch not in _HEX:
            raise WireFormatError(f"Invalid character {ch!r} at position {i}")
    return bytes.fromhex(body)


def byte_string_to_json(b: bytes) -> str:
    """deserializers.rs:70-79"""
    return "0x" + bytes(b).hex()


def _fixed_hash_from_json(v: Any, n: int, what: str) -> bytes:
    if not isinstance(v, str):
        raise WireFormatError(f"{what}: expected a 0x-prefixed hex string")
    if not v.startswith("0x"):
        raise WireFormatError(f"{what}: 0x prefix is missing")
    body = v[2:]
    if len(body) != 2 * n:
        raise WireFormatError(f"{what}: expected {2 * n} hex digits, got {len(body)}")
    for i, ch in enumerate(body):
        if ch not in _HEX:
            raise WireFormatError(f"{what}: invalid hex character {ch!r} at {i}")
    return bytes.fromhex(body)


def address_from_json(v: Any) -> bytes:
    return _fixed_hash_from_json(v, 20, "Address")


def h256_from_json(v: Any) -> bytes:
    return _fixed_hash_from_json(v, 32, "H256")


def u256_from_json(v: Any) -> int:
    if not isinstance(v, str):
        raise WireFormatError("U256: expected a 0x-prefixed hex string")
    if not v.startswith("0x"):
        raise WireFormatError("U256: 0x prefix is missing")
    body = v[2:]
    if len(body) > 64:
        raise WireFormatError(f"U256: expected at most 64 hex digits, got {len(body)}")
    for i, ch in enumerate(body):
        if ch not in _HEX:
            raise WireFormatError(f"U256: invalid hex character {ch!r} at {i}")
    return int(body, 16) if body else 0


def u256_to_json(v: int) -> str:
    if not 0 <= v < 1 << 256:
        raise WireFormatError("U256 out of range")
    return hex(v)


def _u64_from_json(v: Any, what: str) -> int:
    if isinstance(v, bool) or not isinstance(v, int) or not 0 <= v < 1 << 64:
        raise WireFormatError(f"{what}: expected a u64")
    return v


def _struct(v: Any, what: str, required: tuple, optional: tuple = ()) -> Dict[str, Any]:
    if not isinstance(v, dict):
        raise WireFormatError(f"{what}: expected a map")
    for k in required:
        if k not in v:
            raise WireFormatError(f"{what}: missing field `{k}`")
    return v  # serde ignores unknown fields unless deny_unknown_fields is set (it is not)


def _enum(v: Any, what: str, variants: tuple):
    """externally tagged enum: a map with exactly one key, the snake_case variant name"""
    if not isinstance(v, dict) or len(v) != 1:
        raise WireFormatError(f"{what}: expected a map with a single key, one of {variants}")
    (tag, body), = v.items()
    if tag not in variants:
        raise WireFormatError(f"{what}: unknown variant `{tag}`, expected one of {variants}")
    return tag, body


# ---- structures --------------------------------------------------------------------------------------------------
def code_usage_from_json(v: Any) -> ContractCodeUsage:
    tag, body = _enum(v, "ContractCodeUsage", ("read", "write"))
    if tag == "read":
        return ContractCodeUsage(read=h256_from_json(body))
    return ContractCodeUsage(write=byte_string_from_json(body))


def code_usage_to_json(c: ContractCodeUsage) -> dict:
    if c.read is not None:
        return {"read": "0x" + bytes(c.read).hex()}
    return {"write": byte_string_to_json(c.write or b"")}


def txn_trace_from_json(v: Any) -> TxnTrace:
    d = _struct(v, "TxnTrace", ())
    t = TxnTrace()
    if d.get("balance") is not None:
        t.balance = u256_from_json(d["balance"])
    if d.get("nonce") is not None:
        t.nonce = u256_from_json(d["nonce"])
    if d.get("storage_read") is not None:
        if not isinstance(d["storage_read"], list):
            raise WireFormatError("TxnTrace.storage_read: expected a sequence")
        t.storage_read = [h256_from_json(x) for x in d["storage_read"]]
    if d.get("storage_written") is not None:
        if not isinstance(d["storage_written"], dict):
            raise WireFormatError("TxnTrace.storage_written: expected a map")
        t.storage_written = {h256_from_json(k): u256_from_json(x) for k, x in d["storage_written"].items()}
    if d.get("code_usage") is not None:
        t.code_usage = code_usage_from_json(d["code_usage"])
    if d.get("self_destructed") is not None:
        if not isinstance(d["self_destructed"], bool):
            raise WireFormatError("TxnTrace.self_destructed: expected a bool")
        t.self_destructed = d["self_destructed"]
    return t


def txn_trace_to_json(t: TxnTrace) -> dict:
    d: Dict[str, Any] = {}  # skip_serializing_if = "Option::is_none" on every field
    if t.balance is not None:
        d["balance"] = u256_to_json(t.balance)
    if t.nonce is not None:
        d["nonce"] = u256_to_json(t.nonce)
    if t.storage_read is not None:
        d["storage_read"] = ["0x" + bytes(k).hex() for k in t.storage_read]
    if t.storage_written is not None:
        d["storage_written"] = {"0x" + bytes(k).hex(): u256_to_json(x) for k, x in t.storage_written.items()}
    if t.code_usage is not None:
        d["code_usage"] = code_usage_to_json(t.code_usage)
    if t.self_destructed is not None:
        d["self_destructed"] = bool(t.self_destructed)
    return d


def txn_meta_from_json(v: Any) -> TxnMeta:
    d = _struct(v, "TxnMeta", ("byte_code", "new_txn_trie_node_byte", "new_receipt_trie_node_byte", "gas_used"))
    return TxnMeta(
        byte_code=byte_string_from_json(d["byte_code"]),
        new_txn_trie_node_byte=byte_string_from_json(d["new_txn_trie_node_byte"]),
        new_receipt_trie_node_byte=byte_string_from_json(d["new_receipt_trie_node_byte"]),
        gas_used=_u64_from_json(d["gas_used"], "TxnMeta.gas_used"),
    )


def txn_meta_to_json(m: TxnMeta) -> dict:
    return {
        "byte_code": byte_string_to_json(m.byte_code),
        "new_txn_trie_node_byte": byte_string_to_json(m.new_txn_trie_node_byte),
        "new_receipt_trie_node_byte": byte_string_to_json(m.new_receipt_trie_node_byte),
        "gas_used": int(m.gas_used),
    }


def txn_info_from_json(v: Any) -> TxnInfo:
    d = _struct(v, "TxnInfo", ("traces", "meta"))
    if not isinstance(d["traces"], dict):
        raise WireFormatError("TxnInfo.traces: expected a map")
    return TxnInfo(traces={address_from_json(a): txn_trace_from_json(t) for a, t in d["traces"].items()}, meta=txn_meta_from_json(d["meta"]))


def txn_info_to_json(t: TxnInfo) -> dict:
    return {"traces": {"0x" + bytes(a).hex(): txn_trace_to_json(tr) for a, tr in t.traces.items()}, "meta": txn_meta_to_json(t.meta)}


# ---- TrieDirect(HashedPartialTrie), trace_protocol.rs:97-99 ------------------------------------------------------
# The serde form of `HashedPartialTrie` belongs to eth_trie_utils (rev 7fc3c3f5), which is NOT under /root/reference:
# what follows is RECALLED and unpinned.  A trie is {"node": Node, "hash": null | H256} (the cached hash is ignored on
# input and written as null); Node is an externally tagged enum with the variant names as written in Rust: "Empty",
# {"Hash": H256}, {"Branch": {"children": [16 tries], "value": [u8]}}, {"Extension": {"nibbles": Nibbles, "child": trie}},
# {"Leaf": {"nibbles": Nibbles, "value": [u8]}}; Nibbles is {"count": usize, "packed": U512 as 0x-hex}, the first nibble
# in the most significant position.  A bare Node (without the {"node": ...} wrapper) and hex strings for byte vectors
# are accepted on input too.  Parsed tries are the node tuples of flat.encode_node.
def _vec_u8_from_json(v: Any, what: str) -> bytes:
    if isinstance(v, str):
        return byte_string_from_json(v)
    if not isinstance(v, list) or any(isinstance(x, bool) or not isinstance(x, int) or not 0 <= x < 256 for x in v):
        raise WireFormatError(f"{what}: expected a sequence of u8")
    return bytes(v)


def nibbles_from_json(v: Any) -> list:
    d = _struct(v, "Nibbles", ("count", "packed"))
    count = d["count"]
    if isinstance(count, bool) or not isinstance(count, int) or not 0 <= count <= 64:
        raise WireFormatError("Nibbles.count: expected 0..64")
    if not isinstance(d["packed"], str) or not d["packed"].startswith("0x"):
        raise WireFormatError("Nibbles.packed: expected a 0x-prefixed hex string")
    body = d["packed"][2:]
    if len(body) > 128 or any(ch not in _HEX for ch in body):
        raise WireFormatError("Nibbles.packed: expected at most 128 hex digits")
    packed = int(body, 16) if body else 0
    if packed >> (4 * count):
        raise WireFormatError("Nibbles.packed: more nibbles than `count`")
    return [(packed >> (4 * (count - 1 - i))) & 15 for i in range(count)]


def nibbles_to_json(nibs) -> dict:
    packed = 0
    for n in nibs:
        packed = (packed << 4) | int(n)
    return {"count": len(nibs), "packed": hex(packed)}


def direct_trie_from_json(v: Any, depth: int = 0):
    if depth > 140:
        raise WireFormatError("HashedPartialTrie: nested deeper than any 64-nibble key allows")
    if isinstance(v, dict) and "node" in v:
        v = v["node"]
    if v == "Empty":
        return ("empty",)
    tag, body = _enum(v, "Node", ("Empty", "Hash", "Branch", "Extension", "Leaf"))
    if tag == "Empty":
        return ("empty",)
    if tag == "Hash":
        return ("hash", h256_from_json(body))
    if tag == "Branch":
        d = _struct(body, "Node::Branch", ("children", "value"))
        if not isinstance(d["children"], list) or len(d["children"]) != 16:
            raise WireFormatError("Node::Branch.children: expected 16 entries")
        return ("branch", [direct_trie_from_json(c, depth + 1) for c in d["children"]], _vec_u8_from_json(d["value"], "Node::Branch.value"))
    if tag == "Extension":
        d = _struct(body, "Node::Extension", ("nibbles", "child"))
        return ("extension", nibbles_from_json(d["nibbles"]), direct_trie_from_json(d["child"], depth + 1))
    d = _struct(body, "Node::Leaf", ("nibbles", "value"))
    return ("leaf", nibbles_from_json(d["nibbles"]), _vec_u8_from_json(d["value"], "Node::Leaf.value"))


def direct_trie_to_json(node) -> dict:
    k = node[0]
    if k == "empty":
        n: Any = "Empty"
    elif k == "hash":
        n = {"Hash": "0x" + bytes(node[1]).hex()}
    elif k == "branch":
        n = {"Branch": {"children": [direct_trie_to_json(c) for c in node[1]], "value": list(bytes(node[2]))}}
    elif k == "extension":
        n = {"Extension": {"nibbles": nibbles_to_json(node[1]), "child": direct_trie_to_json(node[2])}}
    else:
        n = {"Leaf": {"nibbles": nibbles_to_json(node[1]), "value": list(bytes(node[2]))}}
    return {"node": n, "hash": None}


def pre_images_from_json(v: Any) -> dict:
    """BlockTraceTriePreImages (trace_protocol.rs:50-108).  `combined` is decoded to {"combined": {"compact": bytes}},
    the form `BlockTrace.to_flat` takes.  `separate` with a Direct state trie and MultipleTries of Direct tries is decoded
    to node tuples ({"separate": {"state": {"direct": trie}, "storage": {"multiple_tries": {hashed address: {"direct":
    trie}}}}}: FlatBlock kind 2).  The other `separate` forms are accepted structurally and kept as parsed JSON: the
    reference has no decode path behind them (todo!() at processed_block_trace.rs:144,161), so `into_txn_proof_gen_ir`
    reports them as unimplemented."""
    tag, body = _enum(v, "BlockTraceTriePreImages", ("separate", "combined"))
    if tag == "combined":
        d = _struct(body, "CombinedPreImages", ("compact",))
        return {"combined": {"compact": byte_string_from_json(d["compact"])}}
    d = _struct(body, "SeparateTriePreImages", ("state", "storage"))
    stag, sbody = _enum(d["state"], "SeparateTriePreImage", ("uncompressed", "direct"))
    ttag, tbody = _enum(d["storage"], "SeparateStorageTriesPreImage", ("single_trie", "multiple_tries"))
    if stag != "direct" or ttag != "multiple_tries":
        return {"separate": d}
    if not isinstance(tbody, dict):
        raise WireFormatError("SeparateStorageTriesPreImage::MultipleTries: expected a map")
    tries = {}
    for h, t in tbody.items():
        ktag, kbody = _enum(t, "SeparateTriePreImage", ("uncompressed", "direct"))
        if ktag != "direct":
            return {"separate": d}
        tries[h256_from_json(h)] = {"direct": direct_trie_from_json(kbody)}
    return {"separate": {"state": {"direct": direct_trie_from_json(sbody)}, "storage": {"multiple_tries": tries}}}


def pre_images_to_json(p: dict) -> dict:
    if "combined" in p:
        return {"combined": {"compact": byte_string_to_json(p["combined"]["compact"])}}
    sep = p["separate"]
    state, storage = sep.get("state"), sep.get("storage")
    if isinstance(state, dict) and isinstance(state.get("direct"), tuple) and isinstance(storage, dict) and isinstance(storage.get("multiple_tries"), dict):
        return {"separate": {"state": {"direct": direct_trie_to_json(state["direct"])},
                             "storage": {"multiple_tries": {"0x" + bytes(h).hex(): {"direct": direct_trie_to_json(t["direct"])}
                                                            for h, t in storage["multiple_tries"].items()}}}}
    return {"separate": sep}


def block_trace_from_json(src: Any) -> BlockTrace:
    """`src`: JSON text (str / bytes) or the already parsed value."""
    if isinstance(src, (str, bytes, bytearray)):
        try:
            src = json.loads(src)
        except json.JSONDecodeError as e:
            raise WireFormatError(f"invalid JSON: {e}") from None
    d = _struct(src, "BlockTrace", ("trie_pre_images", "txn_info"))
    if not isinstance(d["txn_info"], list):
        raise WireFormatError("BlockTrace.txn_info: expected a sequence")
    return BlockTrace(trie_pre_images=pre_images_from_json(d["trie_pre_images"]), txn_info=[txn_info_from_json(t) for t in d["txn_info"]])


def block_trace_to_json(bt: BlockTrace) -> dict:
    return {"trie_pre_images": pre_images_to_json(bt.trie_pre_images), "txn_info": [txn_info_to_json(t) for t in bt.txn_info]}


def block_trace_dumps(bt: BlockTrace) -> str:
    return json.dumps(block_trace_to_json(bt), separators=(",", ":"))
