#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the compact pre-image path.

Run in the build container (where /root/reference exists); the output JSON is
committed so the tests never read /root/reference at run time.

Sources (all data, no code):
  protocol_decoder/src/compact/complex_test_payloads.rs:14-30   six (witness hex, state root) pairs
  protocol_decoder/src/compact/large_test_payloads/test_payload_{5,6}.txt
  protocol_decoder/src/compact/compact_prestate_processing.rs:1439   SIMPLE_PAYLOAD_STR (instruction KAT, :1483-1492)
  protocol_decoder/src/types.rs:24-41   EMPTY_CODE_HASH / EMPTY_TRIE_HASH / EMPTY_ACCOUNT_BYTES_RLPED
"""
import json, os, re, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "protocol_decoder/src")


def byte_array(src, name):
    m = re.search(name + r"[^=]*=\s*(?:H256\()?\[([0-9,\s]+)\]", src)
    return bytes(int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()).hex()


def main():
    out = {"compact_goldens": []}
    payloads = open(os.path.join(SRC, "compact/complex_test_payloads.rs")).read()
    for m in re.finditer(r'TEST_PAYLOAD_(\d): TestProtocolInputAndRoot = TestProtocolInputAndRoot \{\s*byte_str: "([0-9a-f]+)",\s*root_str: "([0-9a-f]+)"', payloads):
        out["compact_goldens"].append({"name": f"payload_{m.group(1)}", "witness_hex": m.group(2), "state_root": m.group(3)})
    for m in re.finditer(r'TEST_PAYLOAD_(\d): TestProtocolInputAndRoot = TestProtocolInputAndRoot \{\s*byte_str: include_str!\("([^"]+)"\),\s*root_str: "([0-9a-f]+)"', payloads):
        hexs = open(os.path.join(SRC, "compact", m.group(2))).read().strip()
        out["compact_goldens"].append({"name": f"payload_{m.group(1)}", "witness_hex": hexs, "state_root": m.group(3)})
    out["compact_goldens"].sort(key=lambda g: g["name"])
    assert len(out["compact_goldens"]) == 6

    cpp = open(os.path.join(SRC, "compact/compact_prestate_processing.rs")).read()
    simple = re.search(r'SIMPLE_PAYLOAD_STR: &str = "([0-9a-f]+)"', cpp).group(1)
    # expected instruction list, transcribed from compact_prestate_processing.rs:1483-1492
    out["simple_payload"] = {
        "witness_hex": simple,
        "instructions": [
            {"op": "leaf", "key_bytes": "10", "value": "31323334"},
            {"op": "leaf", "key_bytes": "10", "value": "31323334"},
            {"op": "branch", "mask": 0b00110000},
            {"op": "leaf", "key_bytes": "0350", "value": "31323335"},
            {"op": "branch", "mask": 0b00011000},
            {"op": "extension", "key_bytes": "0000000000000000000000000000000000000000000000000000000000000012"},
        ],
    }
    types = open(os.path.join(SRC, "types.rs")).read()
    out["constants"] = {
        "EMPTY_CODE_HASH": byte_array(types, "EMPTY_CODE_HASH"),
        "EMPTY_TRIE_HASH": byte_array(types, "EMPTY_TRIE_HASH"),
        "EMPTY_ACCOUNT_BYTES_RLPED": byte_array(types, r"EMPTY_ACCOUNT_BYTES_RLPED: \[u8; 70\]"),
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, {k: (len(v) if isinstance(v, list) else "ok") for k, v in out.items()})


if __name__ == "__main__":
    main()
