"""Size-independent properties of an IrDump, checked with a third, independent implementation
(pure-Python RLP + the generator's numpy Keccak): a subset trie hashes to the root of the trie it
was cut from, so consecutive IRs chain."""
from proof_protocol_decoder_b200 import synth


def _hp(nib, leaf):
    flag = (2 if leaf else 0) + (len(nib) & 1)
    if len(nib) & 1:
        out = [(flag << 4) | nib[0]]
        rest = nib[1:]
    else:
        out = [flag << 4]
        rest = nib
    for i in range(0, len(rest), 2):
        out.append((rest[i] << 4) | rest[i + 1])
    return bytes(out)


def _enc(node):
    """-> (raw_or_hash bytes, is_hash)"""
    k = node[0]
    if k == "empty":
        return b"\x80", False
    if k == "hash":
        return node[1], True
    if k == "leaf":
        raw = synth.rlp_list([synth.rlp_str(_hp(node[1], True)), synth.rlp_str(node[2])])
    elif k == "extension":
        c, is_h = _enc(node[2])
        raw = synth.rlp_list([synth.rlp_str(_hp(node[1], False)), synth.rlp_str(c) if is_h else c])
    else:
        items = []
        for ch in node[1]:
            c, is_h = _enc(ch)
            items.append(synth.rlp_str(c) if is_h else c)
        items.append(synth.rlp_str(node[2]) if node[2] else b"\x80")
        raw = synth.rlp_list(items)
    if len(raw) >= 32:
        return synth.keccak256(raw), True
    return raw, False


def subset_root(node) -> bytes:
    c, is_h = _enc(node)
    return c if is_h else synth.keccak256(c)


def test_subset_root_of_a_single_leaf():
    # the one-slot storage trie of golden 4
    key = bytes.fromhex("8015657e298d35290e69628be03d91f74d613caf3fdc9b3c8a8b0b2c2f502f50")
    nib = [x for b in key for x in (b >> 4, b & 15)]
    root = subset_root(("leaf", nib, bytes.fromhex("84deadbeef")))
    assert len(root) == 32
