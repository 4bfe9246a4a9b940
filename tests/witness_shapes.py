"""Hand-built compact witnesses whose tree is NOT the canonical trie of its items.

The reference re-inserts every leaf / hashed-out subtree with its full key
(protocol_decoder/src/compact/compact_to_partial_trie.rs:49-139), so a witness with a single-child
branch, an extension over a leaf, an EmptyRoot child ... still decodes to the canonical trie.  The
CUDA library converts canonical witnesses directly and falls back to the general path for these:
both must agree with the oracle, and each non-canonical witness must give the same root as its
canonical twin."""
from proof_protocol_decoder_b200.synth import cbor_bytes, cbor_uint, compact_key

OP_LEAF, OP_EXT, OP_BRANCH, OP_HASH, OP_CODE, OP_ACCOUNT, OP_EMPTY = range(7)


def nibs(hexstr):
    return [int(c, 16) for c in hexstr]


def account(key_nibbles, nonce=None, balance=None):
    flags = (4 if nonce is not None else 0) | (8 if balance is not None else 0)
    b = bytes([OP_ACCOUNT]) + cbor_bytes(compact_key(key_nibbles)) + bytes([flags])
    if nonce is not None:
        b += cbor_uint(nonce)
    if balance is not None:
        b += cbor_bytes(balance.to_bytes((balance.bit_length() + 7) // 8 or 1, "big"))
    return b


def account_with_storage(key_nibbles, storage_stream, balance=7):
    # stream order: storage subtree, then the leaf (flags: storage | balance)
    return storage_stream + bytes([OP_ACCOUNT]) + cbor_bytes(compact_key(key_nibbles)) + bytes([2 | 8]) + cbor_bytes(bytes([balance]))


def leaf(key_nibbles, value):
    return bytes([OP_LEAF]) + cbor_bytes(compact_key(key_nibbles)) + cbor_bytes(value)


def ext(key_nibbles):
    return bytes([OP_EXT]) + cbor_bytes(compact_key(key_nibbles))


def branch(mask):
    return bytes([OP_BRANCH]) + cbor_uint(mask)


def hashnode(h):
    return bytes([OP_HASH]) + bytes(h)


HDR = b"\x01"
K62 = "ab" * 31  # 62 nibbles
K60 = "cd" * 30

# (name, non-canonical witness, canonical twin)
PAIRS = [
    (
        "single_child_branch",
        HDR + account(nibs(K62 + "1"), balance=5) + branch(1 << 3),
        HDR + account(nibs("3" + K62 + "1"), balance=5),
    ),
    (
        "extension_over_leaf",
        HDR + account(nibs(K60), nonce=9) + ext(nibs("12")),
        HDR + account(nibs("12" + K60), nonce=9),
    ),
    (
        "extension_over_extension",
        HDR + account(nibs(K60 + "1"), balance=1) + account(nibs(K60 + "2"), balance=2) + branch((1 << 4) | (1 << 9)) + ext(nibs("7")) + ext(nibs("e")),
        HDR + account(nibs(K60 + "1"), balance=1) + account(nibs(K60 + "2"), balance=2) + branch((1 << 4) | (1 << 9)) + ext(nibs("e7")),
    ),
    (
        "empty_root_child",
        HDR + bytes([OP_EMPTY]) + account(nibs(K62), balance=3) + branch((1 << 0) | (1 << 5)),
        HDR + account(nibs("5" + K62), balance=3),
    ),
    (
        "nested_single_children",
        HDR + account(nibs(K60), balance=300) + branch(1 << 15) + branch(1 << 0) + account(nibs(K62), nonce=1) + branch((1 << 2) | (1 << 8)),
        HDR + account(nibs("0f" + K60), balance=300) + account(nibs(K62), nonce=1) + branch((1 << 2) | (1 << 8)),
    ),
    (
        "hash_under_single_child_branch",
        HDR + hashnode(bytes(range(32))) + branch(1 << 6) + account(nibs(K62), balance=1) + branch((1 << 1) | (1 << 2)),
        HDR + hashnode(bytes(range(32))) + ext(nibs("6")) + account(nibs(K62), balance=1) + branch((1 << 1) | (1 << 2)),
    ),
    (
        "noncanonical_storage_trie",
        HDR + account_with_storage(nibs("1" + K62), leaf(nibs(K62), b"\x2a") + branch(1 << 9) + ext(nibs("4")))
        + account(nibs(K62), balance=1) + branch((1 << 0) | (1 << 3)),
        HDR + account_with_storage(nibs("1" + K62), leaf(nibs("49" + K62), b"\x2a")) + account(nibs(K62), balance=1) + branch((1 << 0) | (1 << 3)),
    ),
]


def code(b):
    return bytes([OP_CODE]) + cbor_bytes(b)


def account_with_code(key_nibbles, code_stream, code_len, balance=7, storage_stream=None):
    # stream order: code, then storage, then the leaf; the code size follows the balance (read and discarded)
    flags = 1 | 8 | (2 if storage_stream is not None else 0)
    return (code_stream + (storage_stream or b"") + bytes([OP_ACCOUNT]) + cbor_bytes(compact_key(key_nibbles)) + bytes([flags])
            + cbor_bytes(bytes([balance])) + cbor_uint(code_len))


def long_code_witness(code_sizes, seed=7):
    """A canonical witness of len(code_sizes) <= 16 accounts under one branch, account i carrying code_sizes[i]
    bytes of inline code: with sizes of several tiles (4 KiB) the instruction chain jumps over whole tiles, and
    over the first tile of a tile group, in the GPU parser's boundary search."""
    import random

    rnd = random.Random(seed)
    out = HDR
    mask = 0
    for i, size in enumerate(code_sizes):
        body = bytes(rnd.getrandbits(8) for _ in range(size))
        out += account_with_code(nibs(("%x" % ((i * 7 + 3) % 16)) * 63), code(body), size, balance=1 + i)
        mask |= 1 << i
    return out + branch(mask)
