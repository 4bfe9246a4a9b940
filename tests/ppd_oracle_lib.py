"""ctypes binding of the CPU oracle.  Test infrastructure: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import struct
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libppd_oracle.so")


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"oracle status {code}: {msg}")
        self.code = code
        self.msg = msg


def build(force=False):
    src = os.path.join(ORACLE_DIR, "ppd_oracle.cpp")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return SO


class Oracle:
    def __init__(self, path):
        L = ctypes.CDLL(path)
        self.L = L
        u8p = ctypes.POINTER(ctypes.c_uint8)
        for name in ("oracle_compact_instructions", "oracle_compact_decode", "oracle_block_decode", "oracle_compact_to_direct"):
            getattr(L, name).argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(u8p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
            getattr(L, name).restype = ctypes.c_int
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_keccak256.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        L.oracle_keccak256_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.oracle_trie_root_from_leaves.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p]
        L.oracle_last_stats.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_key_bytes_to_nibbles.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]

    def _call(self, fn, data):
        u8p = ctypes.POINTER(ctypes.c_uint8)
        out, n = u8p(), ctypes.c_size_t()
        err = ctypes.create_string_buffer(512)
        rc = fn(bytes(data), len(data), ctypes.byref(out), ctypes.byref(n), err, 512)
        if rc != 0:
            raise OracleError(rc, err.value.decode(errors="replace"))
        res = ctypes.string_at(out, n.value)
        self.L.oracle_free(out)
        return res

    def keccak256(self, data: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.L.oracle_keccak256(bytes(data), len(data), out)
        return out.raw

    def keccak256_batch(self, data, offsets):
        """data: np.uint8 array, offsets: np.uint64 array of n+1 entries -> np.uint8 [n,32]"""
        import numpy as np

        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.empty((n, 32), dtype=np.uint8)
        self.L.oracle_keccak256_batch(data.ctypes.data, offsets.ctypes.data, n, out.ctypes.data)
        return out

    def key_bytes_to_nibbles(self, key: bytes):
        out = ctypes.create_string_buffer(64)
        n = ctypes.c_size_t()
        rc = self.L.oracle_key_bytes_to_nibbles(bytes(key), len(key), out, ctypes.byref(n))
        if rc:
            raise OracleError(rc, "key")
        return list(out.raw[: n.value])

    def compact_instructions(self, witness: bytes) -> bytes:
        return self._call(self.L.oracle_compact_instructions, witness)

    def compact_decode(self, witness: bytes) -> bytes:
        return self._call(self.L.oracle_compact_decode, witness)

    def block_decode(self, flat: bytes) -> bytes:
        return self._call(self.L.oracle_block_decode, flat)

    def compact_to_direct(self, witness: bytes) -> bytes:
        """The tries a compact witness decodes to, as a DirectPreImage payload (include/ppd_flat.h, pre_image_kind 2)."""
        return self._call(self.L.oracle_compact_to_direct, witness)

    def trie_root_from_leaves(self, keys, val_off, vals) -> bytes:
        import numpy as np

        keys = np.ascontiguousarray(keys, dtype=np.uint8)
        val_off = np.ascontiguousarray(val_off, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.uint8)
        out = ctypes.create_string_buffer(32)
        rc = self.L.oracle_trie_root_from_leaves(keys.ctypes.data, val_off.ctypes.data, vals.ctypes.data, len(val_off) - 1, out)
        if rc:
            raise OracleError(rc, "trie_root_from_leaves")
        return out.raw

    def last_stats(self):
        a = (ctypes.c_uint64 * 4)()
        self.L.oracle_last_stats(a)
        return {"nodes_hashed": a[0], "node_perms": a[1], "other_hashes": a[2], "other_perms": a[3]}


def parse_pre_image_dump(b: bytes):
    magic, ver = struct.unpack_from("<IB", b, 0)
    assert magic == 0x50445050
    pos = 5
    root = b[pos : pos + 32]
    pos += 32
    (ns,) = struct.unpack_from("<I", b, pos)
    pos += 4
    storage = {}
    for _ in range(ns):
        storage[b[pos : pos + 32]] = b[pos + 32 : pos + 64]
        pos += 64
    (nc,) = struct.unpack_from("<I", b, pos)
    pos += 4
    code = {}
    for _ in range(nc):
        (ln,) = struct.unpack_from("<I", b, pos + 32)
        code[b[pos : pos + 32]] = ln
        pos += 36
    nodes, perms = struct.unpack_from("<QQ", b, pos)
    return {"version": ver, "state_root": root, "storage": storage, "code": code, "nodes_hashed": nodes, "perms": perms}


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = Oracle(build())
    return _cached
