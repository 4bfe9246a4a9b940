// arena.h — the node arena the GPU hashes: a forest (DAG) of MPT nodes in HBM, sorted by level.
//
// Replaces eth_trie_utils' heap-allocated `HashedPartialTrie` nodes with per-node
// `Arc<RwLock<Option<H256>>>` caches.  Here a node is one 16-byte
// record; every byte string it refers to lives in a pool; its result ("ref": the bytes a parent
// embeds, i.e. the raw RLP when shorter than 32 bytes, else the Keccak-256) is one 32-byte slot
// plus a length byte.  Records are ordered by level (all children of a node have a smaller
// level), so hashing is one launch per level over a contiguous range, bottom-up.
#pragma once
#include <cstdint>

namespace ppd {

enum NodeKind : uint32_t {
  NK_HASH = 0,          // never stored: a hashed-out subtree is the id HASH_ID_BASE + index into hash_pool (Node::Hash)
  NK_LEAF = 1,          // a0 = key byte offset, a1 = value offset, a2 = value length   (Node::Leaf)
  NK_LEAF_ACCOUNT = 2,  // a0 = key byte offset, a1 = account record index (Node::Leaf holding rlp(AccountRlp))
  NK_EXT = 3,           // a0 = key byte offset, a1 = child node           (Node::Extension)
  NK_BRANCH = 4,        // a0 = first child slot in child_pool, a1 = 16-bit child mask (Node::Branch, empty value)
  NK_ROOT = 5,          // a1 = child node or NODE_EMPTY: PartialTrie::hash() of a trie whose root is that node
};

static const uint32_t NODE_EMPTY = 0xffffffffu;  // Node::Empty as a child / as a trie root

// word0 = kind | nib_start << 8 | nib_len << 16 ; nibble i of a key is the high (i even) or low
// (i odd) half of key_pool[key_off + i / 2]
struct alignas(16) NodeRec {  // (one 128-bit load / store on the device)
  uint32_t w0, a0, a1, a2;
};
static inline uint32_t node_w0(uint32_t kind, uint32_t nib_start, uint32_t nib_len) { return kind | (nib_start << 8) | (nib_len << 16); }

// plonky2_evm AccountRlp with a late-bound storage root (decoding.rs:438-452: the root of the
// account's storage trie *after* this txn's storage writes)
struct AccountRec {
  uint8_t nonce[32];        // big-endian U256
  uint8_t balance[32];      // big-endian U256
  uint8_t storage_root[32]; // used when storage_src == NODE_EMPTY
  uint8_t code_hash[32];
  uint32_t storage_src;     // NK_ROOT node whose ref is the storage root, or NODE_EMPTY
  uint32_t pad[3];
};

// IrDump plan segments (ppd_dump.cu): seg_b is the root of a trie to cut and serialise (a node id, a hashed-out id or
// NODE_EMPTY), or one of these kinds
static const uint32_t IR_SEG_LITERAL = 0xfffffffeu;  // seg_a bytes the HOST writes after the copy back (host-shaped blocks)
static const uint32_t IR_SEG_REF = 0xfffffffdu;      // 32 bytes of ref[seg_a] (a trie root after the txn)
static const uint32_t IR_SEG_FLAT = 0xfffffffcu;     // seg_a bytes of the FlatBlock resident in HBM, from offset seg_c
static const uint32_t IR_SEG_KEY32 = 0xfffffffbu;    // 32 bytes of key_pool at seg_a (a hashed address)
static const uint32_t IR_SEG_LIT_DEV = 0xfffffffau;  // seg_a bytes of the uploaded literal pool, from offset seg_c
static const uint32_t IR_SEG_ROOT_ONLY = 0xfffffff9u;  // the trie rooted at seg_a with only its root kept (a dummy entry's tries)
static const uint32_t IR_SEG_KIND_MIN = 0xfffffff9u;

// Device-side view of one arena (all pointers are device pointers).
struct ArenaView {
  const NodeRec* nodes;
  const uint8_t* key_pool;
  const uint8_t* val_pool;       // values start on 4-byte boundaries
  const uint8_t* hash_pool;      // 32 bytes each
  const uint32_t* child_pool;
  const AccountRec* accounts;
  uint8_t* ref;                  // [n_nodes][32]
  uint8_t* ref_len;              // [n_nodes]
  unsigned long long* counters;  // [0] nodes hashed, [1] permutations
};

}  // namespace ppd
