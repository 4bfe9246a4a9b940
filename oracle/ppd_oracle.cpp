// ppd_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain, single-threaded C++ restatement of the reference's algorithm for the
// hot path (compact pre-image -> tries -> per-txn proof-gen IR, with recursive
// RLP + Keccak-256 root computation).  It follows the reference's structure:
// insert-built pointer tries, recursive hashing with a per-node cache, scalar
// portable Keccak.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it; the product (libppd_b200.so)
// never does.
//
// PARITY PINNING: pinned by the reference's own fixtures for the pre-image half
// (six witness -> state-root goldens, the instruction-list KAT, three constants;
// tests/golden/reference_goldens.json, checked in tests/test_oracle_goldens.py).
// The per-txn half (decoding.rs, processed_block_trace.rs:209-343) has no tests
// or fixtures in the reference and its trie engine (eth_trie_utils, git rev
// 7fc3c3f5, NOT under /root/reference) is restated from its published behaviour:
// that half is "parity unpinned" (see DESIGN.md).
//
// Reference files followed (relative to /root/reference/protocol_decoder/src):
//   compact/compact_prestate_processing.rs:683-875   byte grammar
//   compact/compact_prestate_processing.rs:896-1003  field readers (ciborium subset)
//   compact/compact_prestate_processing.rs:1338-1390 key_bytes_to_nibbles
//   compact/compact_prestate_processing.rs:387-668   collapse rules (a stack machine)
//   compact/compact_to_partial_trie.rs:49-190        tree -> trie, account RLP, storage remap
//   processed_block_trace.rs:52-108, 209-343         trace -> per-txn node sets
//   decoding.rs:80-607                               txn loop, deltas, subsets, dummies, withdrawals
//   types.rs:24-44                                   constants
// Third-party semantics restated (not in /root/reference): eth_trie_utils 0.6.0
// (insert/delete/get/items/create_trie_subset/hash), tiny-keccak 2.0.2, rlp 0.5.2,
// ciborium 0.2.1.

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../include/ppd_flat.h"
#include "../include/ppd_status.h"

namespace orc {

using Bytes = std::vector<uint8_t>;

struct Err {
  int code;
  std::string msg;
};
[[noreturn]] static void fail(int code, const std::string& msg) { throw Err{code, msg}; }
// The payloads of the TraceParsingError variants (decoding.rs:31-49) follow the sentence as "; key=value ..." words: the
// form the product's ppd_last_error uses, so that a caller can rebuild the reference's error value.
static std::string hex_str(const uint8_t* p, size_t n) {
  std::string o;
  char t[3];
  for (size_t i = 0; i < n; i++) snprintf(t, sizeof t, "%02x", p[i]), o += t;
  return o;
}

// ---------------------------------------------------------------------------------------------
// Keccak-256 (tiny-keccak 2.0.2 behaviour: rate 136, padding 0x01 .. 0x80, 24 rounds)
// ---------------------------------------------------------------------------------------------
struct Stats {
  uint64_t nodes_hashed = 0;   // keccak invocations over trie-node encodings
  uint64_t node_perms = 0;     // keccak-f permutations spent on those
  uint64_t other_hashes = 0;   // address / slot / code hashes
  uint64_t other_perms = 0;
};
static thread_local Stats g_stats;

static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static inline uint64_t rotl64(uint64_t x, int s) { return (x << s) | (x >> (64 - s)); }

static void keccak_f1600(uint64_t a[25]) {
  for (int round = 0; round < 24; round++) {
    uint64_t c[5];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) {
      uint64_t d = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
      for (int y = 0; y < 25; y += 5) a[y + x] ^= d;
    }
    uint64_t last = a[1];
    for (int i = 0; i < 24; i++) {
      int j = KECCAK_PIL[i];
      uint64_t t = a[j];
      a[j] = rotl64(last, KECCAK_ROT[i]);
      last = t;
    }
    for (int y = 0; y < 25; y += 5) {
      uint64_t r[5];
      for (int x = 0; x < 5; x++) r[x] = a[y + x];
      for (int x = 0; x < 5; x++) a[y + x] = r[x] ^ (~r[(x + 1) % 5] & r[(x + 2) % 5]);
    }
    a[0] ^= KECCAK_RC[round];
  }
}

static uint64_t keccak256_raw(const uint8_t* data, size_t len, uint8_t out[32]) {
  uint64_t st[25];
  memset(st, 0, sizeof st);
  uint64_t perms = 0;
  while (len >= 136) {
    for (int i = 0; i < 17; i++) {
      uint64_t w;
      memcpy(&w, data + 8 * i, 8);
      st[i] ^= w;
    }
    keccak_f1600(st);
    perms++;
    data += 136;
    len -= 136;
  }
  uint8_t last[136];
  memset(last, 0, sizeof last);
  memcpy(last, data, len);
  last[len] ^= 0x01;
  last[135] ^= 0x80;
  for (int i = 0; i < 17; i++) {
    uint64_t w;
    memcpy(&w, last + 8 * i, 8);
    st[i] ^= w;
  }
  keccak_f1600(st);
  perms++;
  memcpy(out, st, 32);
  return perms;
}

struct H256 {
  uint8_t b[32];
  bool operator==(const H256& o) const { return memcmp(b, o.b, 32) == 0; }
  bool operator!=(const H256& o) const { return !(*this == o); }
  bool operator<(const H256& o) const { return memcmp(b, o.b, 32) < 0; }
};
struct H256Hash {
  size_t operator()(const H256& h) const {
    size_t v;
    memcpy(&v, h.b, sizeof v);
    return v;
  }
};

// utils.rs:11-13 `hash`
static H256 hash_bytes(const uint8_t* p, size_t n) {
  H256 h;
  g_stats.other_perms += keccak256_raw(p, n, h.b);
  g_stats.other_hashes++;
  return h;
}
static H256 hash_bytes(const Bytes& b) { return hash_bytes(b.data(), b.size()); }

static H256 h256_from_hex(const char* s) {
  H256 h;
  for (int i = 0; i < 32; i++) {
    unsigned v;
    sscanf(s + 2 * i, "%2x", &v);
    h.b[i] = (uint8_t)v;
  }
  return h;
}
// types.rs:24-34
static const H256 EMPTY_CODE_HASH = h256_from_hex("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470");
static const H256 EMPTY_TRIE_HASH = h256_from_hex("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421");

// ---------------------------------------------------------------------------------------------
// RLP (rlp 0.5.2 behaviour)
// ---------------------------------------------------------------------------------------------
static void rlp_len_prefix(Bytes& out, size_t len, uint8_t short_base, uint8_t long_base) {
  if (len < 56) {
    out.push_back((uint8_t)(short_base + len));
  } else {
    uint8_t tmp[8];
    int n = 0;
    for (size_t v = len; v; v >>= 8) tmp[n++] = (uint8_t)v;
    out.push_back((uint8_t)(long_base + n));
    for (int i = n - 1; i >= 0; i--) out.push_back(tmp[i]);
  }
}
static void rlp_append_str(Bytes& out, const uint8_t* p, size_t n) {
  if (n == 1 && p[0] < 0x80) {
    out.push_back(p[0]);
    return;
  }
  rlp_len_prefix(out, n, 0x80, 0xb7);
  out.insert(out.end(), p, p + n);
}
static void rlp_append_str(Bytes& out, const Bytes& b) { rlp_append_str(out, b.data(), b.size()); }
static Bytes rlp_wrap_list(const Bytes& payload) {
  Bytes out;
  rlp_len_prefix(out, payload.size(), 0xc0, 0xf7);
  out.insert(out.end(), payload.begin(), payload.end());
  return out;
}
// U256 / usize as a minimal big-endian string (0 -> 0x80)
static void rlp_append_u256(Bytes& out, const uint8_t be[32]) {
  int i = 0;
  while (i < 32 && be[i] == 0) i++;
  rlp_append_str(out, be + i, 32 - i);
}
static Bytes rlp_encode_u256(const uint8_t be[32]) {
  Bytes out;
  rlp_append_u256(out, be);
  return out;
}
static Bytes rlp_encode_u64(uint64_t v) {
  uint8_t be[32];
  memset(be, 0, 32);
  for (int i = 0; i < 8; i++) be[31 - i] = (uint8_t)(v >> (8 * i));
  return rlp_encode_u256(be);
}

struct RlpItem {
  bool is_list;
  const uint8_t* payload;
  size_t payload_len;
  size_t total_len;
};
// Decode one item header with the canonical-form checks the rlp crate applies.
static bool rlp_decode_item(const uint8_t* p, size_t n, RlpItem& it) {
  if (n == 0) return false;
  uint8_t b = p[0];
  if (b < 0x80) {
    it = {false, p, 1, 1};
    return true;
  }
  bool is_list = b >= 0xc0;
  uint8_t short_base = is_list ? 0xc0 : 0x80, long_base = is_list ? 0xf7 : 0xb7;
  size_t hdr, len;
  if (b <= long_base) {
    hdr = 1;
    len = b - short_base;
    if (!is_list && len == 1) {
      if (n < 2) return false;
      if (p[1] < 0x80) return false;  // RlpInvalidIndirection
    }
  } else {
    size_t ll = b - long_base;
    if (ll > 8 || n < 1 + ll) return false;
    if (p[1] == 0) return false;  // RlpDataLenWithZeroPrefix
    len = 0;
    for (size_t i = 0; i < ll; i++) len = (len << 8) | p[1 + i];
    if (len < 56) return false;  // RlpInvalidIndirection
    hdr = 1 + ll;
  }
  if (len > n - hdr) return false;
  it = {is_list, p + hdr, len, hdr + len};
  return true;
}

// plonky2_evm::generation::mpt::AccountRlp
struct Account {
  uint8_t nonce[32];
  uint8_t balance[32];
  H256 storage_root;
  H256 code_hash;
};
static Bytes account_encode(const Account& a) {
  Bytes pl;
  rlp_append_u256(pl, a.nonce);
  rlp_append_u256(pl, a.balance);
  rlp_append_str(pl, a.storage_root.b, 32);
  rlp_append_str(pl, a.code_hash.b, 32);
  return rlp_wrap_list(pl);
}
static bool rlp_take_u256(const uint8_t*& p, size_t& n, uint8_t out[32]) {
  RlpItem it;
  if (!rlp_decode_item(p, n, it) || it.is_list) return false;
  if (it.payload_len > 32) return false;
  if (it.payload_len > 0 && it.payload[0] == 0) return false;
  memset(out, 0, 32);
  memcpy(out + 32 - it.payload_len, it.payload, it.payload_len);
  p += it.total_len;
  n -= it.total_len;
  return true;
}
static bool rlp_take_h256(const uint8_t*& p, size_t& n, H256& out) {
  RlpItem it;
  if (!rlp_decode_item(p, n, it) || it.is_list || it.payload_len != 32) return false;
  memcpy(out.b, it.payload, 32);
  p += it.total_len;
  n -= it.total_len;
  return true;
}
static bool account_decode(const uint8_t* p, size_t n, Account& a) {
  RlpItem top;
  if (!rlp_decode_item(p, n, top) || !top.is_list) return false;
  const uint8_t* q = top.payload;
  size_t m = top.payload_len;
  return rlp_take_u256(q, m, a.nonce) && rlp_take_u256(q, m, a.balance) && rlp_take_h256(q, m, a.storage_root) &&
         rlp_take_h256(q, m, a.code_hash);
}
// EMPTY_ACCOUNT_BYTES_RLPED, types.rs:36-41
static Bytes empty_account_bytes() {
  Account a;
  memset(a.nonce, 0, 32);
  memset(a.balance, 0, 32);
  a.storage_root = EMPTY_TRIE_HASH;
  a.code_hash = EMPTY_CODE_HASH;
  return account_encode(a);
}

// Does `raw` decode as plonky2_evm's LegacyReceiptRlp {status: bool, cum_gas_used: U256,
// bloom: Bytes, logs: Vec<LogRlp{address: Address, topics: Vec<H256>, data: Bytes}>} ?
static bool is_legacy_receipt(const uint8_t* p, size_t n) {
  RlpItem top;
  if (!rlp_decode_item(p, n, top) || !top.is_list) return false;
  const uint8_t* q = top.payload;
  size_t m = top.payload_len;
  RlpItem it;
  // status: bool decodes as a u8 that must be 0 or 1
  if (!rlp_decode_item(q, m, it) || it.is_list || it.payload_len > 1) return false;
  if (it.payload_len == 1 && (it.payload[0] == 0 || it.payload[0] > 1)) return false;
  q += it.total_len, m -= it.total_len;
  uint8_t tmp[32];
  if (!rlp_take_u256(q, m, tmp)) return false;
  if (!rlp_decode_item(q, m, it) || it.is_list) return false;  // bloom
  q += it.total_len, m -= it.total_len;
  if (!rlp_decode_item(q, m, it) || !it.is_list) return false;  // logs
  const uint8_t* lq = it.payload;
  size_t lm = it.payload_len;
  while (lm > 0) {
    RlpItem log;
    if (!rlp_decode_item(lq, lm, log) || !log.is_list) return false;
    const uint8_t* f = log.payload;
    size_t fm = log.payload_len;
    RlpItem x;
    if (!rlp_decode_item(f, fm, x) || x.is_list || x.payload_len != 20) return false;
    f += x.total_len, fm -= x.total_len;
    if (!rlp_decode_item(f, fm, x) || !x.is_list) return false;
    const uint8_t* tq = x.payload;
    size_t tm = x.payload_len;
    while (tm > 0) {
      RlpItem t;
      if (!rlp_decode_item(tq, tm, t) || t.is_list || t.payload_len != 32) return false;
      tq += t.total_len, tm -= t.total_len;
    }
    f += x.total_len, fm -= x.total_len;
    if (!rlp_decode_item(f, fm, x) || x.is_list) return false;
    lq += log.total_len, lm -= log.total_len;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Nibbles (eth_trie_utils::nibbles::Nibbles; at most 64 nibbles at the pinned rev)
// ---------------------------------------------------------------------------------------------
struct Nibs {
  uint8_t n = 0;
  uint8_t d[64];
  void push(uint8_t v) {
    if (n >= 64) fail(PPD_ERR_KEY_ERROR, "nibble key longer than 64");
    d[n++] = v;
  }
  Nibs slice(int from) const {
    Nibs r;
    r.n = (uint8_t)(n - from);
    memcpy(r.d, d + from, r.n);
    return r;
  }
  Nibs prefix(int cnt) const {
    Nibs r;
    r.n = (uint8_t)cnt;
    memcpy(r.d, d, cnt);
    return r;
  }
  Nibs merge(const Nibs& o) const {
    Nibs r = *this;
    for (int i = 0; i < o.n; i++) r.push(o.d[i]);
    return r;
  }
  bool operator==(const Nibs& o) const { return n == o.n && memcmp(d, o.d, n) == 0; }
};
static Nibs nibs_from_bytes(const uint8_t* p, size_t n) {
  Nibs r;
  for (size_t i = 0; i < n; i++) {
    r.push(p[i] >> 4);
    r.push(p[i] & 15);
  }
  return r;
}
static Nibs nibs_from_h256(const H256& h) { return nibs_from_bytes(h.b, 32); }
static int common_prefix_len(const Nibs& a, const Nibs& b) {
  int m = std::min(a.n, b.n), i = 0;
  while (i < m && a.d[i] == b.d[i]) i++;
  return i;
}
// Nibbles::bytes_be at the pinned rev is value-minimal (leading zero bytes dropped): that is
// why utils.rs:49-59 pads.  RECALLED, unpinned (SURVEY.md 8c hazard 4).
static Bytes nibs_bytes_be_minimal(const Nibs& k) {
  Bytes full;
  int i = 0;
  if (k.n & 1) {
    full.push_back(k.d[0]);
    i = 1;
  }
  for (; i + 1 < k.n; i += 2) full.push_back((uint8_t)((k.d[i] << 4) | k.d[i + 1]));
  size_t z = 0;
  while (z < full.size() && full[z] == 0) z++;
  return Bytes(full.begin() + z, full.end());
}
// utils.rs:49-59
static H256 h_addr_nibs_to_h256(const Nibs& k) {
  Bytes b = nibs_bytes_be_minimal(k);
  H256 h;
  memset(h.b, 0, 32);
  if (b.size() > 32) fail(PPD_PANIC_H256_FROM_SLICE, "key longer than 32 bytes");
  memcpy(h.b + 32 - b.size(), b.data(), b.size());
  return h;
}
// H256::from_slice(&nibbles.bytes_be()) — panics unless exactly 32 bytes (decoding.rs:202,228-230)
static H256 h256_from_slice_of_nibs(const Nibs& k) {
  Bytes b = nibs_bytes_be_minimal(k);
  if (b.size() != 32) fail(PPD_PANIC_H256_FROM_SLICE, "H256::from_slice on a short bytes_be()");
  H256 h;
  memcpy(h.b, b.data(), 32);
  return h;
}
// hex-prefix encoding (Nibbles::to_hex_prefix_encoding)
static Bytes hex_prefix(const Nibs& k, bool is_leaf) {
  Bytes out;
  uint8_t flag = (uint8_t)((is_leaf ? 2 : 0) + (k.n & 1));
  int i = 0;
  if (k.n & 1) {
    out.push_back((uint8_t)((flag << 4) | k.d[0]));
    i = 1;
  } else {
    out.push_back((uint8_t)(flag << 4));
  }
  for (; i < k.n; i += 2) out.push_back((uint8_t)((k.d[i] << 4) | k.d[i + 1]));
  return out;
}

// ---------------------------------------------------------------------------------------------
// HashedPartialTrie (eth_trie_utils::partial_trie): immutable shared nodes, per-node hash cache
// ---------------------------------------------------------------------------------------------
enum Kind : uint8_t { EMPTY = 0, HASH = 1, BRANCH = 2, EXT = 3, LEAF = 4 };
struct Node;
using NodeP = std::shared_ptr<const Node>;
struct Node {
  Kind kind = EMPTY;
  H256 hash;                    // HASH
  std::array<NodeP, 16> ch;     // BRANCH
  Bytes value;                  // BRANCH value / LEAF value
  Nibs nib;                     // EXT / LEAF
  NodeP child;                  // EXT
  mutable int8_t ref_len = -1;  // cache: -1 unknown, 0..31 raw encoding inlined, 32 hashed
  mutable uint8_t ref[32];
};
// per thread: a process-wide singleton would make every branch copy contend on one reference count
static thread_local NodeP EMPTY_NODE = std::make_shared<Node>();
static NodeP mk_hash(const H256& h) {
  auto n = std::make_shared<Node>();
  n->kind = HASH;
  n->hash = h;
  return n;
}
static NodeP mk_leaf(const Nibs& k, const Bytes& v) {
  auto n = std::make_shared<Node>();
  n->kind = LEAF;
  n->nib = k;
  n->value = v;
  return n;
}
static NodeP mk_ext(const Nibs& k, const NodeP& c) {
  auto n = std::make_shared<Node>();
  n->kind = EXT;
  n->nib = k;
  n->child = c;
  return n;
}
static NodeP mk_branch(const std::array<NodeP, 16>& ch, const Bytes& v) {
  auto n = std::make_shared<Node>();
  n->kind = BRANCH;
  n->ch = ch;
  n->value = v;
  return n;
}
static std::array<NodeP, 16> empty_children() {
  std::array<NodeP, 16> c;
  for (auto& x : c) x = EMPTY_NODE;
  return c;
}

// --- hashing: Node::hash_intern / EncodedNode ---
static void node_ref(const Node& n);  // fills n.ref / n.ref_len
static void append_child_ref(Bytes& out, const Node& c) {
  node_ref(c);
  if (c.ref_len == 32) {
    out.push_back(0xa0);
    out.insert(out.end(), c.ref, c.ref + 32);
  } else {
    out.insert(out.end(), c.ref, c.ref + c.ref_len);  // append_raw
  }
}
static void node_ref(const Node& n) {
  if (n.ref_len >= 0) return;
  Bytes enc;
  switch (n.kind) {
    case EMPTY:
      n.ref[0] = 0x80;
      n.ref_len = 1;
      return;
    case HASH:
      memcpy(n.ref, n.hash.b, 32);
      n.ref_len = 32;
      return;
    case BRANCH: {
      Bytes pl;
      for (int i = 0; i < 16; i++) append_child_ref(pl, *n.ch[i]);
      if (n.value.empty())
        pl.push_back(0x80);
      else
        rlp_append_str(pl, n.value);
      enc = rlp_wrap_list(pl);
      break;
    }
    case EXT: {
      Bytes pl;
      rlp_append_str(pl, hex_prefix(n.nib, false));
      append_child_ref(pl, *n.child);
      enc = rlp_wrap_list(pl);
      break;
    }
    case LEAF: {
      Bytes pl;
      rlp_append_str(pl, hex_prefix(n.nib, true));
      rlp_append_str(pl, n.value);
      enc = rlp_wrap_list(pl);
      break;
    }
  }
  if (enc.size() >= 32) {
    g_stats.node_perms += keccak256_raw(enc.data(), enc.size(), n.ref);
    g_stats.nodes_hashed++;
    n.ref_len = 32;
  } else {
    memcpy(n.ref, enc.data(), enc.size());
    n.ref_len = (int8_t)enc.size();
  }
}
// PartialTrie::hash(): the root is always hashed, even when its encoding is < 32 bytes
static H256 trie_hash(const NodeP& root) {
  node_ref(*root);
  H256 h;
  if (root->ref_len == 32) {
    memcpy(h.b, root->ref, 32);
  } else {
    g_stats.node_perms += keccak256_raw(root->ref, (size_t)root->ref_len, h.b);
    g_stats.nodes_hashed++;
  }
  return h;
}

// --- insert (trie_ops.rs insert_into_trie_rec) ---
struct InsertVal {
  bool is_hash;
  Bytes val;
  H256 h;
};
static NodeP node_from_insert_val(const Nibs& k, const InsertVal& v) {
  if (!v.is_hash) return mk_leaf(k, v.val);
  // A hash inserted below its parent keeps the remaining nibbles as an extension (golden 6)
  return k.n == 0 ? mk_hash(v.h) : mk_ext(k, mk_hash(v.h));
}
static NodeP place_branch(const Nibs& common, const Nibs& existing_postfix, const NodeP& existing, const Nibs& new_postfix,
                          const InsertVal& v) {
  if (existing_postfix.n == 0 || new_postfix.n == 0) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "one key is a prefix of another");
  auto ch = empty_children();
  ch[existing_postfix.d[0]] = existing;
  ch[new_postfix.d[0]] = node_from_insert_val(new_postfix.slice(1), v);
  NodeP br = mk_branch(ch, Bytes());
  return common.n == 0 ? br : mk_ext(common, br);
}
static NodeP trie_insert(const NodeP& node, const Nibs& key, const InsertVal& v) {
  switch (node->kind) {
    case EMPTY:
      return node_from_insert_val(key, v);
    case HASH:
      fail(PPD_PANIC_INSERT_INTO_HASH_NODE, "insert traversed a Hash node");
    case BRANCH: {
      if (key.n == 0) {
        if (v.is_hash) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "hash inserted at a branch");
        return mk_branch(node->ch, v.val);
      }
      auto ch = node->ch;
      ch[key.d[0]] = trie_insert(node->ch[key.d[0]], key.slice(1), v);
      return mk_branch(ch, node->value);
    }
    case EXT: {
      int cp = common_prefix_len(node->nib, key);
      if (cp == node->nib.n) return mk_ext(node->nib, trie_insert(node->child, key.slice(cp), v));
      Nibs existing_postfix = node->nib.slice(cp);
      NodeP existing = existing_postfix.n == 1 ? node->child : mk_ext(existing_postfix.slice(1), node->child);
      return place_branch(key.prefix(cp), existing_postfix, existing, key.slice(cp), v);
    }
    case LEAF: {
      if (node->nib == key) {
        if (v.is_hash) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "hash inserted over a leaf");
        return mk_leaf(key, v.val);
      }
      int cp = common_prefix_len(node->nib, key);
      Nibs existing_postfix = node->nib.slice(cp);
      if (existing_postfix.n == 0) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "leaf key is a prefix of the new key");
      NodeP existing = mk_leaf(existing_postfix.slice(1), node->value);
      return place_branch(key.prefix(cp), existing_postfix, existing, key.slice(cp), v);
    }
  }
  return node;
}

// --- delete (trie_ops.rs delete_intern + collapse helpers; RECALLED) ---
static NodeP collapse_ext(const Nibs& ext, const NodeP& child) {
  switch (child->kind) {
    case EXT:
      return mk_ext(ext.merge(child->nib), child->child);
    case LEAF:
      return mk_leaf(ext.merge(child->nib), child->value);
    default:  // Branch, Hash
      return mk_ext(ext, child);
  }
}
// returns nullptr when the key is not present (trie unchanged)
static NodeP trie_delete(const NodeP& node, const Nibs& key) {
  switch (node->kind) {
    case EMPTY:
    case HASH:
      return nullptr;
    case LEAF:
      return node->nib == key ? EMPTY_NODE : nullptr;
    case EXT: {
      if (key.n < node->nib.n || common_prefix_len(node->nib, key) != node->nib.n) return nullptr;
      NodeP upd = trie_delete(node->child, key.slice(node->nib.n));
      if (!upd) return nullptr;
      if (upd->kind == EMPTY) return EMPTY_NODE;
      return collapse_ext(node->nib, upd);
    }
    case BRANCH: {
      if (key.n == 0) {
        if (node->value.empty()) return nullptr;
        return mk_branch(node->ch, Bytes());
      }
      uint8_t nib = key.d[0];
      NodeP upd = trie_delete(node->ch[nib], key.slice(1));
      if (!upd) return nullptr;
      auto ch = node->ch;
      ch[nib] = upd;
      if (upd->kind != EMPTY) return mk_branch(ch, node->value);
      int live = 0, last = -1;
      for (int i = 0; i < 16; i++)
        if (ch[i]->kind != EMPTY) live++, last = i;
      if (live >= 2 || !node->value.empty()) return mk_branch(ch, node->value);
      if (live == 0) return EMPTY_NODE;
      Nibs one;
      one.push((uint8_t)last);
      return collapse_ext(one, ch[last]);
    }
  }
  return nullptr;
}

static const Bytes* trie_get(const NodeP& node, const Nibs& key) {
  const Node* n = node.get();
  int pos = 0;
  for (;;) {
    switch (n->kind) {
      case EMPTY:
      case HASH:
        return nullptr;
      case LEAF:
        return (n->nib == key.slice(pos)) ? &n->value : nullptr;
      case EXT: {
        Nibs rest = key.slice(pos);
        if (rest.n < n->nib.n || common_prefix_len(n->nib, rest) != n->nib.n) return nullptr;
        pos += n->nib.n;
        n = n->child.get();
        break;
      }
      case BRANCH:
        if (pos == key.n) return n->value.empty() ? nullptr : &n->value;
        n = n->ch[key.d[pos]].get();
        pos++;
        break;
    }
  }
}

struct Item {
  Nibs key;
  bool is_hash;
  Bytes val;
  H256 h;
};
static void trie_items(const NodeP& node, const Nibs& prefix, std::vector<Item>& out) {
  switch (node->kind) {
    case EMPTY:
      return;
    case HASH:
      out.push_back({prefix, true, {}, node->hash});
      return;
    case LEAF:
      out.push_back({prefix.merge(node->nib), false, node->value, {}});
      return;
    case EXT:
      trie_items(node->child, prefix.merge(node->nib), out);
      return;
    case BRANCH:
      if (!node->value.empty()) out.push_back({prefix, false, node->value, {}});
      for (int i = 0; i < 16; i++) {
        Nibs p = prefix;
        p.push((uint8_t)i);
        trie_items(node->ch[i], p, out);
      }
      return;
  }
}

// --- create_trie_subset (trie_subsets.rs; RECALLED) ---
using Touched = std::unordered_set<const Node*>;
static void subset_mark(const NodeP& node, const Nibs& key, int pos, Touched& touched) {
  touched.insert(node.get());
  switch (node->kind) {
    case EMPTY:
    case LEAF:
      return;
    case HASH:
      if (pos < key.n) fail(PPD_ERR_MISSING_KEYS_CREATING_SUB_PARTIAL_TRIE, "subset key runs into a hashed-out node");
      return;
    case BRANCH:
      if (pos >= key.n) return;
      subset_mark(node->ch[key.d[pos]], key, pos + 1, touched);
      return;
    case EXT: {
      int avail = key.n - pos;
      int m = std::min<int>(avail, node->nib.n);
      if (memcmp(node->nib.d, key.d + pos, m) != 0) return;
      if (avail < node->nib.n) return;
      subset_mark(node->child, key, pos + node->nib.n, touched);
      return;
    }
  }
}
static NodeP subset_build(const NodeP& node, const Touched& touched) {
  if (!touched.count(node.get())) {
    node_ref(*node);
    if (node->ref_len == 32) {
      H256 h;
      memcpy(h.b, node->ref, 32);
      return mk_hash(h);
    }
    // too small to hash: kept as it is
  }
  switch (node->kind) {
    case EMPTY:
    case HASH:
    case LEAF:
      return node;
    case EXT:
      return mk_ext(node->nib, subset_build(node->child, touched));
    case BRANCH: {
      auto ch = node->ch;
      for (int i = 0; i < 16; i++) ch[i] = subset_build(node->ch[i], touched);
      return mk_branch(ch, node->value);
    }
  }
  return node;
}
static NodeP create_trie_subset(const NodeP& trie, const std::vector<Nibs>& keys) {
  Touched touched;
  for (const Nibs& k : keys) subset_mark(trie, k, 0, touched);
  return subset_build(trie, touched);
}

// ---------------------------------------------------------------------------------------------
// Compact witness: bytes -> instructions (compact_prestate_processing.rs:683-875, 896-1003)
// ---------------------------------------------------------------------------------------------
struct Instr {
  uint8_t op;
  Nibs key;        // leaf / extension / account leaf
  Bytes key_bytes; // raw compact key bytes (for the instruction dump)
  Bytes value;     // leaf value / code
  uint32_t mask = 0;
  H256 hash;
  uint8_t nonce[32], balance[32];
  bool has_code = false, has_storage = false;
};

struct Cursor {
  const uint8_t* p;
  size_t n, pos = 0;
  bool eof() const { return pos == n; }
  uint8_t read_byte() {
    if (pos >= n) fail(PPD_ERR_UNEXPECTED_END_OF_STREAM, "read_byte at end of stream");
    return p[pos++];
  }
  // CBOR head: major type + argument.  ciborium subset: definite lengths only.
  bool cbor_head(uint8_t& major, uint64_t& arg) {
    if (pos >= n) return false;
    uint8_t b = p[pos++];
    major = b >> 5;
    uint8_t ai = b & 31;
    if (ai < 24) {
      arg = ai;
      return true;
    }
    if (ai > 27) return false;
    size_t w = (size_t)1 << (ai - 24);
    if (n - pos < w) return false;
    arg = 0;
    for (size_t i = 0; i < w; i++) arg = (arg << 8) | p[pos++];
    return true;
  }
  Bytes read_cbor_bytes(int err_code, const char* field) {
    uint8_t major;
    uint64_t len;
    if (!cbor_head(major, len) || major != 2 || len > n - pos) fail(err_code, std::string("bad CBOR byte string: ") + field);
    Bytes out(p + pos, p + pos + len);
    pos += len;
    return out;
  }
  uint64_t read_cbor_uint(uint64_t max, const char* field) {
    uint8_t major;
    uint64_t v;
    if (!cbor_head(major, v) || major != 0 || v > max) fail(PPD_ERR_INVALID_BYTES_FOR_TYPE, std::string("bad CBOR uint: ") + field);
    return v;
  }
  H256 read_h256() {
    if (n - pos < 32) fail(PPD_ERR_INVALID_BYTES_FOR_TYPE, "short raw hash");
    H256 h;
    memcpy(h.b, p + pos, 32);
    pos += 32;
    return h;
  }
};

// compact_prestate_processing.rs:1338-1390
static Nibs key_bytes_to_nibbles(const Bytes& bytes) {
  Nibs key;
  if (bytes.empty()) return key;
  if (bytes.size() == 1) key.push(bytes[0] & 0x0f);
  bool is_odd = bytes[0] & 1;
  size_t m = bytes.size() - 1;  // actual key bytes
  if (m == 0) return key;
  for (size_t i = 0; i + 1 < m; i++) {
    key.push(bytes[1 + i] >> 4);
    key.push(bytes[1 + i] & 15);
  }
  uint8_t fin = bytes[m];
  key.push(fin >> 4);
  if (!is_odd) key.push(fin & 15);
  return key;
}

static uint8_t parse_instructions(const uint8_t* w, size_t n, std::vector<Instr>& out) {
  Cursor c{w, n};
  if (n == 0) fail(PPD_ERR_MISSING_HEADER, "missing header");
  uint8_t version = c.read_byte();
  while (!c.eof()) {
    Instr in;
    in.op = c.read_byte();
    switch (in.op) {
      case PPD_OP_LEAF:
        in.key_bytes = c.read_cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR, "leaf key");
        in.key = key_bytes_to_nibbles(in.key_bytes);
        in.value = c.read_cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR, "leaf value");
        break;
      case PPD_OP_EXTENSION:
        in.key_bytes = c.read_cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR, "extension key");
        in.key = key_bytes_to_nibbles(in.key_bytes);
        break;
      case PPD_OP_BRANCH:
        in.mask = (uint32_t)c.read_cbor_uint(0xffffffffull, "mask");
        break;
      case PPD_OP_HASH:
        in.hash = c.read_h256();
        break;
      case PPD_OP_CODE:
        in.value = c.read_cbor_bytes(PPD_ERR_INVALID_BYTES_FOR_TYPE, "code");
        break;
      case PPD_OP_ACCOUNT_LEAF: {
        in.key_bytes = c.read_cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR, "account leaf key");
        in.key = key_bytes_to_nibbles(in.key_bytes);
        uint8_t flags = c.read_byte();
        in.has_code = flags & 1;
        in.has_storage = flags & 2;
        memset(in.nonce, 0, 32);
        memset(in.balance, 0, 32);
        if (flags & 4) {
          uint64_t v = c.read_cbor_uint(~0ull, "account leaf nonce");
          for (int i = 0; i < 8; i++) in.nonce[31 - i] = (uint8_t)(v >> (8 * i));
        }
        if (flags & 8) {
          Bytes b = c.read_cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR, "account leaf balance");
          if (b.size() > 32) fail(PPD_PANIC_U256_FROM_BIG_ENDIAN, "balance wider than 256 bits (U256::from_big_endian panics)");
          memcpy(in.balance + 32 - b.size(), b.data(), b.size());
        }
        if (flags & 1) (void)c.read_cbor_uint(~0ull, "code size");
        break;
      }
      case PPD_OP_EMPTY_ROOT:
        break;
      default:
        fail(PPD_ERR_INVALID_OPERATOR, "invalid opcode");
    }
    out.push_back(std::move(in));
  }
  return version;
}

// ---------------------------------------------------------------------------------------------
// Instructions -> NodeEntry tree (compact_prestate_processing.rs:387-668) -> tries
// (compact_to_partial_trie.rs:37-190)
// ---------------------------------------------------------------------------------------------
struct CNode;
using CNodeP = std::shared_ptr<CNode>;
struct CNode {
  enum K { BRANCH, CODE, EMPTY, HASH, LEAF, EXT } k;
  std::array<CNodeP, 16> ch;  // BRANCH
  Bytes bytes;                // CODE bytes / LEAF raw value
  H256 hash;                  // HASH
  Nibs key;                   // LEAF / EXT
  CNodeP child;               // EXT
  bool is_account = false;    // LEAF
  uint8_t nonce[32], balance[32];
  bool has_storage_root = false, has_code = false, code_is_hash = false;
  H256 storage_root, code_hash_node;
  Bytes code_bytes;
};

struct TrieOut {
  NodeP trie = EMPTY_NODE;
  std::map<H256, Bytes> code;
};
static void compact_node_to_trie(const Nibs& key, const CNode& n, TrieOut& out) {
  switch (n.k) {
    case CNode::BRANCH:
      for (int i = 0; i < 16; i++)
        if (n.ch[i]) {
          Nibs k = key;
          k.push((uint8_t)i);
          compact_node_to_trie(k, *n.ch[i], out);
        }
      return;
    case CNode::CODE:
      out.code[hash_bytes(n.bytes)] = n.bytes;
      return;
    case CNode::EMPTY:
      return;
    case CNode::HASH:
      out.trie = trie_insert(out.trie, key, InsertVal{true, {}, n.hash});
      return;
    case CNode::EXT:
      compact_node_to_trie(key.merge(n.key), *n.child, out);
      return;
    case CNode::LEAF: {
      Nibs full = key.merge(n.key);
      Bytes val;
      if (!n.is_account) {
        rlp_append_str(val, n.bytes);
      } else {
        Account a;
        memcpy(a.nonce, n.nonce, 32);
        memcpy(a.balance, n.balance, 32);
        a.storage_root = n.has_storage_root ? n.storage_root : EMPTY_TRIE_HASH;
        if (n.has_code && !n.code_is_hash) {
          a.code_hash = hash_bytes(n.code_bytes);
          out.code[a.code_hash] = n.code_bytes;
        } else if (n.has_code) {
          a.code_hash = n.code_hash_node;
        } else {
          a.code_hash = EMPTY_CODE_HASH;
        }
        val = account_encode(a);
      }
      out.trie = trie_insert(out.trie, full, InsertVal{false, val, {}});
      return;
    }
  }
}

struct PreImage {
  uint8_t version = 0;
  NodeP state = EMPTY_NODE;
  std::map<H256, NodeP> storage;  // by hashed account address
  std::map<H256, Bytes> code;
};

static PreImage process_compact_prestate(const uint8_t* w, size_t n) {
  std::vector<Instr> instrs;
  PreImage pre;
  pre.version = parse_instructions(w, n, instrs);
  std::vector<CNodeP> stack;
  std::unordered_map<H256, NodeP, H256Hash> storage_by_root;
  for (const Instr& in : instrs) {
    auto node = std::make_shared<CNode>();
    switch (in.op) {
      case PPD_OP_EMPTY_ROOT:
        node->k = CNode::EMPTY;
        break;
      case PPD_OP_HASH:
        node->k = CNode::HASH;
        node->hash = in.hash;
        break;
      case PPD_OP_LEAF:
        node->k = CNode::LEAF;
        node->key = in.key;
        node->bytes = in.value;
        break;
      case PPD_OP_CODE:
        node->k = CNode::CODE;
        node->bytes = in.value;
        break;
      case PPD_OP_EXTENSION:
        if (stack.empty()) fail(PPD_ERR_INVALID_WITNESS_FORMAT, "extension with no preceding node");
        node->k = CNode::EXT;
        node->key = in.key;
        node->child = stack.back();
        stack.pop_back();
        break;
      case PPD_OP_BRANCH: {
        size_t expected = (size_t)__builtin_popcount(in.mask);
        if (stack.size() < expected) fail(PPD_ERR_INCORRECT_NUMBER_OF_NODES_PRECEDING_BRANCH, "branch mask wants more nodes than precede it");
        node->k = CNode::BRANCH;
        size_t base = stack.size() - expected, used = 0;
        for (int i = 0; i < 16; i++)
          if (in.mask & (1u << i)) node->ch[i] = stack[base + used++];  // lowest bit <-> oldest pushed
        if (used != expected) fail(PPD_ERR_MISSING_EXPECTED_NODES_PRECEDING_BRANCH, "branch mask has bits above 15");
        stack.resize(base);
        break;
      }
      case PPD_OP_ACCOUNT_LEAF: {
        node->k = CNode::LEAF;
        node->is_account = true;
        node->key = in.key;
        memcpy(node->nonce, in.nonce, 32);
        memcpy(node->balance, in.balance, 32);
        if (in.has_storage) {
          if (stack.empty() || stack.back()->k == CNode::CODE) fail(PPD_ERR_INVALID_WITNESS_FORMAT, "account leaf: no storage node");
          CNodeP s = stack.back();
          stack.pop_back();
          TrieOut st;
          compact_node_to_trie(Nibs(), *s, st);
          H256 root = trie_hash(st.trie);
          storage_by_root[root] = st.trie;
          node->has_storage_root = true;
          node->storage_root = root;
        }
        if (in.has_code) {
          if (stack.empty() || (stack.back()->k != CNode::CODE && stack.back()->k != CNode::HASH))
            fail(PPD_ERR_INVALID_WITNESS_FORMAT, "account leaf: no code node");
          CNodeP c = stack.back();
          stack.pop_back();
          node->has_code = true;
          if (c->k == CNode::CODE) {
            node->code_bytes = c->bytes;
          } else {
            node->code_is_hash = true;
            node->code_hash_node = c->hash;
          }
        }
        break;
      }
    }
    stack.push_back(node);
  }
  if (stack.size() > 1) fail(PPD_ERR_NON_SINGLE_ENTRY_AFTER_PROCESSING, "more than one entry left");
  if (stack.size() == 1) {
    TrieOut st;
    compact_node_to_trie(Nibs(), *stack[0], st);
    pre.state = st.trie;
    pre.code = st.code;
  }
  // convert_storage_trie_root_keyed_hashmap_to_account_addr_keyed (compact_to_partial_trie.rs:167-190)
  std::vector<Item> items;
  trie_items(pre.state, Nibs(), items);
  for (const Item& it : items) {
    if (it.is_hash) continue;
    Account a;
    if (!account_decode(it.val.data(), it.val.size(), a)) fail(PPD_PANIC_PRE_IMAGE_ACCOUNT_DECODE, "state leaf is not an account");
    auto f = storage_by_root.find(a.storage_root);
    if (f != storage_by_root.end()) pre.storage[h_addr_nibs_to_h256(it.key)] = f->second;
  }
  return pre;
}

// ---------------------------------------------------------------------------------------------
// Flat input reader
// ---------------------------------------------------------------------------------------------
struct Reader {
  const uint8_t* p;
  size_t n, pos = 0;
  void need(size_t k) {
    if (n - pos < k) fail(PPD_ERR_BAD_FLAT_INPUT, "flat block truncated");
  }
  uint8_t u8() {
    need(1);
    return p[pos++];
  }
  uint32_t u32() {
    need(4);
    uint32_t v;
    memcpy(&v, p + pos, 4);
    pos += 4;
    return v;
  }
  uint64_t u64() {
    need(8);
    uint64_t v;
    memcpy(&v, p + pos, 8);
    pos += 8;
    return v;
  }
  void raw(uint8_t* out, size_t k) {
    need(k);
    memcpy(out, p + pos, k);
    pos += k;
  }
  Bytes bytes() {
    uint32_t k = u32();
    need(k);
    Bytes b(p + pos, p + pos + k);
    pos += k;
    return b;
  }
};

// DirectPreImage (include/ppd_flat.h): Trie state_trie; u32 n_storage; n x { hashed_addr[32]; Trie }, tries in the
// pre-order Node form of the IrDump.  trace_protocol.rs:97-99 (TrieDirect(HashedPartialTrie)), :101-108 (MultipleTries).
static NodeP read_direct_node(Reader& r, int depth) {
  if (depth > 130) fail(PPD_ERR_BAD_FLAT_INPUT, "direct trie nested deeper than any 64-nibble key allows");
  switch (r.u8()) {
    case PPD_NODE_EMPTY:
      return EMPTY_NODE;
    case PPD_NODE_HASH: {
      H256 h;
      r.raw(h.b, 32);
      return mk_hash(h);
    }
    case PPD_NODE_BRANCH: {
      std::array<NodeP, 16> ch;
      for (int i = 0; i < 16; i++) ch[i] = read_direct_node(r, depth + 1);
      return mk_branch(ch, r.bytes());
    }
    case PPD_NODE_EXTENSION: {
      Nibs k;
      uint8_t cnt = r.u8();
      if (cnt > 64) fail(PPD_ERR_BAD_FLAT_INPUT, "direct trie: more than 64 nibbles");
      for (int i = 0; i < cnt; i++) {
        uint8_t v = r.u8();
        if (v > 15) fail(PPD_ERR_BAD_FLAT_INPUT, "direct trie: nibble above 15");
        k.push(v);
      }
      return mk_ext(k, read_direct_node(r, depth + 1));
    }
    case PPD_NODE_LEAF: {
      Nibs k;
      uint8_t cnt = r.u8();
      if (cnt > 64) fail(PPD_ERR_BAD_FLAT_INPUT, "direct trie: more than 64 nibbles");
      for (int i = 0; i < cnt; i++) {
        uint8_t v = r.u8();
        if (v > 15) fail(PPD_ERR_BAD_FLAT_INPUT, "direct trie: nibble above 15");
        k.push(v);
      }
      return mk_leaf(k, r.bytes());
    }
    default:
      fail(PPD_ERR_BAD_FLAT_INPUT, "direct trie: unknown node kind");
  }
  return EMPTY_NODE;
}
// process_separate_trie_pre_images (processed_block_trace.rs:130-141): tries as given, no code mappings
static PreImage process_direct_pre_image(const uint8_t* p, size_t n) {
  Reader r{p, n};
  PreImage pre;
  pre.version = 1;
  pre.state = read_direct_node(r, 0);
  uint32_t ns = r.u32();
  for (uint32_t i = 0; i < ns; i++) {
    H256 h;
    r.raw(h.b, 32);
    pre.storage[h] = read_direct_node(r, 0);
  }
  if (r.pos != r.n) fail(PPD_ERR_BAD_FLAT_INPUT, "bytes after the direct pre-image");
  return pre;
}

struct Addr {
  uint8_t b[20];
};
struct Trace {
  Addr addr;
  uint8_t flags;
  uint8_t balance[32], nonce[32];
  std::vector<H256> reads;
  std::vector<std::pair<H256, std::array<uint8_t, 32>>> writes;
  H256 code_read;
  Bytes code_write;
};
struct Txn {
  std::vector<Trace> traces;
  Bytes byte_code, new_txn_node, new_receipt_node;
  uint64_t gas_used;
};
struct Block {
  uint32_t pre_image_kind = 0;  // 0 Combined{compact}; 2 Separate{Direct, MultipleTries{Direct}} (include/ppd_flat.h)
  Bytes compact;                // the TrieCompact bytes, or the DirectPreImage payload
  std::vector<Txn> txns;
  std::map<H256, Bytes> resolved_code;
  std::vector<std::pair<Addr, std::array<uint8_t, 32>>> withdrawals;
  H256 checkpoint;
  Bytes b_meta, b_hashes;
};
static Block read_block(const uint8_t* p, size_t n) {
  Reader r{p, n};
  Block b;
  if (r.u32() != PPD_FLAT_BLOCK_MAGIC || r.u32() != 1) fail(PPD_ERR_BAD_FLAT_INPUT, "bad magic/version");
  b.pre_image_kind = r.u32();
  // processed_block_trace.rs:130-168: Separate{state: Direct} is `t.0`; every other Separate form ends in todo!().
  // Kind 2 (Direct state trie + a Direct trie per hashed address) is this repo's completion of
  // process_multiple_storage_tries -- each entry taken as the trie it holds, as process_state_trie does.
  if (b.pre_image_kind != 0 && b.pre_image_kind != 2)
    fail(PPD_PANIC_UNIMPLEMENTED_PRE_IMAGE, "pre-image variant the reference leaves as todo!()");
  b.compact = r.bytes();
  uint32_t nt = r.u32();
  for (uint32_t t = 0; t < nt; t++) {
    Txn tx;
    uint32_t ntr = r.u32();
    for (uint32_t i = 0; i < ntr; i++) {
      Trace tr;
      r.raw(tr.addr.b, 20);
      tr.flags = r.u8();
      if (tr.flags & PPD_TR_BALANCE) r.raw(tr.balance, 32);
      if (tr.flags & PPD_TR_NONCE) r.raw(tr.nonce, 32);
      if (tr.flags & PPD_TR_STORAGE_READ) {
        uint32_t k = r.u32();
        for (uint32_t j = 0; j < k; j++) {
          H256 h;
          r.raw(h.b, 32);
          tr.reads.push_back(h);
        }
      }
      if (tr.flags & PPD_TR_STORAGE_WRITTEN) {
        uint32_t k = r.u32();
        for (uint32_t j = 0; j < k; j++) {
          H256 h;
          std::array<uint8_t, 32> v;
          r.raw(h.b, 32);
          r.raw(v.data(), 32);
          tr.writes.push_back({h, v});
        }
      }
      if (tr.flags & PPD_TR_CODE_READ) r.raw(tr.code_read.b, 32);
      if (tr.flags & PPD_TR_CODE_WRITE) tr.code_write = r.bytes();
      tx.traces.push_back(std::move(tr));
    }
    tx.byte_code = r.bytes();
    tx.new_txn_node = r.bytes();
    tx.new_receipt_node = r.bytes();
    tx.gas_used = r.u64();
    b.txns.push_back(std::move(tx));
  }
  uint32_t nc = r.u32();
  for (uint32_t i = 0; i < nc; i++) {
    H256 h;
    r.raw(h.b, 32);
    b.resolved_code[h] = r.bytes();
  }
  uint32_t nw = r.u32();
  for (uint32_t i = 0; i < nw; i++) {
    Addr a;
    std::array<uint8_t, 32> v;
    r.raw(a.b, 20);
    r.raw(v.data(), 32);
    b.withdrawals.push_back({a, v});
  }
  r.raw(b.checkpoint.b, 32);
  b.b_meta = r.bytes();
  b.b_hashes = r.bytes();
  return b;
}

// ---------------------------------------------------------------------------------------------
// processed_block_trace.rs:209-343  TxnInfo::into_processed_txn_info
// ---------------------------------------------------------------------------------------------
struct StateWrite {
  H256 haddr;
  bool has_balance, has_nonce, storage_trie_change, has_code_hash;
  uint8_t balance[32], nonce[32];
  H256 code_hash;
};
struct NodesUsedByTxn {
  std::vector<H256> state_accesses;
  std::vector<StateWrite> state_writes;
  std::vector<std::pair<Nibs, std::vector<Nibs>>> storage_accesses;
  std::vector<std::pair<Nibs, std::vector<std::pair<Nibs, Bytes>>>> storage_writes;
  std::map<H256, H256> accounts_with_storage_but_no_accesses;
  std::vector<H256> self_destructed;
};
struct ProcessedTxn {
  NodesUsedByTxn nodes;
  std::map<H256, Bytes> contract_code;
  bool has_txn_bytes;
  Bytes txn_bytes, receipt_node_bytes;
  uint64_t gas_used;
};

static ProcessedTxn process_txn(const Txn& tx, const std::vector<std::pair<H256, Account>>& all_accounts, const PreImage& pre,
                                const Block& blk) {
  ProcessedTxn out;
  out.contract_code[EMPTY_CODE_HASH] = Bytes();
  std::set<H256> with_storage_accesses;
  for (const Trace& tr : tx.traces) {
    H256 haddr = hash_bytes(tr.addr.b, 20);
    Nibs haddr_nibs = nibs_from_h256(haddr);
    std::vector<Nibs> access_keys;
    for (const H256& k : tr.reads) access_keys.push_back(nibs_from_h256(hash_bytes(k.b, 32)));
    for (const auto& w : tr.writes) access_keys.push_back(nibs_from_h256(hash_bytes(w.first.b, 32)));
    if (!access_keys.empty()) with_storage_accesses.insert(haddr);
    out.nodes.storage_accesses.push_back({haddr_nibs, access_keys});

    bool storage_trie_change = !tr.writes.empty();
    bool code_change = tr.flags & (PPD_TR_CODE_READ | PPD_TR_CODE_WRITE);
    if ((tr.flags & (PPD_TR_BALANCE | PPD_TR_NONCE)) || storage_trie_change || code_change) {
      StateWrite sw;
      sw.haddr = haddr;
      sw.has_balance = tr.flags & PPD_TR_BALANCE;
      sw.has_nonce = tr.flags & PPD_TR_NONCE;
      memcpy(sw.balance, tr.balance, 32);
      memcpy(sw.nonce, tr.nonce, 32);
      sw.storage_trie_change = storage_trie_change;
      sw.has_code_hash = code_change;
      if (tr.flags & PPD_TR_CODE_READ)
        sw.code_hash = tr.code_read;
      else if (tr.flags & PPD_TR_CODE_WRITE)
        sw.code_hash = hash_bytes(tr.code_write);
      out.nodes.state_writes.push_back(sw);
    }
    std::vector<std::pair<Nibs, Bytes>> wr;
    for (const auto& w : tr.writes) wr.push_back({nibs_from_h256(w.first), rlp_encode_u256(w.second.data())});
    out.nodes.storage_writes.push_back({haddr_nibs, wr});
    out.nodes.state_accesses.push_back(haddr);

    if (tr.flags & PPD_TR_CODE_READ) {
      if (!out.contract_code.count(tr.code_read)) {
        auto f = pre.code.find(tr.code_read);
        if (f != pre.code.end()) {
          out.contract_code[tr.code_read] = f->second;
        } else {
          auto g = blk.resolved_code.find(tr.code_read);
          if (g == blk.resolved_code.end()) fail(PPD_ERR_UNRESOLVED_CODE_HASH, "code hash not resolvable");
          out.contract_code[tr.code_read] = g->second;
        }
      }
    } else if (tr.flags & PPD_TR_CODE_WRITE) {
      out.contract_code[hash_bytes(tr.code_write)] = tr.code_write;
    }
    if (tr.flags & PPD_TR_SELF_DESTRUCTED) out.nodes.self_destructed.push_back(haddr);
  }
  for (const auto& acc : all_accounts)
    if (acc.second.storage_root != EMPTY_TRIE_HASH && !with_storage_accesses.count(acc.first))
      out.nodes.accounts_with_storage_but_no_accesses[acc.first] = acc.second.storage_root;

  out.has_txn_bytes = !tx.byte_code.empty();
  out.txn_bytes = tx.byte_code;
  // process_rlped_receipt_node_bytes, processed_block_trace.rs:335-343
  const Bytes& raw = tx.new_receipt_node;
  if (is_legacy_receipt(raw.data(), raw.size())) {
    out.receipt_node_bytes = raw;
  } else {
    RlpItem it;
    if (!rlp_decode_item(raw.data(), raw.size(), it) || it.is_list) fail(PPD_PANIC_RECEIPT_DECODE, "receipt is neither legacy nor a byte string");
    out.receipt_node_bytes.assign(it.payload, it.payload + it.payload_len);
  }
  out.gas_used = tx.gas_used;
  return out;
}

// ---------------------------------------------------------------------------------------------
// decoding.rs:80-607
// ---------------------------------------------------------------------------------------------
struct TrieState {
  NodeP state = EMPTY_NODE;
  std::map<H256, NodeP> storage;
  NodeP txn = EMPTY_NODE, receipt = EMPTY_NODE;
};
struct TrieInputs {
  NodeP state, txn, receipt;
  std::vector<std::pair<H256, NodeP>> storage;
};
struct GenInputs {
  uint64_t txn_number_before, gas_used_before, gas_used_after;
  bool has_signed_txn = false;
  Bytes signed_txn;
  std::vector<std::pair<Addr, std::array<uint8_t, 32>>> withdrawals;
  TrieInputs tries;
  H256 state_root, txn_root, receipt_root;
  std::map<H256, Bytes> contract_code;
};

static Nibs txn_key(size_t idx) {
  Bytes k = rlp_encode_u64(idx);
  return nibs_from_bytes(k.data(), k.size());
}
// The key `0_u64` used by create_fully_hashed_out_sub_partial_trie (decoding.rs:468-471) converts
// to zero nibbles (count = ceil(bits/4) = 0): the root is marked and nothing below it.  RECALLED.
static std::vector<Nibs> dummy_subset_keys() { return {Nibs()}; }

// create_trie_subset_wrapped (decoding.rs:595-602): the SubsetTrieError becomes MissingKeysCreatingSubPartialTrie(trie_type)
static NodeP subset_wrapped(const NodeP& trie, const std::vector<Nibs>& keys, const char* trie_type) {
  try {
    return create_trie_subset(trie, keys);
  } catch (const Err& e) {
    if (e.code == PPD_ERR_MISSING_KEYS_CREATING_SUB_PARTIAL_TRIE) fail(e.code, e.msg + "; trie_type=" + trie_type);
    throw;
  }
}

static TrieInputs minimal_tries_for_txn(TrieState& cur, const NodesUsedByTxn& nodes, size_t txn_idx) {
  TrieInputs ti;
  std::vector<Nibs> state_keys;
  for (const H256& h : nodes.state_accesses) state_keys.push_back(nibs_from_h256(h));
  ti.state = subset_wrapped(cur.state, state_keys, "State");
  ti.txn = subset_wrapped(cur.txn, {txn_key(txn_idx)}, "Txn");
  ti.receipt = subset_wrapped(cur.receipt, {txn_key(txn_idx)}, "Receipt");
  // decoding.rs:199-203: storage_access_vec is collected (every key through H256::from_slice, the panic site) before
  // create_minimal_storage_partial_tries cuts the first storage subset
  for (const auto& acc : nodes.storage_accesses) (void)h256_from_slice_of_nibs(acc.first);
  for (const auto& acc : nodes.storage_accesses) {
    H256 haddr = h256_from_slice_of_nibs(acc.first);
    auto f = cur.storage.find(haddr);
    if (f == cur.storage.end()) {
      auto g = nodes.accounts_with_storage_but_no_accesses.find(haddr);
      NodeP t = g != nodes.accounts_with_storage_but_no_accesses.end() ? mk_hash(g->second) : EMPTY_NODE;
      f = cur.storage.insert({haddr, t}).first;
    }
    ti.storage.push_back({haddr, subset_wrapped(f->second, acc.second, "Storage")});
  }
  return ti;
}

static void apply_deltas(TrieState& ts, const ProcessedTxn& tx, size_t txn_idx) {
  for (const auto& sw : tx.nodes.storage_writes) {
    H256 haddr = h256_from_slice_of_nibs(sw.first);
    auto f = ts.storage.find(haddr);
    if (f == ts.storage.end()) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, "no storage trie for a written account; hashed_addr=" + hex_str(haddr.b, 32));
    for (const auto& kv : sw.second) {
      Bytes pre = nibs_bytes_be_minimal(kv.first);
      Nibs slot = nibs_from_h256(hash_bytes(pre));
      if (kv.second.size() == 1 && kv.second[0] == 0x80) {
        NodeP d = trie_delete(f->second, slot);
        if (d) f->second = d;
      } else {
        f->second = trie_insert(f->second, slot, InsertVal{false, kv.second, {}});
      }
    }
  }
  static const Bytes EMPTY_ACCOUNT = empty_account_bytes();
  for (const StateWrite& w : tx.nodes.state_writes) {
    Nibs k = nibs_from_h256(w.haddr);
    const Bytes* cur = trie_get(ts.state, k);
    if (!cur) cur = &EMPTY_ACCOUNT;
    Account a;
    if (!account_decode(cur->data(), cur->size(), a)) fail(PPD_ERR_ACCOUNT_DECODE, "state leaf is not an account; bytes=" + hex_str(cur->data(), cur->size()));
    if (w.storage_trie_change) {
      auto f = ts.storage.find(w.haddr);
      if (f == ts.storage.end()) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, "no storage trie for a changed account; hashed_addr=" + hex_str(w.haddr.b, 32));
      a.storage_root = trie_hash(f->second);
    }
    if (w.has_balance) memcpy(a.balance, w.balance, 32);
    if (w.has_nonce) memcpy(a.nonce, w.nonce, 32);
    if (w.has_code_hash) a.code_hash = w.code_hash;
    ts.state = trie_insert(ts.state, k, InsertVal{false, account_encode(a), {}});
  }
  for (const H256& h : tx.nodes.self_destructed) {
    if (!ts.storage.erase(h)) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, "self-destructed account has no storage trie; hashed_addr=" + hex_str(h.b, 32));
    NodeP d = trie_delete(ts.state, nibs_from_h256(h));
    if (d) ts.state = d;
  }
  Nibs tk = txn_key(txn_idx);
  ts.txn = trie_insert(ts.txn, tk, InsertVal{false, tx.has_txn_bytes ? tx.txn_bytes : Bytes(), {}});
  ts.receipt = trie_insert(ts.receipt, tk, InsertVal{false, tx.receipt_node_bytes, {}});
}

static GenInputs dummy_gen_input(const TrieState& ts, uint64_t txn_number, uint64_t gas_used) {
  GenInputs g;
  auto keys = dummy_subset_keys();
  g.tries.state = create_trie_subset(ts.state, keys);
  g.tries.txn = create_trie_subset(ts.txn, keys);
  g.tries.receipt = create_trie_subset(ts.receipt, keys);
  for (const auto& s : ts.storage) g.tries.storage.push_back({s.first, create_trie_subset(s.second, keys)});
  g.state_root = trie_hash(g.tries.state);
  g.txn_root = trie_hash(g.tries.txn);
  g.receipt_root = trie_hash(g.tries.receipt);
  g.txn_number_before = txn_number;
  g.gas_used_before = gas_used;
  g.gas_used_after = gas_used;
  return g;
}

static void u256_add(uint8_t a[32], const uint8_t b[32]) {
  unsigned carry = 0;
  for (int i = 31; i >= 0; i--) {
    unsigned s = (unsigned)a[i] + b[i] + carry;
    a[i] = (uint8_t)s;
    carry = s >> 8;
  }
}
static void apply_withdrawals(const Block& blk, NodeP& state) {
  for (const auto& w : blk.withdrawals) {
    H256 h = hash_bytes(w.first.b, 20);
    Nibs k = nibs_from_h256(h);
    const Bytes* cur = trie_get(state, k);
    if (!cur) fail(PPD_ERR_MISSING_WITHDRAWAL_ACCOUNT, "withdrawal to an account that is not in the state trie; addr=" + hex_str(w.first.b, 20) + " hashed_addr=" + hex_str(h.b, 32) + " amount=" + hex_str(w.second.data(), 32));
    Account a;
    if (!account_decode(cur->data(), cur->size(), a)) fail(PPD_ERR_ACCOUNT_DECODE, "withdrawal account does not decode; bytes=" + hex_str(cur->data(), cur->size()));
    u256_add(a.balance, w.second.data());
    state = trie_insert(state, k, InsertVal{false, account_encode(a), {}});
  }
}

static std::vector<GenInputs> decode_block(const Block& blk) {
  PreImage pre = blk.pre_image_kind == 2 ? process_direct_pre_image(blk.compact.data(), blk.compact.size())
                                         : process_compact_prestate(blk.compact.data(), blk.compact.size());
  if (pre.version != 1) fail(PPD_PANIC_INCOMPATIBLE_HEADER_VERSION, "compact header version is not 1");

  std::vector<std::pair<H256, Account>> all_accounts;
  {
    std::vector<Item> items;
    trie_items(pre.state, Nibs(), items);
    for (const Item& it : items) {
      if (it.is_hash) continue;
      Account a;
      if (!account_decode(it.val.data(), it.val.size(), a)) fail(PPD_PANIC_PRE_IMAGE_ACCOUNT_DECODE, "state leaf is not an account");
      all_accounts.push_back({h_addr_nibs_to_h256(it.key), a});
    }
  }
  std::vector<ProcessedTxn> txns;
  for (const Txn& t : blk.txns) txns.push_back(process_txn(t, all_accounts, pre, blk));

  TrieState cur, initial;
  cur.state = initial.state = pre.state;
  cur.storage = initial.storage = pre.storage;
  uint64_t txn_before = 0, txn_after = 0, gas_before = 0, gas_after = 0;
  std::vector<GenInputs> irs;
  for (size_t i = 0; i < txns.size(); i++) {
    GenInputs g;
    g.tries = minimal_tries_for_txn(cur, txns[i].nodes, i);
    txn_after += 1;
    gas_after += txns[i].gas_used;
    apply_deltas(cur, txns[i], i);
    g.state_root = trie_hash(cur.state);
    g.txn_root = trie_hash(cur.txn);
    g.receipt_root = trie_hash(cur.receipt);
    g.txn_number_before = txn_before;
    g.gas_used_before = gas_before;
    g.gas_used_after = gas_after;
    g.has_signed_txn = txns[i].has_txn_bytes;
    g.signed_txn = txns[i].txn_bytes;
    g.contract_code = txns[i].contract_code;
    txn_before += 1;
    gas_before = gas_after;
    irs.push_back(std::move(g));
  }
  // pad_gen_inputs_with_dummy_inputs_if_needed, decoding.rs:304-347.  The dummy sanity asserts
  // (:498-505) compare txn_number_before/after and gas_used_before/after of `extra_data`, which are
  // equal after the loop; a dummy carries those values.
  bool has_withdrawals = !blk.withdrawals.empty();
  bool dummies_added = false;
  if (irs.empty()) {
    irs.push_back(dummy_gen_input(initial, txn_before, gas_before));
    irs.push_back(dummy_gen_input(initial, txn_before, gas_before));
    dummies_added = true;
  } else if (irs.size() == 1) {
    if (!has_withdrawals)
      irs.insert(irs.begin(), dummy_gen_input(initial, txn_before, gas_before));
    else
      irs.push_back(dummy_gen_input(cur, txn_before, gas_before));
    dummies_added = true;
  }
  // add_withdrawals_to_txns, decoding.rs:356-402
  if (has_withdrawals) {
    if (!dummies_added) {
      GenInputs wd = dummy_gen_input(cur, txn_before, gas_before);
      apply_withdrawals(blk, cur.state);
      wd.withdrawals = blk.withdrawals;
      wd.state_root = trie_hash(cur.state);
      irs.push_back(std::move(wd));
    } else {
      apply_withdrawals(blk, cur.state);
      irs[1].withdrawals = blk.withdrawals;
      irs[1].state_root = trie_hash(cur.state);
    }
  }
  return irs;
}

// ---------------------------------------------------------------------------------------------
// Canonical dumps (include/ppd_flat.h)
// ---------------------------------------------------------------------------------------------
struct Writer {
  Bytes b;
  void u8(uint8_t v) { b.push_back(v); }
  void u32(uint32_t v) {
    for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i)));
  }
  void u64(uint64_t v) {
    for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i)));
  }
  void raw(const uint8_t* p, size_t n) { b.insert(b.end(), p, p + n); }
  void bytes(const Bytes& x) {
    u32((uint32_t)x.size());
    raw(x.data(), x.size());
  }
  void u256_from_u64(uint64_t v) {
    uint8_t be[32];
    memset(be, 0, 32);
    for (int i = 0; i < 8; i++) be[31 - i] = (uint8_t)(v >> (8 * i));
    raw(be, 32);
  }
};
static void dump_node(Writer& w, const Node& n) {
  switch (n.kind) {
    case EMPTY:
      w.u8(PPD_NODE_EMPTY);
      return;
    case HASH:
      w.u8(PPD_NODE_HASH);
      w.raw(n.hash.b, 32);
      return;
    case BRANCH:
      w.u8(PPD_NODE_BRANCH);
      for (int i = 0; i < 16; i++) dump_node(w, *n.ch[i]);
      w.bytes(n.value);
      return;
    case EXT:
      w.u8(PPD_NODE_EXTENSION);
      w.u8(n.nib.n);
      w.raw(n.nib.d, n.nib.n);
      dump_node(w, *n.child);
      return;
    case LEAF:
      w.u8(PPD_NODE_LEAF);
      w.u8(n.nib.n);
      w.raw(n.nib.d, n.nib.n);
      w.bytes(n.value);
      return;
  }
}
static Bytes dump_irs(const std::vector<GenInputs>& irs, const Block& blk) {
  Writer w;
  w.u32(PPD_IR_DUMP_MAGIC);
  w.u32((uint32_t)irs.size());
  for (const GenInputs& g : irs) {
    w.u256_from_u64(g.txn_number_before);
    w.u256_from_u64(g.gas_used_before);
    w.u256_from_u64(g.gas_used_after);
    w.u8(g.has_signed_txn);
    w.bytes(g.has_signed_txn ? g.signed_txn : Bytes());
    w.u32((uint32_t)g.withdrawals.size());
    for (const auto& x : g.withdrawals) {
      w.raw(x.first.b, 20);
      w.raw(x.second.data(), 32);
    }
    dump_node(w, *g.tries.state);
    dump_node(w, *g.tries.txn);
    dump_node(w, *g.tries.receipt);
    auto st = g.tries.storage;
    std::stable_sort(st.begin(), st.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    w.u32((uint32_t)st.size());
    for (const auto& s : st) {
      w.raw(s.first.b, 32);
      dump_node(w, *s.second);
    }
    w.raw(g.state_root.b, 32);
    w.raw(g.txn_root.b, 32);
    w.raw(g.receipt_root.b, 32);
    w.raw(blk.checkpoint.b, 32);
    w.u32((uint32_t)g.contract_code.size());
    for (const auto& c : g.contract_code) {
      w.raw(c.first.b, 32);
      w.bytes(c.second);
    }
    w.bytes(blk.b_meta);
    w.bytes(blk.b_hashes);
  }
  return w.b;
}

static Bytes dump_pre_image(const PreImage& pre) {
  Writer w;
  w.u32(PPD_PRE_IMAGE_MAGIC);
  w.u8(pre.version);
  Stats before = g_stats;
  H256 root = trie_hash(pre.state);
  (void)before;
  w.raw(root.b, 32);
  w.u32((uint32_t)pre.storage.size());
  for (const auto& s : pre.storage) {
    w.raw(s.first.b, 32);
    H256 r = trie_hash(s.second);
    w.raw(r.b, 32);
  }
  w.u32((uint32_t)pre.code.size());
  for (const auto& c : pre.code) {
    w.raw(c.first.b, 32);
    w.u32((uint32_t)c.second.size());
  }
  w.u64(g_stats.nodes_hashed);
  w.u64(g_stats.node_perms);
  return w.b;
}

static Bytes dump_instructions(const std::vector<Instr>& ins, uint8_t version) {
  // u8 version, u32 n, then per instruction: u8 op and its operands:
  //   leaf: u8 n_nibbles, nibbles, u32 len, value   extension: u8 n, nibbles   branch: u32 mask
  //   hash: 32 B   code: u32 len, bytes   account leaf: u8 n, nibbles, nonce[32], balance[32], u8 has_code, u8 has_storage
  Writer w;
  w.u8(version);
  w.u32((uint32_t)ins.size());
  for (const Instr& in : ins) {
    w.u8(in.op);
    switch (in.op) {
      case PPD_OP_LEAF:
        w.u8(in.key.n);
        w.raw(in.key.d, in.key.n);
        w.bytes(in.value);
        break;
      case PPD_OP_EXTENSION:
        w.u8(in.key.n);
        w.raw(in.key.d, in.key.n);
        break;
      case PPD_OP_BRANCH:
        w.u32(in.mask);
        break;
      case PPD_OP_HASH:
        w.raw(in.hash.b, 32);
        break;
      case PPD_OP_CODE:
        w.bytes(in.value);
        break;
      case PPD_OP_ACCOUNT_LEAF:
        w.u8(in.key.n);
        w.raw(in.key.d, in.key.n);
        w.raw(in.nonce, 32);
        w.raw(in.balance, 32);
        w.u8(in.has_code);
        w.u8(in.has_storage);
        break;
    }
  }
  return w.b;
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------
// C ABI for ctypes (tests, smoke, bench cpu_baseline)
// ---------------------------------------------------------------------------------------------
using namespace orc;

static int finish(const Bytes& b, uint8_t** out, size_t* out_len) {
  *out = (uint8_t*)malloc(b.size() ? b.size() : 1);
  memcpy(*out, b.data(), b.size());
  *out_len = b.size();
  return PPD_OK;
}
template <class F>
static int guarded(F f, char* err, size_t err_cap) {
  try {
    return f();
  } catch (const Err& e) {
    if (err && err_cap) snprintf(err, err_cap, "%s", e.msg.c_str());
    return e.code;
  } catch (const std::exception& e) {
    if (err && err_cap) snprintf(err, err_cap, "%s", e.what());
    return PPD_ERR_BAD_ARGUMENT;
  }
}

extern "C" {

void oracle_free(void* p) { free(p); }

int oracle_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
  keccak256_raw(data, len, out);
  return PPD_OK;
}

// n messages, message i = data[offsets[i] .. offsets[i+1])
int oracle_keccak256_batch(const uint8_t* data, const uint64_t* offsets, size_t n, uint8_t* out32n) {
  for (size_t i = 0; i < n; i++) keccak256_raw(data + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), out32n + 32 * i);
  return PPD_OK;
}

int oracle_key_bytes_to_nibbles(const uint8_t* key, size_t len, uint8_t* nibbles_out, size_t* n_out) {
  return guarded(
      [&] {
        Nibs k = key_bytes_to_nibbles(Bytes(key, key + len));
        memcpy(nibbles_out, k.d, k.n);
        *n_out = k.n;
        return PPD_OK;
      },
      nullptr, 0);
}

int oracle_compact_instructions(const uint8_t* w, size_t n, uint8_t** out, size_t* out_len, char* err, size_t err_cap) {
  return guarded(
      [&] {
        std::vector<Instr> ins;
        uint8_t v = parse_instructions(w, n, ins);
        return finish(dump_instructions(ins, v), out, out_len);
      },
      err, err_cap);
}

// process_compact_prestate -> PreImageDump (include/ppd_flat.h)
int oracle_compact_decode(const uint8_t* w, size_t n, uint8_t** out, size_t* out_len, char* err, size_t err_cap) {
  return guarded(
      [&] {
        g_stats = Stats();
        PreImage pre = process_compact_prestate(w, n);
        return finish(dump_pre_image(pre), out, out_len);
      },
      err, err_cap);
}

// process_compact_prestate -> the same tries as a DirectPreImage payload (include/ppd_flat.h): what a tracer that
// sends Separate{Direct} pre-images would send for this state.  Test infrastructure for the kind-2 FlatBlock.
int oracle_compact_to_direct(const uint8_t* w, size_t n, uint8_t** out, size_t* out_len, char* err, size_t err_cap) {
  return guarded(
      [&] {
        PreImage pre = process_compact_prestate(w, n);
        Writer wr;
        dump_node(wr, *pre.state);
        wr.u32((uint32_t)pre.storage.size());
        for (const auto& s2 : pre.storage) {
          wr.raw(s2.first.b, 32);
          dump_node(wr, *s2.second);
        }
        return finish(wr.b, out, out_len);
      },
      err, err_cap);
}

// BlockTrace::into_txn_proof_gen_ir on a FlatBlock -> IrDump (include/ppd_flat.h)
int oracle_block_decode(const uint8_t* flat, size_t n, uint8_t** out, size_t* out_len, char* err, size_t err_cap) {
  return guarded(
      [&] {
        g_stats = Stats();
        Block blk = read_block(flat, n);
        std::vector<GenInputs> irs = decode_block(blk);
        return finish(dump_irs(irs, blk), out, out_len);
      },
      err, err_cap);
}

// Root of the trie holding n (key32, value) leaves, built by repeated insert like the reference
// builds every trie.  values[i] = vals[val_off[i] .. val_off[i+1]) is stored as given.
int oracle_trie_root_from_leaves(const uint8_t* keys32, const uint64_t* val_off, const uint8_t* vals, size_t n, uint8_t root_out[32]) {
  return guarded(
      [&] {
        g_stats = Stats();
        NodeP t = EMPTY_NODE;
        for (size_t i = 0; i < n; i++) {
          Nibs k = nibs_from_bytes(keys32 + 32 * i, 32);
          t = trie_insert(t, k, InsertVal{false, Bytes(vals + val_off[i], vals + val_off[i + 1]), {}});
        }
        H256 r = trie_hash(t);
        memcpy(root_out, r.b, 32);
        return PPD_OK;
      },
      nullptr, 0);
}

// counters of the last call on this thread: nodes_hashed, node_perms, other_hashes, other_perms
void oracle_last_stats(uint64_t out[4]) {
  out[0] = g_stats.nodes_hashed;
  out[1] = g_stats.node_perms;
  out[2] = g_stats.other_hashes;
  out[3] = g_stats.other_perms;
}

int oracle_rlp_account(const uint8_t nonce[32], const uint8_t balance[32], const uint8_t storage_root[32], const uint8_t code_hash[32],
                       uint8_t* out, size_t* out_len) {
  Account a;
  memcpy(a.nonce, nonce, 32);
  memcpy(a.balance, balance, 32);
  memcpy(a.storage_root.b, storage_root, 32);
  memcpy(a.code_hash.b, code_hash, 32);
  Bytes b = account_encode(a);
  memcpy(out, b.data(), b.size());
  *out_len = b.size();
  return PPD_OK;
}

}  // extern "C"
