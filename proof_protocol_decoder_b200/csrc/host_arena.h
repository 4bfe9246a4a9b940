// host_arena.h — host-side construction of the node arena: persistent (path-copying) Merkle
// Patricia tries whose every version stays addressable, so that ALL versions of ALL tries of a
// block (or of a batch of blocks) are hashed by one level-synchronous GPU sweep.
//
// Replaces eth_trie_utils' insert / delete / get / create_trie_subset marking (SURVEY.md rows
// a12, a13, a19; call sites decoding.rs:185-209, 239-289, 414-424 and
// compact_to_partial_trie.rs:105,125).  The host only shapes the tries — which needs no hash —
// and never computes a Keccak: every node hash, key hash and root comes from the CUDA kernels.
//
// Nibble strings are never materialised: a node refers to a range [nib_start, nib_start+nib_len)
// of a full key stored once in key_pool, with ABSOLUTE positions (a leaf reached at depth d has
// nib_start == d), so splitting and merging paths is index arithmetic.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ppd_status.h"
#include "arena.h"
#include "err_detail.h"

namespace ppd {

struct Fail {
  int code;
  std::string msg;
};
[[noreturn]] inline void fail(int code, const char* msg) { throw Fail{code, msg}; }
[[noreturn]] inline void fail(int code, const std::string& msg) { throw Fail{code, msg}; }

static const uint32_t UNCHANGED = 0xfffffffeu;
// Hashed-out subtrees (Node::Hash) are not arena nodes: their id is HASH_ID_BASE + index into hash_pool,
// their ref IS the pool entry, and nothing has to be computed for them.
static const uint32_t HASH_ID_BASE = 0x80000000u, HASH_ID_END = 0xf0000000u;
static inline bool is_hash_id(uint32_t n) { return n >= HASH_ID_BASE && n < HASH_ID_END; }

// Growable array of trivially copyable elements whose storage survives clear() and may come from a
// caller-supplied allocator (page-locked host memory for everything that is copied to or from the
// device: the context keeps these buffers across calls, so steady state allocates nothing and the
// copies run at full PCIe rate).  resize() does not initialise new elements.
typedef void* (*PvecAlloc)(size_t);
typedef void (*PvecFree)(void*);
template <class T>
struct PVec {
  T* d = nullptr;
  size_t n = 0, cap = 0;
  PvecAlloc alloc_fn = nullptr;
  PvecFree free_fn = nullptr;
  PVec() {}
  PVec(const PVec&) = delete;
  PVec& operator=(const PVec&) = delete;
  ~PVec() { release(); }
  void release() {
    if (d) (free_fn ? free_fn : ::free)(d);
    d = nullptr, n = cap = 0;
  }
  void reserve(size_t want) {
    if (want <= cap) return;
    size_t nc = cap ? cap * 2 : 1024;
    if (nc < want) nc = want;
    T* nd = (T*)(alloc_fn ? alloc_fn : ::malloc)(nc * sizeof(T));
    if (!nd) fail(PPD_ERR_BAD_ARGUMENT, "out of host memory");
    if (n) memcpy(nd, d, n * sizeof(T));
    if (d) (free_fn ? free_fn : ::free)(d);
    d = nd, cap = nc;
  }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T* data() { return d; }
  const T* data() const { return d; }
  T& operator[](size_t i) { return d[i]; }
  const T& operator[](size_t i) const { return d[i]; }
  T& back() { return d[n - 1]; }
  void clear() { n = 0; }
  void resize(size_t m) {
    reserve(m);
    n = m;
  }
  void push_back(const T& v) {
    if (n == cap) reserve(n + 1);
    d[n++] = v;
  }
  void append(const T* p, size_t k) {
    reserve(n + k);
    memcpy(d + n, p, k * sizeof(T));
    n += k;
  }
};

struct HostArena {
  PVec<NodeRec> nodes;
  PVec<uint16_t> level;
  PVec<uint8_t> key_pool, val_pool, hash_pool;
  PVec<uint32_t> child_pool;
  PVec<AccountRec> accounts;
  void set_allocator(PvecAlloc a, PvecFree f) {
    nodes.alloc_fn = key_pool.alloc_fn = val_pool.alloc_fn = hash_pool.alloc_fn = a;
    nodes.free_fn = key_pool.free_fn = val_pool.free_fn = hash_pool.free_fn = f;
    child_pool.alloc_fn = a, child_pool.free_fn = f;
    accounts.alloc_fn = a, accounts.free_fn = f;
  }

  void clear() {
    nodes.clear(), level.clear(), key_pool.clear(), val_pool.clear(), hash_pool.clear(), child_pool.clear(), accounts.clear();
  }

  // ---- pools -------------------------------------------------------------------------------
  uint32_t add_key_nibbles(const uint8_t* nib, uint32_t n) {
    uint32_t off = (uint32_t)key_pool.size();
    key_pool.resize(off + (n + 1) / 2 + 1);  // one slack byte
    memset(key_pool.data() + off, 0, (n + 1) / 2 + 1);
    for (uint32_t i = 0; i < n; i++) key_pool[off + (i >> 1)] |= (i & 1) ? nib[i] : (uint8_t)(nib[i] << 4);
    return off;
  }
  uint32_t add_key_bytes(const uint8_t* bytes, uint32_t nbytes) {
    uint32_t off = (uint32_t)key_pool.size();
    key_pool.append(bytes, nbytes);
    key_pool.push_back(0);
    return off;
  }
  uint32_t add_val(const uint8_t* p, uint32_t n) {
    uint32_t off = (uint32_t)((val_pool.size() + 3) & ~(size_t)3);
    val_pool.resize(off + n);
    if (n) memcpy(val_pool.data() + off, p, n);
    return off;
  }
  uint32_t add_hash(const uint8_t* h) {
    uint32_t idx = (uint32_t)(hash_pool.size() / 32);
    hash_pool.append(h, 32);
    return idx;
  }
  uint32_t key_nib(uint32_t koff, uint32_t i) const {
    uint8_t b = key_pool[koff + (i >> 1)];
    return (i & 1) ? (b & 15u) : (uint32_t)(b >> 4);
  }

  // ---- node accessors ----------------------------------------------------------------------
  uint32_t kind(uint32_t n) const { return is_hash_id(n) ? (uint32_t)NK_HASH : nodes[n].w0 & 0xff; }
  uint32_t nstart(uint32_t n) const { return (nodes[n].w0 >> 8) & 0xff; }
  uint32_t nlen(uint32_t n) const { return (nodes[n].w0 >> 16) & 0xff; }
  bool is_leaf(uint32_t n) const { return kind(n) == NK_LEAF || kind(n) == NK_LEAF_ACCOUNT; }
  bool is_opaque(uint32_t n) const { return kind(n) == NK_HASH || kind(n) == NK_ROOT; }  // Node::Hash
  const uint8_t* hash_of(uint32_t n) const { return hash_pool.data() + 32ull * (n - HASH_ID_BASE); }
  uint32_t child_at(uint32_t br, uint32_t nib) const {
    uint32_t mask = nodes[br].a1, bit = 1u << nib;
    if (!(mask & bit)) return NODE_EMPTY;
    return child_pool[nodes[br].a0 + __builtin_popcount(mask & (bit - 1))];
  }
  uint16_t lvl(uint32_t n) const { return (n == NODE_EMPTY || is_hash_id(n)) ? 0 : level[n]; }

  // ---- constructors ------------------------------------------------------------------------
  uint32_t push(const NodeRec& r, uint32_t lv) {
    if (lv > 0xffff) fail(PPD_ERR_BAD_ARGUMENT, "trie deeper than 65535 levels");
    if (nodes.size() >= HASH_ID_BASE - 1) fail(PPD_ERR_BAD_ARGUMENT, "arena holds at most 2^31 nodes");
    nodes.push_back(r);
    level.push_back((uint16_t)lv);
    return (uint32_t)nodes.size() - 1;
  }
  uint32_t new_hash(uint32_t hash_idx) {
    if (hash_idx >= HASH_ID_END - HASH_ID_BASE) fail(PPD_ERR_BAD_ARGUMENT, "too many hashed-out nodes");
    return HASH_ID_BASE + hash_idx;
  }
  uint32_t new_leaf(uint32_t koff, uint32_t start, uint32_t len, uint32_t val_off, uint32_t val_len) {
    return push({node_w0(NK_LEAF, start, len), koff, val_off, val_len}, 0);
  }
  uint32_t new_account_leaf(uint32_t koff, uint32_t start, uint32_t len, uint32_t rec) {
    uint32_t src = accounts[rec].storage_src;
    return push({node_w0(NK_LEAF_ACCOUNT, start, len), koff, rec, 0}, src == NODE_EMPTY ? 0 : lvl(src) + 1u);
  }
  // the same leaf payload under a different key range
  uint32_t releaf(uint32_t leaf, uint32_t koff, uint32_t start, uint32_t len) {
    NodeRec r = nodes[leaf];
    r.w0 = node_w0(r.w0 & 0xff, start, len);
    r.a0 = koff;
    return push(r, level[leaf]);
  }
  uint32_t new_ext(uint32_t koff, uint32_t start, uint32_t len, uint32_t child) {
    return push({node_w0(NK_EXT, start, len), koff, child, 0}, lvl(child) + 1u);
  }
  uint32_t new_root(uint32_t child) { return push({node_w0(NK_ROOT, 0, 0), 0, child, 0}, child == NODE_EMPTY ? 0 : lvl(child) + 1u); }
  uint32_t new_branch(uint32_t mask, const uint32_t* kids) {
    uint32_t base = (uint32_t)child_pool.size(), lv = 0;
    uint32_t k = (uint32_t)__builtin_popcount(mask);
    for (uint32_t i = 0; i < k; i++) {
      child_pool.push_back(kids[i]);
      if (lvl(kids[i]) > lv) lv = lvl(kids[i]);
    }
    return push({node_w0(NK_BRANCH, 0, 0), base, mask, 0}, lv + 1u);
  }
  // copy of branch `br` with slot `nib` set to `child` (NODE_EMPTY removes it).  The compact child list is
  // copied as a block.  The level is an upper bound (the old branch's level, or the new child's + 1): any
  // level above those of everything a node reads is a valid place in the bottom-up sweep.
  uint32_t branch_with(uint32_t br, uint32_t nib, uint32_t child) {
    const uint32_t mask = nodes[br].a1 & 0xffffu, src = nodes[br].a0, bit = 1u << nib;
    const uint32_t k = (uint32_t)__builtin_popcount(mask), r = (uint32_t)__builtin_popcount(mask & (bit - 1));
    const uint32_t has = (mask & bit) ? 1u : 0u, put = child != NODE_EMPTY ? 1u : 0u;
    const uint32_t nk = k - has + put, nmask = (mask & ~bit) | (put ? bit : 0u);
    const uint32_t base = (uint32_t)child_pool.size();
    child_pool.resize(base + nk);
    uint32_t* d = child_pool.data() + base;
    const uint32_t* o = child_pool.data() + src;
    for (uint32_t i = 0; i < r; i++) d[i] = o[i];
    if (put) d[r] = child;
    for (uint32_t i = r + has; i < k; i++) d[i - has + put] = o[i];
    uint32_t lv = level[br];
    if (put && lvl(child) + 1u > lv) lv = lvl(child) + 1u;
    return push({node_w0(NK_BRANCH, 0, 0), base, nmask, 0}, lv);
  }

  // ---- persistent operations; keys are (koff, klen) full keys in key_pool --------------------
  uint32_t common_prefix(uint32_t koff_a, uint32_t start_a, uint32_t len_a, uint32_t koff_b, uint32_t start_b, uint32_t len_b) const {
    uint32_t m = len_a < len_b ? len_a : len_b, i = 0;
    while (i < m && key_nib(koff_a, start_a + i) == key_nib(koff_b, start_b + i)) i++;
    return i;
  }

  struct Payload {  // what a new leaf holds
    bool account;
    uint32_t a1, a2;  // LEAF: val_off, val_len; LEAF_ACCOUNT: record
  };
  uint32_t make_leaf(uint32_t koff, uint32_t start, uint32_t len, const Payload& p) {
    return p.account ? new_account_leaf(koff, start, len, p.a1) : new_leaf(koff, start, len, p.a1, p.a2);
  }
  uint32_t split(uint32_t koff, uint32_t klen, uint32_t pos, uint32_t cp, uint32_t existing_nib, uint32_t existing, const Payload& p) {
    // branch at depth pos+cp holding `existing` and a new leaf for the key; extension above for the cp common nibbles
    uint32_t at = pos + cp;
    if (at >= klen) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "inserted key is a prefix of an existing key");
    uint32_t new_nib = key_nib(koff, at);
    uint32_t leaf = make_leaf(koff, at + 1, klen - at - 1, p);
    uint32_t kids[2], mask = (1u << existing_nib) | (1u << new_nib);
    if (existing_nib < new_nib)
      kids[0] = existing, kids[1] = leaf;
    else
      kids[0] = leaf, kids[1] = existing;
    uint32_t br = new_branch(mask, kids);
    return cp == 0 ? br : new_ext(koff, pos, cp, br);
  }

  uint32_t insert(uint32_t node, uint32_t koff, uint32_t klen, uint32_t pos, const Payload& p) {
    if (node == NODE_EMPTY) return make_leaf(koff, pos, klen - pos, p);
    switch (kind(node)) {
      case NK_HASH:
      case NK_ROOT:
        fail(PPD_PANIC_INSERT_INTO_HASH_NODE, "insert traversed a hashed-out node");
      case NK_BRANCH: {
        if (pos >= klen) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "inserted key ends at a branch");
        uint32_t nib = key_nib(koff, pos);
        uint32_t nc = insert(child_at(node, nib), koff, klen, pos + 1, p);
        return branch_with(node, nib, nc);
      }
      case NK_EXT: {
        uint32_t ek = nodes[node].a0, es = nstart(node), el = nlen(node), child = nodes[node].a1;
        uint32_t cp = common_prefix(ek, es, el, koff, pos, klen - pos);
        if (cp == el) return new_ext(ek, es, el, insert(child, koff, klen, pos + el, p));
        uint32_t rem = el - cp - 1;
        uint32_t existing = rem == 0 ? child : new_ext(ek, es + cp + 1, rem, child);
        return split(koff, klen, pos, cp, key_nib(ek, es + cp), existing, p);
      }
      default: {  // leaves
        uint32_t lk = nodes[node].a0, ls = nstart(node), ll = nlen(node);
        uint32_t cp = common_prefix(lk, ls, ll, koff, pos, klen - pos);
        if (cp == ll && ll == klen - pos) return make_leaf(koff, pos, klen - pos, p);  // overwrite
        if (cp == ll) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "existing key is a prefix of the inserted key");
        uint32_t existing = releaf(node, lk, ls + cp + 1, ll - cp - 1);
        return split(koff, klen, pos, cp, key_nib(lk, ls + cp), existing, p);
      }
    }
  }

  // Inserts of many distinct keys into one trie version (all the accounts one txn writes): items sorted by key,
  // all of them sharing the first `pos` nibbles.  The result is the trie inserting them one by one gives (a
  // Merkle-Patricia trie is canonical for its key set), but every node on the shared upper part of the paths is
  // created once instead of once per key: the versions in between, which nobody observes (the reference hashes
  // a trie once per txn, decoding.rs:458-464), are neither built nor hashed.
  struct BatchItem {
    uint32_t koff, klen;
    Payload payload;
  };
  uint32_t insert_many(uint32_t node, const BatchItem* it, size_t lo, size_t hi, uint32_t pos) {
    if (lo == hi) return node;
    if (hi - lo == 1) return insert(node, it[lo].koff, it[lo].klen, pos, it[lo].payload);
    const uint32_t k0 = node == NODE_EMPTY ? (uint32_t)NK_LEAF : kind(node);
    if (k0 == NK_HASH || k0 == NK_ROOT) fail(PPD_PANIC_INSERT_INTO_HASH_NODE, "insert traversed a hashed-out node");
    if (k0 == NK_EXT) {
      const uint32_t ek = nodes[node].a0, es = nstart(node), el = nlen(node), child = nodes[node].a1;
      // sorted keys: when the first and the last run through the whole extension, all of them do
      if (common_prefix(ek, es, el, it[lo].koff, pos, it[lo].klen - pos) == el &&
          common_prefix(ek, es, el, it[hi - 1].koff, pos, it[hi - 1].klen - pos) == el) {
        uint32_t nc = insert_many(child, it, lo, hi, pos + el);
        return new_ext(ek, es, el, nc);
      }
    }
    if (k0 != NK_BRANCH) {  // an empty slot, a leaf, an extension that has to split: one by one
      for (size_t i = lo; i < hi; i++) node = insert(node, it[i].koff, it[i].klen, pos, it[i].payload);
      return node;
    }
    uint32_t kids[16];
    {
      const uint32_t mask = nodes[node].a1 & 0xffffu, base = nodes[node].a0;
      uint32_t r = 0;
      for (uint32_t nib = 0; nib < 16; nib++) kids[nib] = (mask >> nib) & 1u ? child_pool[base + r++] : NODE_EMPTY;
    }
    uint32_t lv = level[node];
    if (hi - lo >= 3) {  // the children the groups below descend into: fetched while the first group is being worked on
      for (size_t i = lo; i < hi; i++) {
        if (pos >= it[i].klen) break;
        const uint32_t c = kids[key_nib(it[i].koff, pos)];
        if (c != NODE_EMPTY && !is_hash_id(c)) __builtin_prefetch(&nodes[c]);
      }
    }
    for (size_t i = lo; i < hi;) {
      if (pos >= it[i].klen) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "inserted key ends at a branch");
      const uint32_t nib = key_nib(it[i].koff, pos);
      size_t j = i + 1;
      while (j < hi && pos < it[j].klen && key_nib(it[j].koff, pos) == nib) j++;
      kids[nib] = insert_many(kids[nib], it, i, j, pos + 1);  // (the arena may have been reallocated: nothing of `node` is held across this call)
      if (lvl(kids[nib]) + 1u > lv) lv = lvl(kids[nib]) + 1u;
      i = j;
    }
    const uint32_t base = (uint32_t)child_pool.size();
    uint32_t nmask = 0;
    for (uint32_t nib = 0; nib < 16; nib++)
      if (kids[nib] != NODE_EMPTY) child_pool.push_back(kids[nib]), nmask |= 1u << nib;
    return push({node_w0(NK_BRANCH, 0, 0), base, nmask, 0}, lv);
  }

  // an extension (ek, es, el) over `child`, merged into the child when that is a leaf / extension
  uint32_t collapse_ext(uint32_t ek, uint32_t es, uint32_t el, uint32_t child) {
    switch (kind(child)) {
      case NK_EXT:
        return new_ext(nodes[child].a0, nstart(child) - el, nlen(child) + el, nodes[child].a1);
      case NK_LEAF:
      case NK_LEAF_ACCOUNT:
        return releaf(child, nodes[child].a0, nstart(child) - el, nlen(child) + el);
      default:  // branch, hashed-out node
        return new_ext(ek, es, el, child);
    }
  }
  // returns UNCHANGED when the key is absent
  uint32_t remove(uint32_t node, uint32_t koff, uint32_t klen, uint32_t pos) {
    if (node == NODE_EMPTY) return UNCHANGED;
    switch (kind(node)) {
      case NK_HASH:
      case NK_ROOT:
        return UNCHANGED;
      case NK_EXT: {
        uint32_t ek = nodes[node].a0, es = nstart(node), el = nlen(node);
        if (klen - pos < el || common_prefix(ek, es, el, koff, pos, el) != el) return UNCHANGED;
        uint32_t r = remove(nodes[node].a1, koff, klen, pos + el);
        if (r == UNCHANGED) return UNCHANGED;
        if (r == NODE_EMPTY) return NODE_EMPTY;
        return collapse_ext(ek, es, el, r);  // the extension's own key spells its nibbles
      }
      case NK_BRANCH: {
        if (pos >= klen) return UNCHANGED;
        uint32_t nib = key_nib(koff, pos);
        uint32_t r = remove(child_at(node, nib), koff, klen, pos + 1);
        if (r == UNCHANGED) return UNCHANGED;
        if (r != NODE_EMPTY) return branch_with(node, nib, r);
        uint32_t left = nodes[node].a1 & ~(1u << nib);
        int cnt = __builtin_popcount(left);
        if (cnt >= 2) return branch_with(node, nib, NODE_EMPTY);
        if (cnt == 0) return NODE_EMPTY;
        uint32_t other_nib = (uint32_t)__builtin_ctz(left);
        uint32_t other = child_at(node, other_nib);
        // a key that runs through the surviving child: the removed key's first `pos` nibbles, then its slot
        uint32_t pk = (uint32_t)key_pool.size();
        key_pool.resize(pk + pos / 2 + 2);
        memcpy(key_pool.data() + pk, key_pool.data() + koff, pos / 2 + 1);
        uint8_t* last = key_pool.data() + pk + pos / 2;
        *last = (pos & 1) ? (uint8_t)((*last & 0xf0) | other_nib) : (uint8_t)(other_nib << 4);
        key_pool[pk + pos / 2 + 1] = 0;
        return collapse_ext(pk, pos, 1, other);
      }
      default: {
        uint32_t lk = nodes[node].a0, ls = nstart(node), ll = nlen(node);
        if (ll == klen - pos && common_prefix(lk, ls, ll, koff, pos, ll) == ll) return NODE_EMPTY;
        return UNCHANGED;
      }
    }
  }

  // node holding the value of the key, or NODE_EMPTY
  uint32_t get(uint32_t node, uint32_t koff, uint32_t klen) const {
    uint32_t pos = 0;
    while (node != NODE_EMPTY) {
      switch (kind(node)) {
        case NK_HASH:
        case NK_ROOT:
          return NODE_EMPTY;
        case NK_BRANCH:
          if (pos >= klen) return NODE_EMPTY;
          node = child_at(node, key_nib(koff, pos));
          pos++;
          break;
        case NK_EXT: {
          uint32_t el = nlen(node);
          if (klen - pos < el || common_prefix(nodes[node].a0, nstart(node), el, koff, pos, el) != el) return NODE_EMPTY;
          pos += el;
          node = nodes[node].a1;
          break;
        }
        default: {
          uint32_t ll = nlen(node);
          if (ll == klen - pos && common_prefix(nodes[node].a0, nstart(node), ll, koff, pos, ll) == ll) return node;
          return NODE_EMPTY;
        }
      }
    }
    return NODE_EMPTY;
  }

  // The marking walks of many keys at once (all the keys one txn touches).  The walks are independent and each
  // step is two dependent cache misses (node record, then child slot), so G of them advance in lock step with
  // the next record / slot of every walk prefetched before any of them is read.  Same result per key as mark().
  struct MarkItem {
    uint32_t root, koff, klen;
    uint32_t leaf;  // out: the leaf holding the key, or NODE_EMPTY
  };
  void mark_many(MarkItem* items, size_t n, std::vector<uint32_t>& touched) const {
    constexpr int G = 32;
    for (size_t g = 0; g < n; g += G) {
      const int m = (int)(n - g < (size_t)G ? n - g : G);
      uint32_t node[G], pos[G];
      const uint32_t* slot[G];
      int live = 0;
      for (int j = 0; j < m; j++) {
        node[j] = items[g + j].root, pos[j] = 0, slot[j] = nullptr;
        items[g + j].leaf = NODE_EMPTY;
        if (node[j] != NODE_EMPTY) {
          live++;
          if (!is_hash_id(node[j])) __builtin_prefetch(&nodes[node[j]]);
        }
      }
      while (live) {
        for (int j = 0; j < m; j++) {
          const uint32_t nd = node[j];
          if (nd == NODE_EMPTY) continue;
          MarkItem& it = items[g + j];
          touched.push_back(nd);
          uint32_t next = NODE_EMPTY;
          switch (kind(nd)) {
            case NK_HASH:
            case NK_ROOT:
              if (pos[j] < it.klen) fail(PPD_ERR_MISSING_KEYS_CREATING_SUB_PARTIAL_TRIE, "subset key runs into a hashed-out node");
              break;
            case NK_BRANCH: {
              if (pos[j] >= it.klen) break;
              const uint32_t mask = nodes[nd].a1, bit = 1u << key_nib(it.koff, pos[j]);
              pos[j]++;
              if (mask & bit) {
                slot[j] = &child_pool[nodes[nd].a0 + __builtin_popcount(mask & (bit - 1))];
                __builtin_prefetch(slot[j]);
                continue;  // node[j] is read from the slot in the second half of the step
              }
              break;
            }
            case NK_EXT: {
              uint32_t el = nlen(nd), avail = it.klen - pos[j];
              uint32_t mm = avail < el ? avail : el;
              if (common_prefix(nodes[nd].a0, nstart(nd), mm, it.koff, pos[j], mm) != mm || avail < el) break;
              pos[j] += el;
              next = nodes[nd].a1;
              break;
            }
            default: {
              uint32_t ll = nlen(nd);
              if (ll == it.klen - pos[j] && common_prefix(nodes[nd].a0, nstart(nd), ll, it.koff, pos[j], ll) == ll) it.leaf = nd;
              break;
            }
          }
          node[j] = next;
          if (next == NODE_EMPTY)
            live--;
          else if (!is_hash_id(next))
            __builtin_prefetch(&nodes[next]);
        }
        for (int j = 0; j < m; j++) {
          if (!slot[j]) continue;
          node[j] = *slot[j];
          slot[j] = nullptr;
          if (!is_hash_id(node[j])) __builtin_prefetch(&nodes[node[j]]);  // a set mask bit never holds NODE_EMPTY
        }
      }
    }
  }

  // create_trie_subset's marking pass (trie_subsets.rs mark_nodes_that_are_needed).  Returns what get() returns
  // for the same key: the leaf holding it, or NODE_EMPTY.
  uint32_t mark(uint32_t node, uint32_t koff, uint32_t klen, std::vector<uint32_t>& touched) const {
    uint32_t pos = 0;
    while (node != NODE_EMPTY) {
      touched.push_back(node);
      switch (kind(node)) {
        case NK_HASH:
        case NK_ROOT:
          if (pos < klen) fail(PPD_ERR_MISSING_KEYS_CREATING_SUB_PARTIAL_TRIE, "subset key runs into a hashed-out node");
          return NODE_EMPTY;
        case NK_BRANCH:
          if (pos >= klen) return NODE_EMPTY;
          node = child_at(node, key_nib(koff, pos));
          pos++;
          break;
        case NK_EXT: {
          uint32_t el = nlen(node), avail = klen - pos;
          uint32_t m = avail < el ? avail : el;
          if (common_prefix(nodes[node].a0, nstart(node), m, koff, pos, m) != m) return NODE_EMPTY;
          if (avail < el) return NODE_EMPTY;
          pos += el;
          node = nodes[node].a1;
          break;
        }
        default: {
          uint32_t ll = nlen(node);
          if (ll == klen - pos && common_prefix(nodes[node].a0, nstart(node), ll, koff, pos, ll) == ll) return node;
          return NODE_EMPTY;
        }
      }
    }
    return NODE_EMPTY;
  }
};

}  // namespace ppd
