// host_witness.cu — the compact witness on the host: the instruction parser with the reference's error order
// (compact_prestate_processing.rs:683-875, 387-668) and the builder of the pre-image tries
// (compact_to_partial_trie.rs:37-190) for the witnesses the GPU parser declines: malformed ones (the error and
// its place in stream order are reported from here), non-canonical ones, and those too small to pay for three
// device round trips.  Structure only: no hashing.
#include <functional>

#include "host_pipeline.h"

namespace ppd {

void parse_witness(const uint8_t* w, size_t n, Witness& out) {
  if (n == 0) fail(PPD_ERR_MISSING_HEADER, "missing header");
  if (n >= 0xffffffffull) fail(PPD_ERR_BAD_ARGUMENT, "witness larger than 4 GiB");
  WCursor c{w, n};
  out.bytes = w, out.len = n;
  out.version = c.read_byte();
  out.ins.clear();
  out.ins.reserve(n / 30 + 16);
  // pass 1: instruction boundaries (compact_prestate_processing.rs:683-875)
  while (c.pos < c.n) {
    WNode in;
    in.op = c.read_byte();
    in.pos = (uint32_t)c.pos;
    in.flags = 0;
    in.first_child = in.next_sibling = -1;
    in.aux = ~0u;
    switch (in.op) {
      case PPD_OP_LEAF:
        check_key_length(c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR));
        c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR);
        break;
      case PPD_OP_EXTENSION:
        check_key_length(c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR));
        break;
      case PPD_OP_BRANCH:
        in.aux = (uint32_t)c.cbor_uint(0xffffffffull);
        break;
      case PPD_OP_HASH:
        if (c.n - c.pos < 32) fail(PPD_ERR_INVALID_BYTES_FOR_TYPE, "short raw hash");
        c.pos += 32;
        break;
      case PPD_OP_CODE:
        c.cbor_bytes(PPD_ERR_INVALID_BYTES_FOR_TYPE);
        break;
      case PPD_OP_ACCOUNT_LEAF: {
        check_key_length(c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR));
        in.flags = c.read_byte();
        if (in.flags & 4) c.cbor_uint(~0ull);
        if (in.flags & 8) {
          Span bal = c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR);
          if (bal.n > 32) fail(PPD_PANIC_U256_FROM_BIG_ENDIAN, "balance wider than 256 bits (U256::from_big_endian panics)");
        }
        if (in.flags & 1) (void)c.cbor_uint(~0ull);
        break;
      }
      case PPD_OP_EMPTY_ROOT:
        break;
      default:
        fail(PPD_ERR_INVALID_OPERATOR, "invalid opcode");
    }
    out.ins.push_back(in);
  }
  // pass 2: the stack machine (compact_prestate_processing.rs:387-668): instructions arrive in post-order
  std::vector<int32_t> stack;
  stack.reserve(256);
  WNode* ins = out.ins.data();
  for (int32_t i = 0; i < (int32_t)out.ins.size(); i++) {
    WNode& in = ins[i];
    switch (in.op) {
      case PPD_OP_EXTENSION:
        if (stack.empty()) fail(PPD_ERR_INVALID_WITNESS_FORMAT, "extension with no preceding node");
        in.first_child = stack.back();
        stack.pop_back();
        break;
      case PPD_OP_BRANCH: {
        size_t expected = (size_t)__builtin_popcount(in.aux);
        if (stack.size() < expected) fail(PPD_ERR_INCORRECT_NUMBER_OF_NODES_PRECEDING_BRANCH, "branch mask wants more nodes than precede it");
        if (in.aux >> 16) fail(PPD_ERR_MISSING_EXPECTED_NODES_PRECEDING_BRANCH, "branch mask has bits above 15");
        size_t base = stack.size() - expected;
        for (size_t k = 0; k < expected; k++) {  // lowest set bit <-> oldest pushed
          if (k == 0)
            in.first_child = stack[base];
          else
            ins[stack[base + k - 1]].next_sibling = stack[base + k];
        }
        if (expected) ins[stack[base + expected - 1]].next_sibling = -1;
        stack.resize(base);
        break;
      }
      case PPD_OP_ACCOUNT_LEAF:
        if (in.flags & 2) {
          if (stack.empty() || ins[stack.back()].op == PPD_OP_CODE) fail(PPD_ERR_INVALID_WITNESS_FORMAT, "account leaf: no storage node");
          in.first_child = stack.back();
          stack.pop_back();
        }
        if (in.flags & 1) {
          if (stack.empty() || (ins[stack.back()].op != PPD_OP_CODE && ins[stack.back()].op != PPD_OP_HASH))
            fail(PPD_ERR_INVALID_WITNESS_FORMAT, "account leaf: no code node");
          in.aux = (uint32_t)stack.back();
          stack.pop_back();
        }
        break;
      default:
        break;
    }
    stack.push_back(i);
  }
  if (stack.size() > 1) fail(PPD_ERR_NON_SINGLE_ENTRY_AFTER_PROCESSING, "more than one entry left");
  out.root = stack.empty() ? -1 : stack[0];
}

// ============================================================================================
// Items of one trie (what HashedPartialTrie::items() would list) and the canonical build
// ============================================================================================
// Canonical trie over sorted, prefix-free items [lo, hi) whose keys agree on the first `depth`
// nibbles.  Equals what inserting them one by one produces (compact_to_partial_trie.rs:105,125).
uint32_t build_range(HostArena& A, const std::vector<TrieItem>& it, size_t lo, size_t hi, uint32_t depth) {
  if (lo == hi) return NODE_EMPTY;
  if (hi - lo == 1) {
    const TrieItem& x = it[lo];
    if (x.kind == 2) {
      uint32_t h = A.new_hash(x.a1);
      return x.klen == depth ? h : A.new_ext(x.koff, depth, x.klen - depth, h);
    }
    if (x.kind == 1) return A.new_account_leaf(x.koff, depth, x.klen - depth, x.a1);
    return A.new_leaf(x.koff, depth, x.klen - depth, x.a1, x.a2);
  }
  const TrieItem& f = it[lo];
  const TrieItem& l = it[hi - 1];
  uint32_t cp = A.common_prefix(f.koff, depth, f.klen - depth, l.koff, depth, l.klen - depth);
  uint32_t at = depth + cp;
  uint32_t kids[16], mask = 0, k = 0;
  size_t i = lo;
  while (i < hi) {
    if (it[i].klen <= at) fail(PPD_PANIC_KEY_IS_PREFIX_OF_KEY, "a key is a prefix of another key");
    uint32_t nib = A.key_nib(it[i].koff, at);
    size_t j = i + 1;
    while (j < hi && it[j].klen > at && A.key_nib(it[j].koff, at) == nib) j++;
    if (mask & (1u << nib)) fail(PPD_ERR_UNSORTED_KEYS, "trie items are not sorted");
    kids[k++] = build_range(A, it, i, j, at + 1);
    mask |= 1u << nib;
    i = j;
  }
  uint32_t br = A.new_branch(mask, kids);
  return cp == 0 ? br : A.new_ext(f.koff, depth, cp, br);
}

// ---- step 2: pre-image tries -------------------------------------------------------------------
// The reference turns the witness tree into a trie by re-inserting every leaf and hashed-out subtree
// with its full key (compact_to_partial_trie.rs:49-139), so the result is the canonical trie of those
// items whatever shape the witness had.  WitnessTrie does the same in two ways:
//   * convert(): one DFS that maps witness nodes to arena nodes directly.  That is only the canonical
//     trie when every branch keeps at least two non-empty children and every extension has a non-empty
//     key over a branch or a hashed-out node; the DFS checks exactly that (`canonical`).
//   * items + build_range(): the general path, used for a trie whose witness is not canonical.
struct WitnessTrie {
  Job& J;
  BlockJob& b;
  bool is_storage;
  bool canonical = true;
  bool wrong_leaf_kind = false;  // a value leaf in the state trie / an account leaf in a storage trie
  uint8_t path[160];    // nibbles
  uint8_t packed[84];   // the same path packed two nibbles per byte, maintained incrementally
  std::vector<TrieItem>* items = nullptr;
  // resolves an account leaf instruction to its record (state trie only)
  uint32_t (*account_record)(Job&, BlockJob&, int32_t idx, const uint8_t* path, uint32_t klen) = nullptr;

  void set_nibble(uint32_t d, uint32_t nib) {
    path[d] = (uint8_t)nib;
    packed[d >> 1] = (d & 1) ? (uint8_t)((packed[d >> 1] & 0xf0) | nib) : (uint8_t)(nib << 4);
  }
  uint32_t push_key_nibbles(Span k, uint32_t depth) {  // key_bytes_to_nibbles appended at `depth`; returns the new depth
    uint8_t tmp[72];
    uint32_t n = compact_key_nibbles(k, tmp);
    if (depth + n > 64) fail(PPD_ERR_KEY_ERROR, "key longer than 64 nibbles");
    for (uint32_t i = 0; i < n; i++) set_nibble(depth + i, tmp[i]);
    return depth + n;
  }
  uint32_t add_packed_key(uint32_t n) {  // the current path of n nibbles as a key in the pool
    HostArena& A = J.A;
    uint32_t off = (uint32_t)A.key_pool.size(), nb = (n + 1) / 2;
    A.key_pool.resize(off + nb + 1);  // one slack byte (the device may read key[(j >> 1) + 1])
    uint8_t* d = A.key_pool.data() + off;
    memcpy(d, packed, nb);
    if (n & 1) d[nb - 1] &= 0xf0;
    d[nb] = 0;
    return off;
  }
  uint32_t add_leaf_value(Span v) {  // rlp_str(value), compact_to_partial_trie.rs:119
    HostArena& A = J.A;
    uint8_t hdr[9];
    uint32_t hl = 0;
    if (!(v.n == 1 && v.p[0] < 0x80)) {
      if (v.n < 56) {
        hdr[hl++] = (uint8_t)(0x80 + v.n);
      } else {
        uint8_t tmp[8];
        int k = 0;
        for (size_t x = v.n; x; x >>= 8) tmp[k++] = (uint8_t)x;
        hdr[hl++] = (uint8_t)(0xb7 + k);
        while (k) hdr[hl++] = tmp[--k];
      }
    }
    uint32_t off = (uint32_t)((A.val_pool.size() + 3) & ~(size_t)3);
    A.val_pool.resize(off + hl + v.n);
    memcpy(A.val_pool.data() + off, hdr, hl);
    if (v.n) memcpy(A.val_pool.data() + off + hl, v.p, v.n);
    last_val_len = hl + v.n;
    return off;
  }
  uint32_t last_val_len = 0;

  // ---- the direct conversion ----
  uint32_t convert(int32_t idx, uint32_t depth) {
    HostArena& A = J.A;
    const WNode& in = b.wit.ins[idx];
    switch (in.op) {
      case PPD_OP_BRANCH: {
        uint32_t m = in.aux, kids[16], mask = 0, k = 0;
        for (int32_t c = in.first_child; c >= 0; c = b.wit.ins[c].next_sibling) {
          uint32_t nib = (uint32_t)__builtin_ctz(m);
          m &= m - 1;
          if (depth >= 64) fail(PPD_ERR_KEY_ERROR, "key longer than 64 nibbles");
          set_nibble(depth, nib);
          uint32_t r = convert(c, depth + 1);
          if (r != NODE_EMPTY) kids[k++] = r, mask |= 1u << nib;
        }
        if (k == 0) return NODE_EMPTY;
        if (k < 2) canonical = false;
        return A.new_branch(mask, kids);
      }
      case PPD_OP_CODE: {
        if (!is_storage) {  // code found inside a storage subtree is dropped by the reference
          H256 h = J.kh.digest[b.m_inline_code[idx]];
          b.pre_code[h] = b.wit.code(in);
        }
        return NODE_EMPTY;
      }
      case PPD_OP_EMPTY_ROOT:
        return NODE_EMPTY;
      case PPD_OP_HASH:
        if (is_storage) b.storage_partial = true;
        return A.new_hash(A.add_hash(b.wit.hash(in)));
      case PPD_OP_EXTENSION: {
        uint32_t nd = push_key_nibbles(b.wit.key(in), depth);
        uint32_t r = convert(in.first_child, nd);
        if (r == NODE_EMPTY) return NODE_EMPTY;
        uint32_t kd = A.kind(r);
        if (nd == depth || !(kd == NK_BRANCH || kd == NK_HASH)) canonical = false;
        return A.new_ext(add_packed_key(nd), depth, nd - depth, r);
      }
      case PPD_OP_LEAF: {
        uint32_t nd = push_key_nibbles(b.wit.key(in), depth);
        uint32_t koff = add_packed_key(nd);
        uint32_t voff = add_leaf_value(b.wit.leaf_value(in));
        if (!is_storage) wrong_leaf_kind = true;  // reported after the walk, as the general path does
        return A.new_leaf(koff, depth, nd - depth, voff, last_val_len);
      }
      case PPD_OP_ACCOUNT_LEAF: {
        Span key, bal;
        uint64_t nonce;
        b.wit.account(in, key, nonce, bal);
        uint32_t nd = push_key_nibbles(key, depth);
        uint32_t koff = add_packed_key(nd);
        if (is_storage) {
          wrong_leaf_kind = true;
          return A.new_leaf(koff, depth, nd - depth, 0, 0);
        }
        uint32_t rec = account_record(J, b, idx, path, nd);
        return A.new_account_leaf(koff, depth, nd - depth, rec);
      }
    }
    fail(PPD_ERR_INVALID_OPERATOR, "invalid opcode");
  }

  // ---- the general path: compact_to_partial_trie.rs:49-139 as a DFS with an accumulated key ----
  void walk(int32_t idx, uint32_t depth) {
    const WNode& in = b.wit.ins[idx];
    switch (in.op) {
      case PPD_OP_BRANCH: {
        uint32_t m = in.aux;
        for (int32_t c = in.first_child; c >= 0; c = b.wit.ins[c].next_sibling) {
          uint32_t nib = (uint32_t)__builtin_ctz(m);
          m &= m - 1;
          if (depth >= 64) fail(PPD_ERR_KEY_ERROR, "key longer than 64 nibbles");
          set_nibble(depth, nib);
          walk(c, depth + 1);
        }
        return;
      }
      case PPD_OP_CODE: {
        if (!is_storage) {
          H256 h = J.kh.digest[b.m_inline_code[idx]];
          b.pre_code[h] = b.wit.code(in);
        }
        return;
      }
      case PPD_OP_EMPTY_ROOT:
        return;
      case PPD_OP_HASH: {
        if (is_storage) b.storage_partial = true;
        uint32_t koff = add_packed_key(depth);
        items->push_back({koff, depth, 2, J.A.add_hash(b.wit.hash(in)), 0});
        return;
      }
      case PPD_OP_EXTENSION: {
        uint32_t nd = push_key_nibbles(b.wit.key(in), depth);
        walk(in.first_child, nd);
        return;
      }
      case PPD_OP_LEAF: {
        uint32_t nd = push_key_nibbles(b.wit.key(in), depth);
        uint32_t koff = add_packed_key(nd);
        uint32_t voff = add_leaf_value(b.wit.leaf_value(in));
        items->push_back({koff, nd, 0, voff, last_val_len});
        return;
      }
      case PPD_OP_ACCOUNT_LEAF: {
        Span key, bal;
        uint64_t nonce;
        b.wit.account(in, key, nonce, bal);
        uint32_t nd = push_key_nibbles(key, depth);
        uint32_t koff = add_packed_key(nd);
        items->push_back({koff, nd, 1, (uint32_t)idx /* resolved to a record later */, 0});
        return;
      }
    }
  }
};

struct ArenaMark {
  size_t nodes, keys, vals, hashes, children, accounts;
  static ArenaMark take(const HostArena& A) {
    return {A.nodes.size(), A.key_pool.size(), A.val_pool.size(), A.hash_pool.size(), A.child_pool.size(), A.accounts.size()};
  }
  void rewind(HostArena& A) const {
    A.nodes.resize(nodes), A.level.resize(nodes), A.key_pool.resize(keys), A.val_pool.resize(vals), A.hash_pool.resize(hashes);
    A.child_pool.resize(children), A.accounts.resize(accounts);
  }
};

uint32_t root_node_for(Job& J, BlockJob& b, uint32_t trie_root) {
  if (const uint32_t* f = b.root_of.find(trie_root)) return *f;
  uint32_t r = J.A.new_root(trie_root);
  b.root_of.put(trie_root, r);
  return r;
}

bool trie_root_is_empty_hash(const Job& J, uint32_t root) {
  if (root == NODE_EMPTY) return true;
  if (is_hash_id(root)) return memcmp(J.A.hash_of(root), EMPTY_TRIE_HASH, 32) == 0;
  return false;
}

// The account record of an account leaf instruction and the block's per-account tables
// (compact_to_partial_trie.rs:141-190).  `path` holds the klen nibbles of the leaf's full key.
uint32_t make_account_record(Job& J, BlockJob& b, int32_t idx, const uint8_t* path, uint32_t klen) {
  HostArena& A = J.A;
  const Witness& W = b.wit;
  const WNode& in = W.ins[idx];
  Span key, balance;
  uint64_t nonce;
  W.account(in, key, nonce, balance);
  AccountRec rec;
  memset(&rec, 0, sizeof rec);
  for (int k = 0; k < 8; k++) rec.nonce[31 - k] = (uint8_t)(nonce >> (8 * k));
  if (balance.n) memcpy(rec.balance + 32 - balance.n, balance.p, balance.n);
  memcpy(rec.storage_root, EMPTY_TRIE_HASH, 32);
  rec.storage_src = NODE_EMPTY;
  uint32_t sroot = NODE_EMPTY;
  bool has_trie = false, nonempty = false;
  if (in.flags & 2) {
    sroot = b.storage_root_of_instr[idx];
    nonempty = !trie_root_is_empty_hash(J, sroot);
    has_trie = true;
    if (nonempty) rec.storage_src = root_node_for(J, b, sroot);
  }
  // the reference joins accounts to storage tries by ROOT HASH (compact_to_partial_trie.rs:167-190):
  // every account whose root is EMPTY_TRIE_HASH gets the last witnessed empty-rooted trie, if any
  if (!nonempty) {
    has_trie = b.have_empty_form;
    sroot = b.empty_form;
  }
  if (in.flags & 1) {
    const WNode& c = W.ins[in.aux];
    if (c.op == PPD_OP_CODE) {
      H256 h = J.kh.digest[b.m_inline_code[in.aux]];
      memcpy(rec.code_hash, h.b, 32);
      b.pre_code[h] = W.code(c);
    } else {
      memcpy(rec.code_hash, W.hash(c), 32);
    }
  } else {
    memcpy(rec.code_hash, EMPTY_CODE_HASH, 32);
  }
  uint32_t r = (uint32_t)A.accounts.size();
  A.accounts.push_back(rec);
  // hashed address = the leaf's full key, left-padded (utils.rs:49-59)
  // (value-minimal bytes_be, then left-padded to 32 bytes == the nibbles right-aligned)
  H256 haddr;
  memset(haddr.b, 0, 32);
  for (uint32_t k = 0; k < klen; k++) {
    uint32_t posn = 64 - klen + k;
    haddr.b[posn >> 1] |= (uint8_t)((posn & 1) ? path[k] : (path[k] << 4));
  }
  if (has_trie) b.storage[haddr] = sroot;
  b.pre_accounts.push_back({haddr, r, nonempty, (in.flags & 2) != 0, (in.flags & 2) ? b.storage_root_of_instr[idx] : NODE_EMPTY});
  if (nonempty) b.pre_with_storage[haddr] = r;
  return r;
}

// one trie of the pre-image: the direct conversion, or the general path when the witness is not canonical
uint32_t build_witness_trie(Job& J, BlockJob& b, int32_t root_idx, bool is_storage) {
  HostArena& A = J.A;
  {
    ArenaMark mark = ArenaMark::take(A);
    size_t n_pre_accounts = b.pre_accounts.size();
    WitnessTrie wt{J, b, is_storage};
    wt.account_record = make_account_record;
    uint32_t root = wt.convert(root_idx, 0);
    if (wt.wrong_leaf_kind) {
      if (is_storage) fail(PPD_ERR_INVALID_WITNESS_FORMAT, "account leaf inside a storage trie");
      fail(PPD_PANIC_PRE_IMAGE_ACCOUNT_DECODE, "state leaf is not an account");
    }
    if (wt.canonical) return root;
    // not canonical: undo and rebuild from the items
    mark.rewind(A);
    for (size_t i = n_pre_accounts; i < b.pre_accounts.size(); i++) {
      b.storage.erase(b.pre_accounts[i].haddr);
      b.pre_with_storage.erase(b.pre_accounts[i].haddr);
    }
    b.pre_accounts.resize(n_pre_accounts);
    b.root_of.erase_if([&](uint32_t root, uint32_t root_node) {
      return root_node >= mark.nodes || (is_hash_id(root) ? root - HASH_ID_BASE >= mark.hashes / 32 : (root != NODE_EMPTY && root >= mark.nodes));
    });
  }
  std::vector<TrieItem> items;
  WitnessTrie wt{J, b, is_storage};
  wt.items = &items;
  wt.walk(root_idx, 0);
  for (TrieItem& x : items) {
    if (is_storage && x.kind == 1) fail(PPD_ERR_INVALID_WITNESS_FORMAT, "account leaf inside a storage trie");
    if (!is_storage && x.kind == 0) fail(PPD_PANIC_PRE_IMAGE_ACCOUNT_DECODE, "state leaf is not an account");
    if (x.kind != 1) continue;
    uint8_t nib[64];
    for (uint32_t k = 0; k < x.klen; k++) nib[k] = (uint8_t)A.key_nib(x.koff, k);
    x.a1 = make_account_record(J, b, (int32_t)x.a1, nib, x.klen);
  }
  return build_range(A, items, 0, items.size(), 0);
}

void build_pre_image(Job& J, BlockJob& b) {
  const Witness& W = b.wit;
  if (W.root < 0) return;
  // storage tries, in stream order (compact_prestate_processing.rs:608-625)
  b.storage_root_of_instr.clear();
  {
    size_t n_acct = 0, n_storage = 0;
    for (const WNode& x : W.ins) n_acct += x.op == PPD_OP_ACCOUNT_LEAF, n_storage += (x.op == PPD_OP_ACCOUNT_LEAF && (x.flags & 2));
    b.storage.reserve(n_acct);
    b.pre_accounts.reserve(n_acct);
    b.pre_with_storage.reserve(n_storage);
    b.storage_root_of_instr.reserve(n_storage);
    b.root_of.reserve(2 * n_storage + 1024);
  }
  b.have_empty_form = false;
  b.empty_form = NODE_EMPTY;
  for (int32_t i = 0; i < (int32_t)W.ins.size(); i++) {
    const WNode& in = W.ins[i];
    if (in.op != PPD_OP_ACCOUNT_LEAF || !(in.flags & 2)) continue;
    uint32_t root = build_witness_trie(J, b, in.first_child, true);
    b.storage_root_of_instr[i] = root;
    if (trie_root_is_empty_hash(J, root)) b.have_empty_form = true, b.empty_form = root;
  }
  b.state_root = build_witness_trie(J, b, W.root, false);
  b.pre_image_built = true;
}

// convert_storage_trie_root_keyed_hashmap_to_account_addr_keyed (compact_to_partial_trie.rs:167-190) with the roots in
// hand: every storage trie was hashed and stored under its ROOT HASH while the witness was processed (a later trie
// with the same root replacing an earlier one, compact_prestate_processing.rs:617-619), then every account takes the
// trie stored under its storage root.  The builders above give an account its own witnessed trie, which is the same
// thing unless two accounts with one root are witnessed differently (expanded / hashed-out, or expanded in different
// places); this pass hashes the pre-image on the GPU and re-joins by root.
void join_storage_by_root(Lane* L, Job& J, BlockJob& b) {
#ifndef PPD_HOSTPROF
  bool any = false;
  for (const BlockJob::PreAccount& pa : b.pre_accounts) any |= pa.witnesses_storage;
  if (!any) return;
  const ppd_stats keep = L->stats;
  sweep(L, J, /*refs_to_host=*/true);
  L->stats = keep;  // (the block's own sweep counts these nodes)
  auto root_hash = [&](const BlockJob::PreAccount& pa) {
    H256 h;
    const uint32_t src = J.A.accounts[pa.rec].storage_src;
    memcpy(h.b, src == NODE_EMPTY ? EMPTY_TRIE_HASH : J.ref.data() + 32ull * src, 32);
    return h;
  };
  std::unordered_map<H256, size_t, H256Hasher> last;  // root hash -> the last account that witnesses a trie with it
  for (size_t i = 0; i < b.pre_accounts.size(); i++)
    if (b.pre_accounts[i].witnesses_storage) last[root_hash(b.pre_accounts[i])] = i;
  // "Possibility of identical tries between accounts, so we need to do a clone here" (compact_to_partial_trie.rs:183-185):
  // an account that takes ANOTHER account's trie gets its own copy of the nodes, so that a txn touching both cuts a
  // separate subset out of each (the marks of an IR are kept per node id)
  HostArena& A = J.A;
  std::function<uint32_t(uint32_t)> clone = [&](uint32_t n) -> uint32_t {
    if (n == NODE_EMPTY || is_hash_id(n)) return n;
    const NodeRec r = A.nodes[n];
    const uint32_t lv = A.level[n];
    switch (r.w0 & 0xff) {
      case NK_EXT: {
        NodeRec c = r;
        c.a1 = clone(r.a1);
        return A.push(c, lv);
      }
      case NK_BRANCH: {
        const uint32_t k = (uint32_t)__builtin_popcount(r.a1 & 0xffffu);
        uint32_t kids[16];
        for (uint32_t i = 0; i < k; i++) kids[i] = clone(A.child_pool[r.a0 + i]);
        NodeRec c = r;
        c.a0 = (uint32_t)A.child_pool.size();
        for (uint32_t i = 0; i < k; i++) A.child_pool.push_back(kids[i]);
        return A.push(c, lv);
      }
      default:
        return A.push(r, lv);  // leaves (their key and value bytes are immutable and shared); ROOT nodes do not occur inside a trie
    }
  };
  bool cloned = false;
  for (size_t i = 0; i < b.pre_accounts.size(); i++) {
    const BlockJob::PreAccount& pa = b.pre_accounts[i];
    auto f = last.find(root_hash(pa));
    if (f == last.end()) {
      b.storage.erase(pa.haddr);
      continue;
    }
    const uint32_t root = b.pre_accounts[f->second].own_root;
    const bool other = f->second != i && root != NODE_EMPTY && !is_hash_id(root);
    b.storage[pa.haddr] = other ? clone(root) : root;
    cloned |= other;
  }
  (void)cloned;
  J.refs_on_host = false;
#else
  (void)L, (void)J, (void)b;
#endif
}

}  // namespace ppd
