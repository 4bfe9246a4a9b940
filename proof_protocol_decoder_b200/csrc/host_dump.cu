// host_dump.cu — IrDump serialisation on the host threads (small blocks, and the IRs the device flags).
#include "host_pipeline.h"

namespace ppd {

void account_rlp(const Job& J, const AccountRec& rec, Out& o) {
  // rlp([nonce, balance, storage_root, code_hash]) preceded by its length (u32)
  uint32_t nn = u256_sig(rec.nonce), nb = u256_sig(rec.balance);
  auto str_size = [](const uint8_t* be, uint32_t sig) -> uint32_t { return sig == 0 ? 1 : (sig == 1 && be[31] < 0x80) ? 1 : 1 + sig; };
  uint32_t payload = str_size(rec.nonce, nn) + str_size(rec.balance, nb) + 66;
  o.u32(2 + payload);
  o.need(2 + payload);
  uint8_t* q = o.p + o.n;
  *q++ = 0xf8;
  *q++ = (uint8_t)payload;
  auto put_u256 = [&](const uint8_t* be, uint32_t sig) {
    if (sig == 0) {
      *q++ = 0x80;
      return;
    }
    if (!(sig == 1 && be[31] < 0x80)) *q++ = (uint8_t)(0x80 + sig);
    memcpy(q, be + 32 - sig, sig);
    q += sig;
  };
  put_u256(rec.nonce, nn);
  put_u256(rec.balance, nb);
  const uint8_t* sr = rec.storage_src == NODE_EMPTY ? rec.storage_root : J.ref.data() + 32ull * rec.storage_src;
  *q++ = 0xa0;
  memcpy(q, sr, 32);
  q += 32;
  *q++ = 0xa0;
  memcpy(q, rec.code_hash, 32);
  q += 32;
  o.n += 2 + payload;
}

void dump_nibbles(const Job& J, Out& o, uint32_t node) {
  uint32_t k = J.A.nodes[node].a0, s = J.A.nstart(node), n = J.A.nlen(node);
  o.need(1 + n);
  uint8_t* q = o.p + o.n;
  *q++ = (uint8_t)n;
  for (uint32_t i = 0; i < n; i++) *q++ = (uint8_t)J.A.key_nib(k, s + i);
  o.n += 1 + n;
}

// create_partial_trie_subset_from_tracked_trie (trie_subsets.rs): untouched nodes whose encoding is
// at least 32 bytes become Hash nodes; smaller ones are kept as they are
void dump_subset(const Job& J, const Stamp& st, Out& o, uint32_t node) {
  const HostArena& A = J.A;
  if (node == NODE_EMPTY) {
    o.u8(PPD_NODE_EMPTY);
    return;
  }
  if (is_hash_id(node)) {
    o.need(33);
    o.p[o.n] = PPD_NODE_HASH;
    memcpy(o.p + o.n + 1, A.hash_of(node), 32);
    o.n += 33;
    return;
  }
  bool touched = st.v[node] == st.serial;
  if ((!touched && J.ref_len[node] == 32) || A.is_opaque(node)) {
    o.need(33);
    o.p[o.n] = PPD_NODE_HASH;
    memcpy(o.p + o.n + 1, J.ref.data() + 32ull * node, 32);
    o.n += 33;
    return;
  }
  switch (A.kind(node)) {
    case NK_LEAF:
      o.u8(PPD_NODE_LEAF);
      dump_nibbles(J, o, node);
      o.u32(A.nodes[node].a2);
      o.raw(A.val_pool.data() + A.nodes[node].a1, A.nodes[node].a2);
      return;
    case NK_LEAF_ACCOUNT:
      o.u8(PPD_NODE_LEAF);
      dump_nibbles(J, o, node);
      account_rlp(J, A.accounts[A.nodes[node].a1], o);
      return;
    case NK_EXT:
      o.u8(PPD_NODE_EXTENSION);
      dump_nibbles(J, o, node);
      dump_subset(J, st, o, A.nodes[node].a1);
      return;
    case NK_BRANCH: {
      o.u8(PPD_NODE_BRANCH);
      // the children's refs and records are scattered: start all the misses before the first use
      const uint32_t mask = A.nodes[node].a1 & 0xffff, k = (uint32_t)__builtin_popcount(mask);
      const uint32_t* ch = A.child_pool.data() + A.nodes[node].a0;
      for (uint32_t j = 0; j < k; j++) {
        uint32_t c = ch[j];
        if (is_hash_id(c)) {
          __builtin_prefetch(A.hash_of(c));
        } else {
          __builtin_prefetch(&st.v[c]);
          __builtin_prefetch(&A.nodes[c]);
          __builtin_prefetch(J.ref.data() + 32ull * c);
        }
      }
      // hashed-out and untouched children are written inline (33 bytes each); only expanded children recurse
      o.need(16 * 33 + 8);
      uint8_t* q = o.p + o.n;
      for (uint32_t i = 0, j = 0; i < 16; i++) {
        if (!(mask & (1u << i))) {
          *q++ = PPD_NODE_EMPTY;
          continue;
        }
        const uint32_t c = ch[j++];
        const uint8_t* h = nullptr;
        if (is_hash_id(c))
          h = A.hash_of(c);
        else if ((st.v[c] != st.serial && J.ref_len[c] == 32) || A.is_opaque(c))
          h = J.ref.data() + 32ull * c;
        if (h) {
          *q++ = PPD_NODE_HASH;
          memcpy(q, h, 32);
          q += 32;
        } else {
          o.n = (size_t)(q - o.p);
          dump_subset(J, st, o, c);
          o.need((16 - i) * 33 + 8);
          q = o.p + o.n;
        }
      }
      o.n = (size_t)(q - o.p);
      o.u32(0);
      return;
    }
  }
}

void dump_ir(const Job& J, const BlockJob& b, IrPlan& p, Stamp& st, Out& o) {
  st.serial++;
  for (uint32_t t : p.touched)
    if (!is_hash_id(t)) st.v[t] = st.serial;
  o.u256(p.txn_before);
  o.u256(p.gas_before);
  o.u256(p.gas_after);
  o.u8(p.has_signed_txn);
  o.span(p.has_signed_txn ? p.signed_txn : Span{});
  if (p.has_withdrawals) {
    o.u32((uint32_t)b.withdrawals.size());
    for (auto& w : b.withdrawals) {
      o.raw(w.first, 20);
      o.raw(w.second, 32);
    }
  } else {
    o.u32(0);
  }
  dump_subset(J, st, o, p.state_sub);
  dump_subset(J, st, o, p.txn_sub);
  dump_subset(J, st, o, p.receipt_sub);
  std::stable_sort(p.storage_subs.begin(), p.storage_subs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
  o.u32((uint32_t)p.storage_subs.size());
  for (auto& s : p.storage_subs) {
    o.raw(s.first.b, 32);
    dump_subset(J, st, o, s.second);
  }
  o.raw(J.ref.data() + 32ull * p.root_state, 32);
  o.raw(J.ref.data() + 32ull * p.root_txn, 32);
  o.raw(J.ref.data() + 32ull * p.root_receipt, 32);
  o.raw(b.checkpoint, 32);
  o.u32((uint32_t)p.code.size());
  for (auto& cd : p.code) {
    o.raw(cd.first.b, 32);
    o.span(cd.second);
  }
  o.span(b.b_meta);
  o.span(b.b_hashes);
}

unsigned host_threads() {
  static unsigned n = [] {
    if (const char* e = getenv("PPD_HOST_THREADS")) {
      int v = atoi(e);
      if (v >= 1) return (unsigned)std::min(v, 64);
    }
    unsigned h = std::thread::hardware_concurrency();
    return h == 0 ? 1u : std::min(h, 16u);
  }();
  return n;
}

// Every IR of every block of the job: IRs are serialised independently on the host threads (each
// with its own marks), then copied to their place in the block's output buffer.
void dump_blocks(Job& J, uint8_t** outs, size_t* out_lens, unsigned max_workers) {
  struct Item {
    uint32_t block, ir;
  };
  std::vector<Item> items;
  for (size_t i = 0; i < J.blocks.size(); i++) {
    outs[i] = nullptr, out_lens[i] = 0;
    if (J.blocks[i].status != PPD_OK) continue;
    for (size_t k = 0; k < J.blocks[i].irs.size(); k++) items.push_back({(uint32_t)i, (uint32_t)k});
  }
  const unsigned workers = std::max(1u, std::min<unsigned>(max_workers, (unsigned)items.size()));
  std::vector<Stamp> stamps(workers);
  if (workers == 1) {
    // one thread (a lane of a batch): every IR of a block straight into the block's output buffer
    Stamp& st = stamps[0];
    st.v.assign(J.A.nodes.size(), 0);
    for (size_t i = 0; i < J.blocks.size(); i++) {
      BlockJob& b = J.blocks[i];
      if (b.status != PPD_OK) continue;
      size_t touched = 0;
      for (IrPlan& p : b.irs) touched += p.touched.size();
      Out o;
      o.need(4096 + 600 * touched);
      o.u32(PPD_IR_DUMP_MAGIC);
      o.u32((uint32_t)b.irs.size());
      for (IrPlan& p : b.irs) dump_ir(J, b, p, st, o);
      outs[i] = o.give(&out_lens[i]);
    }
    return;
  }
  std::vector<Out> parts(items.size());
  const size_t n_nodes = J.A.nodes.size();
  parallel_for(items.size(), workers, [&](size_t i, unsigned w) {
    Stamp& st = stamps[w];
    if (st.v.size() != n_nodes) st.v.assign(n_nodes, 0), st.serial = 0;  // first item of this worker
    BlockJob& b = J.blocks[items[i].block];
    parts[i].need(256 << 10);
    dump_ir(J, b, b.irs[items[i].ir], st, parts[i]);
  });
  if (getenv("PPD_TIMING")) fprintf(stderr, "[ppd]   dump: serialise done\n");
  // offsets, then parallel copy
  std::vector<size_t> at(items.size());
  for (size_t i = 0, k = 0; i < J.blocks.size(); i++) {
    if (J.blocks[i].status != PPD_OK) continue;
    size_t total = 8;
    for (size_t q = 0; q < J.blocks[i].irs.size(); q++, k++) {
      at[k] = total;
      total += parts[k].n;
    }
    uint8_t* buf = (uint8_t*)malloc(total);
    if (!buf) fail(PPD_ERR_BAD_ARGUMENT, "out of host memory");
    uint32_t hdr[2] = {PPD_IR_DUMP_MAGIC, (uint32_t)J.blocks[i].irs.size()};
    memcpy(buf, hdr, 8);
    outs[i] = buf, out_lens[i] = total;
  }
  parallel_for(items.size(), workers, [&](size_t i, unsigned) { memcpy(outs[items[i].block] + at[i], parts[i].p, parts[i].n); });
}

}  // namespace ppd
