// host_txn.cu — FlatBlock reader, structural RLP helpers and the host form of the txn loop (decoding.rs:80-177,
// processed_block_trace.rs:209-343) on the persistent arena of host_arena.h.  The GPU runs the txn loop itself
// (gpu_txn.cu); this path takes the blocks it declines (error reporting in the reference's order, witnesses
// built on the host).  Structure only: no hashing.
#include "host_pipeline.h"

namespace ppd {

void read_flat_block(const uint8_t* p, size_t n, BlockJob& b) {
  FlatReader r{p, n};
  if (r.u32() != PPD_FLAT_BLOCK_MAGIC || r.u32() != 1) fail(PPD_ERR_BAD_FLAT_INPUT, "bad magic/version");
  // processed_block_trace.rs:120-168: Combined{compact}, and Separate{Direct state trie, a Direct trie per hashed address}
  // (kind 2, host_direct.cu); every other Separate form is todo!() in the reference
  b.pre_image_kind = r.u32();
  if (b.pre_image_kind != 0 && b.pre_image_kind != 2) fail(PPD_PANIC_UNIMPLEMENTED_PRE_IMAGE, "pre-image variant the reference leaves as todo!()");
  if (b.pre_image_kind == 2)
    b.direct = r.bytes();
  else
    b.compact = r.bytes();
  // counts are checked against the bytes that are left before anything is sized by them (a txn is at least 24 bytes,
  // a trace 21, a resolved-code entry 36, a withdrawal 52)
  uint32_t nt = r.u32();
  if ((uint64_t)nt * 24 > r.n - r.pos) fail(PPD_ERR_BAD_FLAT_INPUT, "txn count exceeds the input");
  b.txns.resize(nt);
  for (uint32_t t = 0; t < nt; t++) {
    TxnV& tx = b.txns[t];
    uint32_t ntr = r.u32();
    if ((uint64_t)ntr * 21 > r.n - r.pos) fail(PPD_ERR_BAD_FLAT_INPUT, "trace count exceeds the input");
    tx.traces.resize(ntr);
    for (uint32_t i = 0; i < ntr; i++) {
      TraceV& tr = tx.traces[i];
      tr.addr = r.raw(20);
      tr.flags = r.u8();
      if (tr.flags & PPD_TR_BALANCE) tr.balance = r.raw(32);
      if (tr.flags & PPD_TR_NONCE) tr.nonce = r.raw(32);
      if (tr.flags & PPD_TR_STORAGE_READ) {
        tr.n_reads = r.u32();
        tr.reads = r.raw(32ull * tr.n_reads);
      }
      if (tr.flags & PPD_TR_STORAGE_WRITTEN) {
        tr.n_writes = r.u32();
        tr.writes = r.raw(64ull * tr.n_writes);
      }
      if (tr.flags & PPD_TR_CODE_READ) tr.code_read = r.raw(32);
      if (tr.flags & PPD_TR_CODE_WRITE) tr.code_write = r.bytes();
    }
    tx.byte_code = r.bytes();
    tx.new_txn_node = r.bytes();
    tx.new_receipt_node = r.bytes();
    tx.gas_used = r.u64();
  }
  uint32_t nc = r.u32();
  if ((uint64_t)nc * 36 > r.n - r.pos) fail(PPD_ERR_BAD_FLAT_INPUT, "code count exceeds the input");
  for (uint32_t i = 0; i < nc; i++) {
    H256 h;
    memcpy(h.b, r.raw(32), 32);
    b.resolved_code[h] = r.bytes();
  }
  uint32_t nw = r.u32();
  if ((uint64_t)nw * 52 > r.n - r.pos) fail(PPD_ERR_BAD_FLAT_INPUT, "withdrawal count exceeds the input");
  for (uint32_t i = 0; i < nw; i++) {
    const uint8_t* a = r.raw(20);
    const uint8_t* v = r.raw(32);
    b.withdrawals.push_back({a, v});
  }
  b.checkpoint = r.raw(32);
  b.b_meta = r.bytes();
  b.b_hashes = r.bytes();
}

// ---- minimal RLP helpers (structure only; no hashing) -----------------------------------------
uint32_t u256_sig(const uint8_t* be) {
  uint32_t i = 0;
  while (i < 32 && be[i] == 0) i++;
  return 32 - i;
}
void rlp_str(std::vector<uint8_t>& out, const uint8_t* p, size_t n) {
  if (n == 1 && p[0] < 0x80) {
    out.push_back(p[0]);
    return;
  }
  if (n < 56) {
    out.push_back((uint8_t)(0x80 + n));
  } else {
    uint8_t tmp[8];
    int k = 0;
    for (size_t v = n; v; v >>= 8) tmp[k++] = (uint8_t)v;
    out.push_back((uint8_t)(0xb7 + k));
    while (k) out.push_back(tmp[--k]);
  }
  out.insert(out.end(), p, p + n);
}
void rlp_u256(std::vector<uint8_t>& out, const uint8_t* be) {
  uint32_t s = u256_sig(be);
  rlp_str(out, be + 32 - s, s);
}
bool rlp_item(const uint8_t* p, size_t n, RlpItem& it) {
  if (n == 0) return false;
  uint8_t b = p[0];
  if (b < 0x80) {
    it = {false, p, 1, 1};
    return true;
  }
  bool is_list = b >= 0xc0;
  uint8_t sb = is_list ? 0xc0 : 0x80, lb = is_list ? 0xf7 : 0xb7;
  size_t hdr, len;
  if (b <= lb) {
    hdr = 1;
    len = b - sb;
    if (!is_list && len == 1) {
      if (n < 2 || p[1] < 0x80) return false;
    }
  } else {
    size_t ll = b - lb;
    if (ll > 8 || n < 1 + ll || p[1] == 0) return false;
    len = 0;
    for (size_t i = 0; i < ll; i++) len = (len << 8) | p[1 + i];
    if (len < 56) return false;
    hdr = 1 + ll;
  }
  if (len > n - hdr) return false;
  it = {is_list, p + hdr, len, hdr + len};
  return true;
}
bool rlp_is_u256(const uint8_t*& q, size_t& m) {
  RlpItem it;
  if (!rlp_item(q, m, it) || it.is_list || it.payload_len > 32) return false;
  if (it.payload_len && it.payload[0] == 0) return false;
  q += it.total_len, m -= it.total_len;
  return true;
}
// plonky2_evm LegacyReceiptRlp {status: bool, cum_gas_used: U256, bloom: Bytes, logs: Vec<LogRlp>}
bool is_legacy_receipt(const uint8_t* p, size_t n) {
  RlpItem top, it;
  if (!rlp_item(p, n, top) || !top.is_list) return false;
  const uint8_t* q = top.payload;
  size_t m = top.payload_len;
  if (!rlp_item(q, m, it) || it.is_list || it.payload_len > 1) return false;
  if (it.payload_len == 1 && (it.payload[0] == 0 || it.payload[0] > 1)) return false;
  q += it.total_len, m -= it.total_len;
  if (!rlp_is_u256(q, m)) return false;
  if (!rlp_item(q, m, it) || it.is_list) return false;
  q += it.total_len, m -= it.total_len;
  if (!rlp_item(q, m, it) || !it.is_list) return false;
  const uint8_t* lq = it.payload;
  size_t lm = it.payload_len;
  while (lm) {
    RlpItem log, x;
    if (!rlp_item(lq, lm, log) || !log.is_list) return false;
    const uint8_t* f = log.payload;
    size_t fm = log.payload_len;
    if (!rlp_item(f, fm, x) || x.is_list || x.payload_len != 20) return false;
    f += x.total_len, fm -= x.total_len;
    if (!rlp_item(f, fm, x) || !x.is_list) return false;
    const uint8_t* tq = x.payload;
    size_t tm = x.payload_len;
    while (tm) {
      RlpItem t;
      if (!rlp_item(tq, tm, t) || t.is_list || t.payload_len != 32) return false;
      tq += t.total_len, tm -= t.total_len;
    }
    f += x.total_len, fm -= x.total_len;
    if (!rlp_item(f, fm, x) || x.is_list) return false;
    lq += log.total_len, lm -= log.total_len;
  }
  return true;
}

// ---- step 1: parse, collect every byte string that must be hashed ----------------------------
void collect_witness_messages(Job& J, BlockJob& b) {
  b.m_inline_code.clear();
  const std::vector<WNode>& ins = b.wit.ins;
  bool any = false;
  for (size_t i = 0; i < ins.size(); i++)
    if (ins[i].op == PPD_OP_CODE) {
      if (!any) b.m_inline_code.assign(ins.size(), ~0u), any = true;
      Span code = b.wit.code(ins[i]);
      b.m_inline_code[i] = J.kh.add(code.p, code.n);
    }
}

void collect_messages(Job& J, BlockJob& b) {
  if (!b.pre_image_on_gpu) {
    parse_witness(b.compact.p, b.compact.n, b.wit);
    if (b.wit.version != 1) fail(PPD_PANIC_INCOMPATIBLE_HEADER_VERSION, "compact header version is not 1");
    collect_witness_messages(J, b);
  }
  for (TxnV& tx : b.txns)
    for (TraceV& tr : tx.traces) {
      tr.m_addr = J.kh.add(tr.addr, 20);
      tr.m_reads = (uint32_t)J.kh.lens.size();
      for (uint32_t k = 0; k < tr.n_reads; k++) J.kh.add(tr.reads + 32 * k, 32);
      tr.m_writes_full = (uint32_t)J.kh.lens.size();
      for (uint32_t k = 0; k < tr.n_writes; k++) J.kh.add(tr.writes + 64 * k, 32);
      // decoding.rs:235 hashes Nibbles::bytes_be() of the raw slot key, which drops leading zero bytes
      tr.m_writes_min = (uint32_t)J.kh.lens.size();
      for (uint32_t k = 0; k < tr.n_writes; k++) {
        const uint8_t* key = tr.writes + 64 * k;
        uint32_t z = 0;
        while (z < 32 && key[z] == 0) z++;
        J.kh.add(key + z, 32 - z);
      }
      if (tr.flags & PPD_TR_CODE_WRITE) tr.m_code = J.kh.add(tr.code_write.p, tr.code_write.n);
    }
  for (auto& w : b.withdrawals) b.m_withdrawal_addr.push_back(J.kh.add(w.first, 20));
}

// ---- step 3: the txn loop (decoding.rs:80-177), shaping only ------------------------------------
uint32_t key_from_digest(Job& J, const H256& h) { return J.A.add_key_bytes(h.b, 32); }

uint32_t txn_index_key(Job& J, size_t idx, uint32_t& klen) {
  // Nibbles::from_bytes_be(rlp::encode(&txn_idx)), decoding.rs:190
  uint8_t be[32];
  memset(be, 0, 32);
  for (int k = 0; k < 8; k++) be[31 - k] = (uint8_t)((uint64_t)idx >> (8 * k));
  std::vector<uint8_t> enc;
  rlp_u256(enc, be);
  klen = (uint32_t)enc.size() * 2;
  return J.A.add_key_bytes(enc.data(), (uint32_t)enc.size());
}

void u256_add(uint8_t a[32], const uint8_t b[32]) {
  unsigned carry = 0;
  for (int i = 31; i >= 0; i--) {
    unsigned s = (unsigned)a[i] + b[i] + carry;
    a[i] = (uint8_t)s;
    carry = s >> 8;
  }
}

void dummy_plan(Job& J, BlockJob& b, IrPlan& p, uint32_t state_root, uint32_t txn_root, uint32_t receipt_root,
                const H256Map& storage, uint64_t txn_number, uint64_t gas_used) {
  // create_dummy_gen_input (decoding.rs:484-549): every trie cut with the key 0_u64, which converts
  // to zero nibbles: the root is the only marked node
  p.txn_before = txn_number;
  p.gas_before = p.gas_after = gas_used;
  p.state_sub = state_root, p.txn_sub = txn_root, p.receipt_sub = receipt_root;
  if (state_root != NODE_EMPTY) p.touched.push_back(state_root);
  if (txn_root != NODE_EMPTY) p.touched.push_back(txn_root);
  if (receipt_root != NODE_EMPTY) p.touched.push_back(receipt_root);
  storage.for_each([&](const H256Map::Entry& s) {
    p.storage_subs.push_back({s.first, s.second});
    if (s.second != NODE_EMPTY) p.touched.push_back(s.second);
  });
  p.root_state = root_node_for(J, b, state_root);
  p.root_txn = root_node_for(J, b, txn_root);
  p.root_receipt = root_node_for(J, b, receipt_root);
}

// AccountDecode's payload (decoding.rs:604-607): the bytes of the leaf that is not an account.  A pre-image built on the
// GPU keeps its value pool on the device until an IR is serialised on the host; a value that is not resident here is
// left out rather than fetched on the error path.
static std::string leaf_value_detail(const Job& J, uint32_t leaf, const char* what) {
  const HostArena& A = J.A;
  const uint32_t off = A.nodes[leaf].a1, n = A.nodes[leaf].a2;
  const bool resident = (J.pools_on_host || off >= J.dev.vals) && (size_t)off + n <= A.val_pool.size();
  return resident ? detail_account_decode(what, A.val_pool.data() + off, n) : std::string(what);
}

void apply_withdrawals(Job& J, BlockJob& b, uint32_t& state_root) {
  HostArena& A = J.A;
  for (size_t i = 0; i < b.withdrawals.size(); i++) {
    const H256& h = J.kh.digest[b.m_withdrawal_addr[i]];
    uint32_t koff = key_from_digest(J, h);
    uint32_t leaf = A.get(state_root, koff, 64);
    if (leaf == NODE_EMPTY) fail(PPD_ERR_MISSING_WITHDRAWAL_ACCOUNT, detail_missing_withdrawal_account(b.withdrawals[i].first, h.b, b.withdrawals[i].second));
    if (A.kind(leaf) != NK_LEAF_ACCOUNT) fail(PPD_ERR_ACCOUNT_DECODE, leaf_value_detail(J, leaf, "withdrawal account does not decode"));
    AccountRec rec = A.accounts[A.nodes[leaf].a1];
    u256_add(rec.balance, b.withdrawals[i].second);
    uint32_t r = (uint32_t)A.accounts.size();
    A.accounts.push_back(rec);
    state_root = A.insert(state_root, koff, 64, 0, HostArena::Payload{true, r, 0});
  }
}

void shape_block(Job& J, BlockJob& b) {
  HostArena& A = J.A;
  SectionTimer sec;
  sec.start();
  if (!b.pre_image_built) build_pre_image(J, b);
  sec.stop(0);
  const uint32_t initial_state = b.state_root;
  // the storage tries before the first txn: only the dummy IRs of a block with at most one txn read them (decoding.rs:304-347)
  H256Map initial_storage;
  if (b.txns.size() <= 1) initial_storage = b.storage;
  uint32_t state = b.state_root, txn_trie = NODE_EMPTY, receipt_trie = NODE_EMPTY;
  uint64_t txn_before = 0, gas_before = 0, gas_after = 0;

  // The reference turns EVERY TxnInfo into its processed form (processed_block_trace.rs:58-66 -> :209-343) before the
  // first txn enters the loop of decoding.rs:106-153: what can fail there -- a code hash nobody resolves, a receipt that
  // is neither legacy nor a byte string -- is reported for the first such txn even when an earlier txn fails in the loop.
  std::vector<H256> written;  // code a trace of the same txn wrote earlier is in the txn's map already (:277-291)
  for (const TxnV& tx : b.txns) {
    written.clear();
    for (const TraceV& tr : tx.traces) {
      if (tr.flags & PPD_TR_CODE_READ) {
        if (memcmp(tr.code_read, EMPTY_CODE_HASH, 32) == 0) continue;
        H256 h;
        memcpy(h.b, tr.code_read, 32);
        if (b.pre_code.find(h) == b.pre_code.end() && b.resolved_code.find(h) == b.resolved_code.end() &&
            std::find(written.begin(), written.end(), h) == written.end())
          fail(PPD_ERR_UNRESOLVED_CODE_HASH, "code hash not resolvable");
      } else if (tr.flags & PPD_TR_CODE_WRITE) {
        written.push_back(J.kh.digest[tr.m_code]);
      }
    }
    RlpItem it;
    const Span receipt = tx.new_receipt_node;
    if (!is_legacy_receipt(receipt.p, receipt.n) && (!rlp_item(receipt.p, receipt.n, it) || it.is_list))
      fail(PPD_PANIC_RECEIPT_DECODE, "receipt is neither legacy nor a byte string");
  }

  for (size_t ti = 0; ti < b.txns.size(); ti++) {
    TxnV& tx = b.txns[ti];
    IrPlan p;
    // ---- into_processed_txn_info (processed_block_trace.rs:209-333): code map, receipt bytes ----
    {
      H256 e;
      memcpy(e.b, EMPTY_CODE_HASH, 32);
      p.code[e] = Span{};
    }
    for (TraceV& tr : tx.traces) {
      if (tr.flags & PPD_TR_CODE_READ) {
        H256 h;
        memcpy(h.b, tr.code_read, 32);
        if (!p.code.count(h)) {
          auto f = b.pre_code.find(h);
          if (f != b.pre_code.end()) {
            p.code[h] = f->second;
          } else {
            auto g = b.resolved_code.find(h);
            if (g == b.resolved_code.end()) fail(PPD_ERR_UNRESOLVED_CODE_HASH, "code hash not resolvable");
            p.code[h] = g->second;
          }
        }
      } else if (tr.flags & PPD_TR_CODE_WRITE) {
        p.code[J.kh.digest[tr.m_code]] = tr.code_write;
      }
    }
    Span receipt = tx.new_receipt_node;
    if (!is_legacy_receipt(receipt.p, receipt.n)) {
      RlpItem it;
      if (!rlp_item(receipt.p, receipt.n, it) || it.is_list) fail(PPD_PANIC_RECEIPT_DECODE, "receipt is neither legacy nor a byte string");
      receipt = Span{it.payload, (uint32_t)it.payload_len};
    }

    sec.stop(1);
    // ---- create_minimal_partial_tries_needed_by_txn (decoding.rs:179-217) ----
    uint32_t tk_len = 0;
    uint32_t tk = txn_index_key(J, ti, tk_len);
    p.state_sub = state, p.txn_sub = txn_trie, p.receipt_sub = receipt_trie;
    // every key this txn touches, marked in one interleaved pass (HostArena::mark_many)
    std::vector<uint32_t>&haddr_key = J.haddr_keys, &haddr_leaf = J.haddr_leaves;
    haddr_key.resize(tx.traces.size()), haddr_leaf.resize(tx.traces.size());
    std::vector<HostArena::MarkItem>& marks = J.mark_items;
    marks.clear();
    for (size_t i = 0; i < tx.traces.size(); i++) {
      haddr_key[i] = key_from_digest(J, J.kh.digest[tx.traces[i].m_addr]);
      marks.push_back({state, haddr_key[i], 64, NODE_EMPTY});
    }
    marks.push_back({txn_trie, tk, tk_len, NODE_EMPTY});
    marks.push_back({receipt_trie, tk, tk_len, NODE_EMPTY});
    p.storage_subs.reserve(tx.traces.size());
    // decoding.rs:199-203 converts EVERY trace's hashed address (H256::from_slice of its bytes_be(): a panic when the hash
    // begins with a zero byte) before the first storage subset is cut, and after the state / txn / receipt subsets: with
    // such an address among the traces no storage key is marked, and the panic is reported after the marks of the three
    // other tries
    bool short_haddr = false;
    for (size_t i = 0; i < tx.traces.size() && !short_haddr; i++) short_haddr = J.kh.digest[tx.traces[i].m_addr].b[0] == 0;
    for (size_t i = 0; i < tx.traces.size() && !short_haddr; i++) {
      TraceV& tr = tx.traces[i];
      const H256& haddr = J.kh.digest[tr.m_addr];
      auto f = b.storage.find(haddr);
      if (f == b.storage.end()) {
        // missing storage trie: Hash(pre-image storage root) when the account had storage in the pre-image
        // and this txn does not access its slots, else an empty trie (decoding.rs:572-582)
        uint32_t t = NODE_EMPTY;
        auto g = b.pre_with_storage.find(haddr);
        if (g != b.pre_with_storage.end() && tr.n_reads + tr.n_writes == 0) t = A.accounts[g->second].storage_src;
        f = b.storage.insert({haddr, t}).first;
      }
      uint32_t sroot = f->second;
      for (uint32_t k = 0; k < tr.n_reads; k++) marks.push_back({sroot, key_from_digest(J, J.kh.digest[tr.m_reads + k]), 64, NODE_EMPTY});
      for (uint32_t k = 0; k < tr.n_writes; k++) marks.push_back({sroot, key_from_digest(J, J.kh.digest[tr.m_writes_full + k]), 64, NODE_EMPTY});
      p.storage_subs.push_back({haddr, sroot});
    }
    p.touched.reserve(marks.size() * 10);
    try {
      A.mark_many(marks.data(), marks.size(), p.touched);
    } catch (const Fail& e) {
      if (e.code != PPD_ERR_MISSING_KEYS_CREATING_SUB_PARTIAL_TRIE) throw;
      // the variant's TrieType is that of the first key, in the reference's order (state accesses, the txn index in the
      // txn and receipt tries, storage slots: decoding.rs:185-209), that runs into a hashed-out node; mark_many walks
      // the keys interleaved, so the keys are walked again one by one on this (cold) path
      std::vector<uint32_t> scratch;
      for (size_t i = 0; i < marks.size(); i++) {
        try {
          A.mark(marks[i].root, marks[i].koff, marks[i].klen, scratch);
        } catch (const Fail&) {
          const size_t nt = tx.traces.size();
          fail(e.code, detail_missing_keys(i < nt ? TRIE_STATE : i == nt ? TRIE_TXN : i == nt + 1 ? TRIE_RECEIPT : TRIE_STORAGE));
        }
      }
      throw;
    }
    // the marking walk of an address also finds its leaf in the pre-txn state: the state writes below read the account
    // from it (other addresses' writes in between only path-copy branches; the leaf's payload stays)
    for (size_t i = 0; i < tx.traces.size() && i < marks.size(); i++) haddr_leaf[i] = marks[i].leaf;
    if (short_haddr) fail(PPD_PANIC_H256_FROM_SLICE, "H256::from_slice on a short bytes_be()");
    gas_after += tx.gas_used;
    sec.stop(2);

    // ---- apply_deltas_to_trie_state (decoding.rs:219-292) ----
    for (TraceV& tr : tx.traces) {
      const H256& haddr = J.kh.digest[tr.m_addr];
      auto f = b.storage.find(haddr);
      if (f == b.storage.end()) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, detail_missing_storage_trie("no storage trie for a written account", haddr.b));
      for (uint32_t k = 0; k < tr.n_writes; k++) {
        const uint8_t* val = tr.writes + 64 * k + 32;
        uint32_t koff = key_from_digest(J, J.kh.digest[tr.m_writes_min + k]);
        uint32_t sig = u256_sig(val);
        if (sig == 0) {  // rlp(0) == [0x80]: a delete
          uint32_t r = A.remove(f->second, koff, 64, 0);
          if (r != UNCHANGED) f->second = r;
        } else {
          uint8_t enc[34];  // rlp(U256): the byte itself below 0x80, else 0x80 + length and the significant bytes
          uint32_t el = 0;
          if (sig == 1 && val[31] < 0x80) {
            enc[el++] = val[31];
          } else {
            enc[el++] = (uint8_t)(0x80 + sig);
            memcpy(enc + el, val + 32 - sig, sig);
            el += sig;
          }
          uint32_t voff = A.add_val(enc, el);
          f->second = A.insert(f->second, koff, 64, 0, HostArena::Payload{false, voff, el});
        }
      }
    }
    sec.stop(3);
    std::vector<HostArena::BatchItem>& batch = J.batch_items;
    batch.clear();
    for (size_t i = 0; i < tx.traces.size(); i++) {
      TraceV& tr = tx.traces[i];
      bool storage_change = tr.n_writes != 0;
      bool code_change = tr.flags & (PPD_TR_CODE_READ | PPD_TR_CODE_WRITE);
      if (!((tr.flags & (PPD_TR_BALANCE | PPD_TR_NONCE)) || storage_change || code_change)) continue;
      const H256& haddr = J.kh.digest[tr.m_addr];
      // the account as state.get() would give it (decoding.rs:251-254): the leaf the marking walk found
      AccountRec rec;
      const uint32_t leaf = haddr_leaf[i];
      if (leaf != NODE_EMPTY) {
        if (A.kind(leaf) != NK_LEAF_ACCOUNT) fail(PPD_ERR_ACCOUNT_DECODE, leaf_value_detail(J, leaf, "state leaf is not an account"));
        rec = A.accounts[A.nodes[leaf].a1];
      } else {
        memset(&rec, 0, sizeof rec);
        memcpy(rec.storage_root, EMPTY_TRIE_HASH, 32);
        memcpy(rec.code_hash, EMPTY_CODE_HASH, 32);
        rec.storage_src = NODE_EMPTY;
      }
      if (storage_change) {
        auto f = b.storage.find(haddr);
        if (f == b.storage.end()) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, detail_missing_storage_trie("no storage trie for a changed account", haddr.b));
        rec.storage_src = root_node_for(J, b, f->second);
      }
      if (tr.flags & PPD_TR_BALANCE) memcpy(rec.balance, tr.balance, 32);
      if (tr.flags & PPD_TR_NONCE) memcpy(rec.nonce, tr.nonce, 32);
      if (tr.flags & PPD_TR_CODE_READ) memcpy(rec.code_hash, tr.code_read, 32);
      if (tr.flags & PPD_TR_CODE_WRITE) memcpy(rec.code_hash, J.kh.digest[tr.m_code].b, 32);
      uint32_t r = (uint32_t)A.accounts.size();
      A.accounts.push_back(rec);
      batch.push_back({haddr_key[i], 64, HostArena::Payload{true, r, 0}});
    }
    // the txn's account writes in one descent (addresses are distinct: TxnInfo.traces is a map, trace_protocol.rs:118)
    std::sort(batch.begin(), batch.end(), [&](const HostArena::BatchItem& x, const HostArena::BatchItem& y) {
      return memcmp(A.key_pool.data() + x.koff, A.key_pool.data() + y.koff, 32) < 0;
    });
    state = A.insert_many(state, batch.data(), 0, batch.size(), 0);
    sec.stop(4);
    for (size_t i = 0; i < tx.traces.size(); i++) {
      if (!(tx.traces[i].flags & PPD_TR_SELF_DESTRUCTED)) continue;
      const H256& haddr = J.kh.digest[tx.traces[i].m_addr];
      if (!b.storage.erase(haddr)) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, detail_missing_storage_trie("self-destructed account has no storage trie", haddr.b));
      uint32_t r = A.remove(state, haddr_key[i], 64, 0);
      if (r != UNCHANGED) state = r;
    }
    {
      uint32_t voff = A.add_val(tx.byte_code.p, tx.byte_code.n);
      txn_trie = A.insert(txn_trie, tk, tk_len, 0, HostArena::Payload{false, voff, tx.byte_code.n});
      uint32_t roff = A.add_val(receipt.p, receipt.n);
      receipt_trie = A.insert(receipt_trie, tk, tk_len, 0, HostArena::Payload{false, roff, receipt.n});
    }
    // ---- calculate_trie_input_hashes + GenerationInputs (decoding.rs:130-145) ----
    p.root_state = root_node_for(J, b, state);
    p.root_txn = root_node_for(J, b, txn_trie);
    p.root_receipt = root_node_for(J, b, receipt_trie);
    p.txn_before = txn_before;
    p.gas_before = gas_before;
    p.gas_after = gas_after;
    p.has_signed_txn = tx.byte_code.n != 0;
    p.signed_txn = tx.byte_code;
    txn_before += 1;
    gas_before = gas_after;
    b.irs.push_back(std::move(p));
    sec.stop(5);
  }
  {
    static const char* const names[] = {"pre-image", "code/receipt", "marks", "storage-wr", "state-wr", "rest"};
    sec.report(names, 6);
  }

  // ---- pad_gen_inputs_with_dummy_inputs_if_needed (decoding.rs:304-347) ----
  bool has_wd = !b.withdrawals.empty(), dummies = false;
  if (b.irs.empty()) {
    for (int k = 0; k < 2; k++) {
      IrPlan d;
      dummy_plan(J, b, d, initial_state, NODE_EMPTY, NODE_EMPTY, initial_storage, txn_before, gas_before);
      b.irs.push_back(std::move(d));
    }
    dummies = true;
  } else if (b.irs.size() == 1) {
    IrPlan d;
    if (!has_wd) {
      dummy_plan(J, b, d, initial_state, NODE_EMPTY, NODE_EMPTY, initial_storage, txn_before, gas_before);
      b.irs.insert(b.irs.begin(), std::move(d));
    } else {
      dummy_plan(J, b, d, state, txn_trie, receipt_trie, b.storage, txn_before, gas_before);
      b.irs.push_back(std::move(d));
    }
    dummies = true;
  }
  // ---- add_withdrawals_to_txns (decoding.rs:356-402) ----
  if (has_wd) {
    if (!dummies) {
      IrPlan d;
      dummy_plan(J, b, d, state, txn_trie, receipt_trie, b.storage, txn_before, gas_before);
      apply_withdrawals(J, b, state);
      d.has_withdrawals = true;
      d.root_state = root_node_for(J, b, state);
      b.irs.push_back(std::move(d));
    } else {
      apply_withdrawals(J, b, state);
      b.irs[1].has_withdrawals = true;
      b.irs[1].root_state = root_node_for(J, b, state);
    }
  }
  b.state_root = state;
}

}  // namespace ppd
