// txn_tables.cu — FlatBlock -> the tables of the device txn loop (see txn_tables.h).  Host only; structure only.
#include "host_pipeline.h"
#include "txn_tables.h"

namespace ppd {

using txn::TxnDesc;
using txn::TxnTrace;

bool txn_tables_phase1(const BlockJob& b, const uint8_t* flat, size_t flat_len, TxnTables& T) {
  T.clear();
  if (flat_len >= 0xfff00000ull) return false;
  size_t n_traces = 0;
  for (const TxnV& tx : b.txns) n_traces += tx.traces.size();
  if (n_traces >= (1u << 24)) return false;
  T.traces.resize(n_traces);
  uint32_t m = (uint32_t)n_traces, op = 0, rec = 0, val = 0, t = 0;
  uint64_t n_items = 0;
  for (size_t ti = 0; ti < b.txns.size(); ti++) {
    const TxnV& tx = b.txns[ti];
    uint32_t item = 0, ops2 = 0;
    const uint32_t op_begin = op;
    for (const TraceV& tr : tx.traces) {
      TxnTrace d;
      memset(&d, 0, sizeof d);
      d.flags = tr.flags;
      d.txn = (uint32_t)ti;
      d.off_addr = (uint32_t)(tr.addr - flat);
      const bool code_change = tr.flags & (PPD_TR_CODE_READ | PPD_TR_CODE_WRITE);
      if ((tr.flags & (PPD_TR_BALANCE | PPD_TR_NONCE)) || tr.n_writes || code_change) d.flags |= txn::TRF_STATE_WRITE;
      if (tr.flags & PPD_TR_BALANCE) d.off_balance = (uint32_t)(tr.balance - flat);
      if (tr.flags & PPD_TR_NONCE) d.off_nonce = (uint32_t)(tr.nonce - flat);
      d.n_reads = tr.n_reads, d.n_writes = tr.n_writes;
      if (tr.n_reads) d.off_reads = (uint32_t)(tr.reads - flat);
      if (tr.n_writes) d.off_writes = (uint32_t)(tr.writes - flat);
      d.m_reads = m, m += tr.n_reads;
      d.m_wfull = m, m += tr.n_writes;
      d.m_wmin = d.m_wfull;
      for (uint32_t k = 0; k < tr.n_writes; k++)
        if (tr.writes[64ull * k] == 0) d.flags |= txn::TRF_MIN_KEYS;
      if (d.flags & txn::TRF_MIN_KEYS) d.m_wmin = m, m += tr.n_writes;
      if (tr.flags & PPD_TR_CODE_READ) {
        d.code_off = (uint32_t)(tr.code_read - flat);
      } else if (tr.flags & PPD_TR_CODE_WRITE) {
        d.code_off = (uint32_t)(tr.code_write.p - flat), d.code_len = tr.code_write.n;
        d.m_code = m++;
        T.code_write_traces.push_back(t);
      }
      T.max_trace_keys = std::max(T.max_trace_keys, tr.n_reads + 2 * tr.n_writes);
      d.n_keys = tr.n_reads + tr.n_writes * ((d.flags & txn::TRF_MIN_KEYS) ? 2u : 1u);
      d.op0 = op, op += d.n_keys;
      item += d.n_keys;
      if (d.flags & txn::TRF_STATE_WRITE) d.rec = rec++;
      d.val0 = val, val += 36u * tr.n_writes;
      ops2++;  // every trace accesses its account
      T.traces[t++] = d;
    }
    op += 2;  // the inserts into the transactions and receipts tries
    T.n_ops2 += ops2;
    T.max_ops = std::max(T.max_ops, op - op_begin + ops2);  // keys of the txn: storage, txn / receipt index, state
    T.max_traces = std::max<uint32_t>(T.max_traces, (uint32_t)tx.traces.size());
    n_items += tx.traces.size() + 2 + item;
  }
  if (n_items * txn::MARK_SLOTS_T >= (1ull << 31)) return false;
  // withdrawals (decoding.rs:404-428): the hashed address of each, and the record the updated account fills
  T.withdrawals.resize(b.withdrawals.size());
  for (size_t w = 0; w < b.withdrawals.size(); w++) T.withdrawals[w] = txn::Withdrawal{m++, (uint32_t)(b.withdrawals[w].second - flat), rec++, 0};
  // the entries of the IrDump (pad_gen_inputs_with_dummy_inputs_if_needed, decoding.rs:304-347; add_withdrawals_to_txns, :356-402)
  {
    const size_t n = b.txns.size();
    const bool wd = !b.withdrawals.empty();
    if (n == 0) {
      T.dummy_initial[0] = 0, T.dummy_initial[1] = 1, T.first_txn_ir = 2, T.n_ir = 2;
    } else if (n == 1 && !wd) {
      T.dummy_initial[0] = 0, T.first_txn_ir = 1, T.n_ir = 2;
    } else if (wd) {
      T.first_txn_ir = 0, T.dummy_final = (int)n, T.n_ir = (uint32_t)n + 1;
    } else {
      T.first_txn_ir = 0, T.n_ir = (uint32_t)n;
    }
  }
  T.n_msgs = m, T.n_ops1 = op, T.n_recs = rec, T.n_items = (uint32_t)n_items, T.val_writes = val, T.val_extra = 0;
  T.est_nodes = 0;
  T.est_nodes = 16ull * (T.n_ops1 + T.n_ops2) + 16ull * b.txns.size() + 80ull * b.withdrawals.size() + 1024;
  T.est_children = 48ull * T.n_ops1 + 128ull * T.n_ops2 + 64ull * b.txns.size() + 1100ull * b.withdrawals.size() + 4096;
  return true;
}

namespace {
struct PlanWriter {
  TxnTables& T;
  size_t lit_from;
  explicit PlanWriter(TxnTables& t) : T(t), lit_from(t.lit.size()) {}
  void seg(uint32_t a, uint32_t b, uint32_t c) { T.seg_a.push_back(a), T.seg_b.push_back(b), T.seg_c.push_back(c); }
  void flush() {
    const size_t len = T.lit.size() - lit_from;
    if (len) seg((uint32_t)len, IR_SEG_LIT_DEV, (uint32_t)lit_from);
    lit_from = T.lit.size();
  }
  void u8(uint8_t v) { T.lit.push_back(v); }
  void u32(uint32_t v) {
    uint8_t t[4];
    memcpy(t, &v, 4);
    T.lit.append(t, 4);
  }
  void raw(const uint8_t* p, size_t n) {
    if (n) T.lit.append(p, n);
  }
  void u256(uint64_t v) {
    uint8_t be[32];
    memset(be, 0, 32);
    for (int i = 0; i < 8; i++) be[31 - i] = (uint8_t)(v >> (8 * i));
    T.lit.append(be, 32);
  }
  void flat(uint32_t off, uint32_t len) {
    flush();
    if (len) seg(len, IR_SEG_FLAT, off);
  }
  uint32_t placeholder(uint32_t kind, uint32_t count) {  // segments the device fills in
    flush();
    const uint32_t at = (uint32_t)T.seg_a.size();
    for (uint32_t k = 0; k < count; k++) seg(0, kind, 0);
    return at;
  }
};
}  // namespace

bool txn_tables_phase2(const BlockJob& b, const uint8_t* flat, const TxnBases& B, const uint8_t* (*code_digest)(void*, uint32_t), void* cd_arg,
                       TxnTables& T) {
  T.txns.resize(b.txns.size());
  T.txn_keys.resize(12 * b.txns.size());
  memset(T.txn_keys.data(), 0, T.txn_keys.size());
  T.seg_a.clear(), T.seg_b.clear(), T.seg_c.clear(), T.lit.clear();
  T.seg_begin.resize(T.n_ir + 1), T.seg_end.resize(T.n_ir + 1), T.touched_begin.resize(T.n_ir + 1);
  for (uint32_t i = 0; i <= T.n_ir; i++) T.seg_begin[i] = T.seg_end[i] = 0, T.touched_begin[i] = 0;
  uint32_t t = 0, op = 0, op2 = 0, val = B.val_base + T.val_writes;
  uint64_t touched = 0, gas_before = 0;
  PlanWriter W(T);
  struct CodeEntry {
    H256 h;
    Span bytes;
  };
  std::vector<CodeEntry> code;
  for (size_t ti = 0; ti < b.txns.size(); ti++) {
    const TxnV& tx = b.txns[ti];
    TxnDesc d;
    memset(&d, 0, sizeof d);
    d.trace_begin = t;
    const uint32_t ntr = (uint32_t)tx.traces.size();
    // ---- code map (processed_block_trace.rs:269-281), sorted by hash as the IrDump wants it ----
    code.clear();
    {
      H256 e;
      memcpy(e.b, EMPTY_CODE_HASH, 32);
      code.push_back({e, Span{}});
    }
    uint32_t items = 0, ops2 = 0;
    for (uint32_t k = 0; k < ntr; k++, t++) {
      const TraceV& tr = tx.traces[k];
      const txn::TxnTrace& dt = T.traces[t];
      items += dt.n_keys;
      ops2++;
      if (tr.flags & PPD_TR_CODE_READ) {
        H256 h;
        memcpy(h.b, tr.code_read, 32);
        bool have = false;
        for (const CodeEntry& c : code) have |= c.h == h;
        if (have) continue;
        auto f = b.pre_code.find(h);
        if (f != b.pre_code.end()) {
          code.push_back({h, f->second});
        } else {
          auto g = b.resolved_code.find(h);
          if (g == b.resolved_code.end()) return false;  // PPD_ERR_UNRESOLVED_CODE_HASH, reported by the host path in order
          code.push_back({h, g->second});
        }
      } else if (tr.flags & PPD_TR_CODE_WRITE) {
        H256 h;
        memcpy(h.b, code_digest(cd_arg, t), 32);
        bool have = false;
        for (CodeEntry& c : code)
          if (c.h == h) c.bytes = tr.code_write, have = true;
        if (!have) code.push_back({h, tr.code_write});
      }
    }
    std::sort(code.begin(), code.end(), [](const CodeEntry& x, const CodeEntry& y) { return x.h < y.h; });
    d.trace_end = t;
    // ---- receipt bytes (process_rlped_receipt_node_bytes, processed_block_trace.rs:335-343) ----
    Span receipt = tx.new_receipt_node;
    if (!is_legacy_receipt(receipt.p, receipt.n)) {
      RlpItem it;
      if (!rlp_item(receipt.p, receipt.n, it) || it.is_list) return false;  // PPD_PANIC_RECEIPT_DECODE
      receipt = Span{it.payload, (uint32_t)it.payload_len};
    }
    d.off_txn_bytes = (uint32_t)(tx.byte_code.p - flat), d.len_txn_bytes = tx.byte_code.n;
    d.off_receipt = receipt.n ? (uint32_t)(receipt.p - flat) : 0, d.len_receipt = receipt.n;
    val = (val + 3) & ~3u;
    d.val_txn = val, val += tx.byte_code.n;
    val = (val + 3) & ~3u;
    d.val_receipt = val, val += receipt.n;
    // ---- the txn index key ----
    {
      uint8_t be[32];
      memset(be, 0, 32);
      for (int k = 0; k < 8; k++) be[31 - k] = (uint8_t)((uint64_t)ti >> (8 * k));
      std::vector<uint8_t> enc;
      rlp_u256(enc, be);
      memcpy(T.txn_keys.data() + 12 * ti, enc.data(), enc.size());
      d.key_off = B.txn_key_base + 12u * (uint32_t)ti, d.key_nibs = 2u * (uint32_t)enc.size();
    }
    d.op1_begin = op;
    for (uint32_t k = d.trace_begin; k < d.trace_end; k++) op += T.traces[k].n_keys;
    op += 2;
    d.op1_end = op;
    d.op2_begin = op2, op2 += ops2, d.op2_end = op2;
    // ---- the IrDump entry (include/ppd_flat.h) as segments ----
    const uint32_t ir = T.first_txn_ir + (uint32_t)ti;
    T.seg_begin[ir] = (uint32_t)T.seg_a.size();
    T.touched_begin[ir] = (uint32_t)touched;
    d.touched_base = (uint32_t)touched;
    touched += (uint64_t)txn::MARK_SLOTS_T * (ntr + 2 + items);
    const uint64_t gas_after = gas_before + tx.gas_used;
    W.u256(ti), W.u256(gas_before), W.u256(gas_after);
    W.u8(tx.byte_code.n != 0);
    W.flat((uint32_t)(tx.byte_code.p - 4 - flat), 4 + tx.byte_code.n);  // u32 length + bytes, as in the FlatBlock
    W.u32(0);                                                           // no withdrawals on a txn's entry
    d.seg_tries = W.placeholder(NODE_EMPTY, 3);
    W.u32(ntr);
    d.seg_storage = W.placeholder(NODE_EMPTY, 2 * ntr);
    d.seg_roots = W.placeholder(IR_SEG_REF, 3);
    W.flat((uint32_t)(b.checkpoint - flat), 32);
    W.u32((uint32_t)code.size());
    for (const CodeEntry& c : code) {
      W.raw(c.h.b, 32);
      W.u32(c.bytes.n);
      if (c.bytes.n) W.flat((uint32_t)(c.bytes.p - flat), c.bytes.n);
    }
    // u32 b_meta_len, bytes, u32 b_hashes_len, bytes: the same bytes as in the FlatBlock
    W.flat((uint32_t)(b.b_meta.p - 4 - flat), 4 + b.b_meta.n + 4 + b.b_hashes.n);
    W.flush();
    T.seg_end[ir] = (uint32_t)T.seg_a.size();
    gas_before = gas_after;
    T.txns[ti] = d;
  }
  // dummy entries touch nothing (their tries keep only their roots): empty ranges of the touched list
  for (uint32_t i = 0; i <= T.n_ir; i++)
    if (i < T.first_txn_ir || i >= T.first_txn_ir + b.txns.size()) T.touched_begin[i] = i < T.first_txn_ir ? 0 : (uint32_t)touched;
  T.val_extra = val - B.val_base;
  return true;
}

// ---- dummy entries (create_dummy_gen_input, decoding.rs:484-549): every trie cut with the key 0_u64, which converts to
// zero nibbles, so only the root of every trie is kept; EVERY storage trie of the state is listed ----
void txn_tables_dummies(const BlockJob& b, const uint8_t* flat, const txn::Cursors& cur, const txn::AcctExport* accounts, uint32_t n_accounts,
                        const txn::AcctState* table, uint32_t table_slots, const uint8_t* digests, TxnTables& T) {
  struct Entry {
    H256 haddr;
    uint32_t root;
  };
  std::vector<Entry> initial, final_;
  for (uint32_t r = 0; r < n_accounts; r++) {
    H256 h;
    memcpy(h.b, accounts[r].haddr, 32);
    if (accounts[r].initial != txn::ST_ABSENT) initial.push_back({h, accounts[r].initial});
    if (accounts[r].final_ != txn::ST_ABSENT) final_.push_back({h, accounts[r].final_});
  }
  for (uint32_t k = 0; k < table_slots; k++) {  // accounts the txns created
    const txn::AcctState& a = table[k];
    if (a.owner == 0xffffffffu || a.pre_rec != txn::NONE || a.storage == txn::ST_ABSENT) continue;
    H256 h;
    memcpy(h.b, digests + 32ull * a.owner, 32);
    final_.push_back({h, a.storage});
  }
  auto by_addr = [](const Entry& x, const Entry& y) { return x.haddr < y.haddr; };
  std::sort(initial.begin(), initial.end(), by_addr);
  std::sort(final_.begin(), final_.end(), by_addr);
  uint64_t gas = 0;
  for (const TxnV& tx : b.txns) gas += tx.gas_used;
  PlanWriter W(T);
  auto emit = [&](int ir, bool on_final, bool with_withdrawals) {
    T.seg_begin[ir] = (uint32_t)T.seg_a.size();
    W.u256(b.txns.size()), W.u256(gas), W.u256(gas);
    W.u8(0), W.u32(0);  // no signed txn
    if (with_withdrawals)
      W.flat((uint32_t)(b.withdrawals[0].first - 4 - flat), 4 + 52 * (uint32_t)b.withdrawals.size());  // u32 count + (address, amount) pairs, as in the FlatBlock
    else
      W.u32(0);
    W.flush();
    W.seg(on_final ? cur.state_before_withdrawals : b.state_root, IR_SEG_ROOT_ONLY, 0);
    W.seg(on_final ? cur.txn_root : NODE_EMPTY, IR_SEG_ROOT_ONLY, 0);
    W.seg(on_final ? cur.receipt_root : NODE_EMPTY, IR_SEG_ROOT_ONLY, 0);
    const std::vector<Entry>& storage = on_final ? final_ : initial;
    W.u32((uint32_t)storage.size());
    for (const Entry& e : storage) {
      W.raw(e.haddr.b, 32);
      W.flush();
      W.seg(e.root, IR_SEG_ROOT_ONLY, 0);
    }
    W.flush();
    const uint32_t r_state = with_withdrawals ? cur.roots[txn::XR_AFTER_WITHDRAWALS] : on_final ? cur.roots[txn::XR_FINAL_STATE] : cur.roots[txn::XR_INITIAL_STATE];
    W.seg(r_state, IR_SEG_REF, 0);
    W.seg(on_final ? cur.roots[txn::XR_FINAL_TXN] : cur.roots[txn::XR_EMPTY], IR_SEG_REF, 0);
    W.seg(on_final ? cur.roots[txn::XR_FINAL_RECEIPT] : cur.roots[txn::XR_EMPTY], IR_SEG_REF, 0);
    W.flat((uint32_t)(b.checkpoint - flat), 32);
    W.u32(0);  // a dummy carries no contract code
    W.flat((uint32_t)(b.b_meta.p - 4 - flat), 4 + b.b_meta.n + 4 + b.b_hashes.n);
    W.flush();
    T.seg_end[ir] = (uint32_t)T.seg_a.size();
  };
  const bool wd = !b.withdrawals.empty();
  if (T.dummy_initial[0] >= 0) emit(T.dummy_initial[0], false, false);
  if (T.dummy_initial[1] >= 0) emit(T.dummy_initial[1], false, wd);  // no txns: the second dummy carries the withdrawals
  if (T.dummy_final >= 0) emit(T.dummy_final, true, wd);
}

}  // namespace ppd
