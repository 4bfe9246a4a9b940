// txn_tables.cu — FlatBlock -> the tables of the device txn loop (see txn_tables.h).  Host only; structure only.
#include "host_pipeline.h"
#include "txn_tables.h"

namespace ppd {

using txn::TxnDesc;
using txn::TxnTrace;

bool txn_tables_phase1(const BlockJob& b, const uint8_t* flat, size_t flat_len, TxnTables& T) {
  T.clear();
  // blocks that need dummy entries or withdrawals (decoding.rs:304-428) are shaped by the host path
  if (b.txns.size() < 2 || !b.withdrawals.empty()) return false;
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
      d.op0 = op, op += tr.n_writes;
      d.item0 = item, item += tr.n_reads + tr.n_writes;
      if (d.flags & txn::TRF_STATE_WRITE) d.rec = rec++;
      d.val0 = val, val += 36u * tr.n_writes;
      if (d.flags & (txn::TRF_STATE_WRITE | PPD_TR_SELF_DESTRUCTED)) ops2++;
      T.traces[t++] = d;
    }
    op += 2;  // the inserts into the transactions and receipts tries
    T.n_ops2 += ops2;
    T.max_ops = std::max(T.max_ops, std::max(op - op_begin, ops2));
    T.max_traces = std::max<uint32_t>(T.max_traces, (uint32_t)tx.traces.size());
    n_items += tx.traces.size() + 2 + item;
  }
  if (n_items * txn::MARK_SLOTS_T >= (1ull << 31)) return false;
  T.n_msgs = m, T.n_ops1 = op, T.n_recs = rec, T.n_items = (uint32_t)n_items, T.val_writes = val, T.val_extra = 0;
  T.est_nodes = 16ull * (T.n_ops1 + T.n_ops2) + 16ull * b.txns.size() + 1024;
  T.est_children = 48ull * T.n_ops1 + 128ull * T.n_ops2 + 64ull * b.txns.size() + 4096;
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
  T.seg_a.clear(), T.seg_b.clear(), T.seg_c.clear(), T.seg_begin.clear(), T.touched_begin.clear(), T.lit.clear();
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
      txn::TxnTrace& dt = T.traces[t];
      if (dt.flags & txn::TRF_STATE_WRITE) dt.rec += B.rec_base;
      dt.val0 += B.val_base;
      items += tr.n_reads + tr.n_writes;
      if (dt.flags & (txn::TRF_STATE_WRITE | PPD_TR_SELF_DESTRUCTED)) ops2++;
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
    for (uint32_t k = d.trace_begin; k < d.trace_end; k++) op += T.traces[k].n_writes;
    op += 2;
    d.op1_end = op;
    d.op2_begin = op2, op2 += ops2, d.op2_end = op2;
    // ---- the IrDump entry (include/ppd_flat.h) as segments ----
    T.seg_begin.push_back((uint32_t)T.seg_a.size());
    T.touched_begin.push_back((uint32_t)touched);
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
    gas_before = gas_after;
    T.txns[ti] = d;
  }
  T.seg_begin.push_back((uint32_t)T.seg_a.size());
  T.touched_begin.push_back((uint32_t)touched);
  T.n_ir = (uint32_t)b.txns.size();
  T.val_extra = val - B.val_base;
  return true;
}

}  // namespace ppd
