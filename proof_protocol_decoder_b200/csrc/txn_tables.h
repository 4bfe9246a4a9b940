// txn_tables.h — what the host lays out for the device txn loop (txn_core.h) from a FlatBlock: one descriptor per
// TxnTrace and per txn (offsets into the FlatBlock, which is resident in HBM), and the plan of every IrDump entry
// (decoding.rs:131-145) as segments: literal bytes, FlatBlock ranges, and the tries / roots / hashed addresses the
// device fills in.  The host reads the flat input and writes small literals; it hashes nothing and shapes no trie.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>

#include "host_arena.h"  // PVec, Fail
#include "flat_maps.h"
#include "txn_core.h"

namespace ppd {

struct BlockJob;

struct TxnTables {
  PVec<txn::TxnTrace> traces;
  PVec<txn::TxnDesc> txns;
  PVec<txn::Withdrawal> withdrawals;
  PVec<uint32_t> seg_a, seg_b, seg_c, seg_begin, seg_end, touched_begin;
  PVec<uint8_t> lit, txn_keys;
  std::vector<uint32_t> code_write_traces;  // traces with a code write: the code maps need their digests
  // counts
  uint32_t n_msgs = 0, n_ops1 = 0, n_ops2 = 0, max_ops = 0, max_traces = 0, n_items = 0, n_recs = 0, n_ir = 0;
  // IrDump entries: txn i is entry first_txn_ir + i; the dummies of decoding.rs:304-347 / 356-402 (-1: none)
  uint32_t first_txn_ir = 0;
  int dummy_initial[2] = {-1, -1}, dummy_final = -1;
  bool needs_dummies() const { return dummy_initial[0] >= 0 || dummy_final >= 0; }
  uint32_t max_trace_keys = 0;  // the most storage keys any trace has
  uint32_t val_writes = 0;  // val_pool bytes of the written slot values (36 each)
  uint32_t val_extra = 0;   // val_pool bytes the loop writes in all (written values, txn bytes, receipts), from val_base (phase 2)
  uint64_t est_nodes = 0, est_children = 0;
  void set_allocator(PvecAlloc a, PvecFree f) {
    traces.alloc_fn = a, traces.free_fn = f, txns.alloc_fn = a, txns.free_fn = f;
    PVec<uint32_t>* u[] = {&seg_a, &seg_b, &seg_c, &seg_begin, &seg_end, &touched_begin};
    withdrawals.alloc_fn = a, withdrawals.free_fn = f;
    for (auto* x : u) x->alloc_fn = a, x->free_fn = f;
    lit.alloc_fn = a, lit.free_fn = f, txn_keys.alloc_fn = a, txn_keys.free_fn = f;
  }
  void clear() {
    traces.clear(), txns.clear(), withdrawals.clear(), seg_a.clear(), seg_b.clear(), seg_c.clear(), seg_begin.clear(), seg_end.clear(), touched_begin.clear();
    lit.clear(), txn_keys.clear(), code_write_traces.clear();
    n_msgs = n_ops1 = n_ops2 = max_ops = max_traces = n_items = n_recs = n_ir = 0;
    first_txn_ir = 0, dummy_initial[0] = dummy_initial[1] = dummy_final = -1;
    max_trace_keys = 0, val_writes = val_extra = 0, est_nodes = est_children = 0;
  }
};

// Where the loop's additions start in the device pools (all known before the loop runs, except nodes / children / the
// keys of collapsed branches, which are allocated by the loop from cursors).
struct TxnBases {
  uint32_t dig_base;      // key_pool: digest 0
  uint32_t txn_key_base;  // key_pool: the txn index keys
  uint32_t key_cursor;    // key_pool: first free byte after them
  uint32_t val_base;      // val_pool: first byte the loop writes
  uint32_t rec_base;      // accounts: first record the loop writes
};

// phase 1 (before anything is known about the witness): trace descriptors, message and op counts.
// Returns false when the block is not one the device loop takes (the host path decodes it).
bool txn_tables_phase1(const BlockJob& b, const uint8_t* flat, size_t flat_len, TxnTables& T);
// phase 2 (pool sizes of the pre-image and the digests of written code known): txn descriptors and the IR plan.
// code_digest(trace) returns the Keccak-256 of the code the trace writes.  Returns false: the host path reports the error.
bool txn_tables_phase2(const BlockJob& b, const uint8_t* flat, const TxnBases& B, const uint8_t* (*code_digest)(void*, uint32_t), void* cd_arg,
                       TxnTables& T);

// phase 3, only for a block with dummy entries (at most one txn, or withdrawals), after the loop: their segments, from
// the storage map the device exported.  `table` / `digests`: the account table and the address digest of every trace.
void txn_tables_dummies(const BlockJob& b, const uint8_t* flat, const txn::Cursors& cur, const txn::AcctExport* accounts, uint32_t n_accounts,
                        const txn::AcctState* table, uint32_t table_slots, const uint8_t* digests, TxnTables& T);

}  // namespace ppd
