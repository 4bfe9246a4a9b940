// tests/cpp/txn_core_check.cpp — TEST INFRASTRUCTURE (CPU): the device txn loop (csrc/txn_core.h, the code the
// txn_loop kernel of ppd_txn.cu instantiates) run as a thread block of one thread on host memory, against the host
// form of the same loop (csrc/host_txn.cu: shape_block), which the GPU parity tests pin to the oracle bit by bit.
//
// Compared per txn, structurally (node ids differ between the two arenas, so every node gets a fingerprint of its
// kind, key nibbles, payload and children): the three tries the subsets are cut from, every touched storage trie
// in hashed-address order, the set of nodes the marking walks touched, and the three tries after the txn.
//
// Built only by `make -C proof_protocol_decoder_b200/csrc txncheck` with -DPPD_HOSTPROF (no GPU, no node hashing:
// the byte strings whose hashes shape the tries are hashed by the development stub tools/hostprof_stub.h).
//   build/txncheck <flat block file>...
#include <cstdio>
#include <set>
#include <unordered_map>

#include "../../proof_protocol_decoder_b200/csrc/host_pipeline.h"
#include "../../proof_protocol_decoder_b200/csrc/txn_core.h"
#include "../../proof_protocol_decoder_b200/csrc/txn_tables.h"
#include "../../tools/hostprof_stub.h"

using namespace ppd;

namespace {

uint64_t mix(uint64_t h, uint64_t x) {
  h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  return h ^ (h >> 33);
}
uint64_t mix_bytes(uint64_t h, const uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; i++) h = mix(h, p[i]);
  return mix(h, n);
}

// a read-only view of an arena, whichever side it comes from
struct ArenaRO {
  const NodeRec* nodes;
  const uint8_t *key_pool, *val_pool, *hash_pool;
  const uint32_t* child_pool;
  const AccountRec* accounts;
  std::unordered_map<uint32_t, uint64_t> memo;
  uint32_t nib(uint32_t koff, uint32_t i) const {
    uint8_t b = key_pool[koff + (i >> 1)];
    return (i & 1) ? (b & 15u) : (uint32_t)(b >> 4);
  }
  uint64_t fp(uint32_t n) {
    if (n == NODE_EMPTY) return 0x1111;
    if (n >= 0x80000000u) return mix_bytes(0x2222, hash_pool + 32ull * (n - 0x80000000u), 32);
    auto f = memo.find(n);
    if (f != memo.end()) return f->second;
    const NodeRec r = nodes[n];
    const uint32_t kind = r.w0 & 0xff, ns = (r.w0 >> 8) & 0xff, nl = (r.w0 >> 16) & 0xff;
    uint64_t h = mix(0x3333, kind);
    if (kind != NK_BRANCH && kind != NK_ROOT) {
      h = mix(h, nl);
      for (uint32_t i = 0; i < nl; i++) h = mix(h, nib(r.a0, ns + i));
    }
    switch (kind) {
      case NK_LEAF:
        h = mix_bytes(h, val_pool + r.a1, r.a2);
        break;
      case NK_LEAF_ACCOUNT: {
        const AccountRec& a = accounts[r.a1];
        h = mix_bytes(h, a.nonce, 32);
        h = mix_bytes(h, a.balance, 32);
        h = mix_bytes(h, a.code_hash, 32);
        if (a.storage_src == NODE_EMPTY)
          h = mix_bytes(h, a.storage_root, 32);
        else
          h = mix(h, fp(a.storage_src));
        break;
      }
      case NK_EXT:
        h = mix(h, fp(r.a1));
        break;
      case NK_BRANCH: {
        const uint32_t mask = r.a1 & 0xffff;
        h = mix(h, mask);
        for (uint32_t j = 0; j < (uint32_t)__builtin_popcount(mask); j++) h = mix(h, fp(child_pool[r.a0 + j]));
        break;
      }
      case NK_ROOT: {
        // a ROOT over a ROOT hashes to the same thing as the inner one
        uint32_t c = r.a1;
        while (c != NODE_EMPTY && c < 0x80000000u && (nodes[c].w0 & 0xff) == NK_ROOT) c = nodes[c].a1;
        h = mix(0x4444, fp(c));
        break;
      }
    }
    memo[n] = h;
    return h;
  }
};

int check_block(const std::vector<uint8_t>& flatv, const char* name) {
  const uint8_t* flat = flatv.data();
  // ---- side 1: the host txn loop ----
  Job J1;
  J1.reset(1);
  BlockJob& b1 = J1.blocks[0];
  read_flat_block(flat, flatv.size(), b1);
  collect_messages(J1, b1);
  J1.kh.run(nullptr);
  // a block the host loop fails on (an error the reference reports) must never come out of the device path as a result:
  // the device side still runs, and has to decline the block or raise a flag (either hands the block to the host path)
  bool host_failed = false;
  Fail host_err{PPD_OK, ""};
  try {
    shape_block(J1, b1);
  } catch (const Fail& e) {
    host_failed = true, host_err = e;
  }
  auto handed_over = [&](const char* how) {
    printf("%s: status %d (%s) from the host path; the device path %s\n", name, host_err.code, host_err.msg.c_str(), how);
    return 4;
  };
  // ---- side 2: the pre-image alone, then the device loop on host memory ----
  Job J2;
  J2.reset(1);
  BlockJob& b2 = J2.blocks[0];
  read_flat_block(flat, flatv.size(), b2);
  collect_messages(J2, b2);
  J2.kh.run(nullptr);
  build_pre_image(J2, b2);
  HostArena& A = J2.A;
  TxnTables T;
  if (!txn_tables_phase1(b2, flat, flatv.size(), T)) {
    if (host_failed) return handed_over("declines the block in phase 1 of its tables");
    printf("%s: not a block the device loop takes: skipped\n", name);
    return 0;
  }
  const uint32_t n_traces = (uint32_t)T.traces.size();
  TxnBases B;
  B.dig_base = (uint32_t)((A.key_pool.size() + 31) & ~(size_t)31);
  B.txn_key_base = B.dig_base + 32 * T.n_msgs;
  B.key_cursor = B.txn_key_base + 12 * (uint32_t)b2.txns.size();
  B.val_base = (uint32_t)((A.val_pool.size() + 3) & ~(size_t)3);
  B.rec_base = (uint32_t)A.accounts.size();
  // memory of the "device"
  const uint32_t n_pre_nodes = (uint32_t)A.nodes.size();
  std::vector<NodeRec> nodes(n_pre_nodes + T.est_nodes);
  std::vector<uint16_t> level(nodes.size());
  memcpy(nodes.data(), A.nodes.data(), 16ull * n_pre_nodes);
  memcpy(level.data(), A.level.data(), 2ull * n_pre_nodes);
  std::vector<uint32_t> children(A.child_pool.size() + T.est_children);
  memcpy(children.data(), A.child_pool.data(), 4 * A.child_pool.size());
  std::vector<uint8_t> keys(B.key_cursor + 65536);
  memcpy(keys.data(), A.key_pool.data(), A.key_pool.size());
  std::vector<AccountRec> accounts(A.accounts.size() + T.n_recs + 1);
  memcpy(accounts.data(), A.accounts.data(), sizeof(AccountRec) * A.accounts.size());
  txn::View v;
  memset(&v, 0, sizeof v);
  v.nodes = nodes.data(), v.level = level.data(), v.key_pool = keys.data(), v.hash_pool = A.hash_pool.data();
  v.child_pool = children.data(), v.accounts = accounts.data();
  v.cap_nodes = (uint32_t)nodes.size(), v.cap_children = (uint32_t)children.size(), v.cap_keys = (uint32_t)keys.size();
  v.flat = flat, v.traces = T.traces.data(), v.n_txns = (uint32_t)b2.txns.size(), v.n_traces = n_traces, v.dig_base = B.dig_base;
  v.rec_base = B.rec_base, v.val_base = B.val_base;
  v.withdrawals = T.withdrawals.data(), v.n_withdrawals = (uint32_t)T.withdrawals.size();
  // digests (the device runs keccak256_batch_kernel over the same (begin, end) table)
  std::vector<uint64_t> se(2ull * T.n_msgs);
  for (uint32_t t = 0; t < n_traces; t++) txn::prep_msgs(v, t, se.data());
  for (uint32_t w = 0; w < v.n_withdrawals; w++) txn::prep_withdrawal_msg(v, w, se.data());
  for (uint32_t m = 0; m < T.n_msgs; m++) hostprof::keccak256(flat + se[2 * m], se[2 * m + 1] - se[2 * m], keys.data() + B.dig_base + 32ull * m);
  struct CD {
    txn::View* v;
    TxnTables* T;
  } cd{&v, &T};
  auto code_digest = [](void* arg, uint32_t t) -> const uint8_t* {
    CD* c = (CD*)arg;
    return c->v->key_pool + c->v->dig_base + 32ull * c->T->traces[t].m_code;
  };
  if (!txn_tables_phase2(b2, flat, B, code_digest, &cd, T)) {
    if (host_failed) return handed_over("declines the block in phase 2 of its tables");
    printf("%s: phase 2 declined the block (the host path reports its error): skipped\n", name);
    return 0;
  }
  memcpy(keys.data() + B.txn_key_base, T.txn_keys.data(), T.txn_keys.size());
  std::vector<uint8_t> vals(B.val_base + T.val_extra + 64);
  memcpy(vals.data(), A.val_pool.data(), A.val_pool.size());
  v.val_pool = vals.data();
  v.txns = T.txns.data();
  // the by-root join as the host builder resolved it
  const uint32_t n_pre_acct = (uint32_t)A.accounts.size();
  std::vector<uint32_t> join_storage(n_pre_acct + 1, txn::ST_ABSENT), join_root(n_pre_acct + 1, txn::NONE);
  std::vector<uint8_t> pre_flags(n_pre_acct + 1, 0);
  for (const BlockJob::PreAccount& pa : b2.pre_accounts) {
    auto f = b2.storage.find(pa.haddr);
    if (f != b2.storage.end()) {
      join_storage[pa.rec] = f->second;
      if (const uint32_t* rn = b2.root_of.find(f->second)) join_root[pa.rec] = *rn;
    }
    pre_flags[pa.rec] = pa.storage_nonempty ? 1 : 0;
  }
  v.pre_flags = pre_flags.data();
  std::vector<uint32_t> pre_slot(n_pre_acct + 1, 0xffffffffu);
  v.pre_slot = pre_slot.data();
  uint32_t table = 64;
  while (table < 2 * n_traces) table <<= 1;
  std::vector<txn::AcctState> acct(table);
  memset(acct.data(), 0xff, sizeof(txn::AcctState) * table);
  v.acct = acct.data();
  std::vector<txn::SOp> ops1(T.n_ops1 + 1), ops2(T.n_ops2 + 1);
  v.ops1 = ops1.data(), v.ops2 = ops2.data();
  std::vector<uint32_t> touched(T.touched_begin.back() + 16, NODE_EMPTY);
  v.touched = touched.data(), v.seg_a = T.seg_a.data(), v.seg_b = T.seg_b.data();
  const size_t np = (size_t)T.max_ops * txn::PATH_CAP + 1;
  std::vector<uint32_t> path_node(np), path_pc(np), tnode(T.max_ops + 1), tpc(T.max_ops + 1), key_hi(T.max_ops + 1);
  std::vector<uint8_t> path_depth(np), plen(T.max_ops + 1), tdepth(T.max_ops + 1), tkind(T.max_ops + 1);
  std::vector<txn::SOp> sh_ops(T.max_ops + 1);
  v.s.path_node = path_node.data(), v.s.path_pc = path_pc.data(), v.s.path_depth = path_depth.data();
  v.s.plen = plen.data(), v.s.tnode = tnode.data(), v.s.tpc = tpc.data(), v.s.tdepth = tdepth.data(), v.s.tkind = tkind.data(), v.s.key_hi = key_hi.data();
  v.s.sh_ops = (n_traces & 1) ? sh_ops.data() : nullptr;  // both ways of reaching a txn's keys get exercised
  // the path-node table in two tiers, the first one tiny so that the spill tier is exercised
  std::vector<txn::PathNode> pc_fast(5), pc_slow(np);
  size_t map_n = 64;
  while (map_n < 2 * np) map_n <<= 1;
  std::vector<uint32_t> pc_map(map_n), pc_map_key(map_n);
  uint32_t pc_count = 0;
  v.s.pc_fast = pc_fast.data(), v.s.pc_n_fast = (uint32_t)pc_fast.size(), v.s.pc_slow = pc_slow.data(), v.s.pc_n_slow = (uint32_t)pc_slow.size();
  v.s.pc_map = pc_map.data(), v.s.pc_map_key = pc_map_key.data(), v.s.pc_map_mask = (uint32_t)map_n - 1, v.s.pc_count = &pc_count;
  txn::Cursors cur;
  memset(&cur, 0, sizeof cur);
  cur.n_nodes = n_pre_nodes, cur.n_children = (uint32_t)A.child_pool.size(), cur.key_bytes = B.key_cursor;
  cur.state_root = b2.state_root, cur.txn_root = NODE_EMPTY, cur.receipt_root = NODE_EMPTY;
  v.cur = &cur;
  v.s.a_nodes = &cur.n_nodes, v.s.a_children = &cur.n_children, v.s.a_keys = &cur.key_bytes, v.s.a_max_level = &cur.max_level;
  // ---- the kernels, in launch order ----
  txn::AcctInit ai{table - 1, b2.state_root, join_storage.data(), join_root.data()};
  for (uint32_t t = 0; t < n_traces; t++) txn::acct_claim(v, ai, t);
  for (uint32_t t = 0; t < n_traces; t++) txn::prep_trace(v, t);
  for (uint32_t t = 0; t < n_traces; t++)
    for (uint32_t k = 0; k < T.traces[t].n_keys; k++) txn::prep_storage_key(v, t, k);
  for (uint32_t ti = 0; ti < v.n_txns; ti++) txn::prep_txn(v, ti);
  for (uint32_t i = 0; i < T.n_ops1; i++) txn::prep_lcp(v, v.ops1, i);
  for (uint32_t i = 0; i < T.n_ops2; i++) txn::prep_lcp(v, v.ops2, i);
  long long sh_clock = 0;
  txn::Ctx c{v, 0, 1, &sh_clock};
  for (uint32_t ti = 0; ti < v.n_txns && !cur.flag; ti++) txn::run_txn(c, ti, EMPTY_TRIE_HASH, EMPTY_CODE_HASH);
  txn::run_finish(c, b2.state_root);
  if (host_failed) {
    if (cur.flag) {
      char how[96];
      snprintf(how, sizeof how, "raises flag %u at txn %u", cur.flag, cur.flag_txn);
      return handed_over(how);
    }
    printf("%s: MISMATCH: the device loop finished a block the host path fails with status %d (%s)\n", name, host_err.code, host_err.msg.c_str());
    return 1;
  }
  if (cur.flag) {
    printf("%s: the loop raised flag %u at txn %u (the host path would redo the block)\n", name, cur.flag, cur.flag_txn);
    return 2;
  }
  if (T.needs_dummies()) {
    // the storage map as the device exports it (ppd_txn.cu: acct_export_kernel)
    std::vector<txn::AcctExport> ex(n_pre_acct + 1);
    for (const BlockJob::PreAccount& pa : b2.pre_accounts) {
      txn::AcctExport e;
      memcpy(e.haddr, pa.haddr.b, 32);
      e.initial = join_storage[pa.rec];
      const uint32_t slot = pre_slot[pa.rec];
      e.final_ = slot < txn::NONE ? acct[slot].storage : e.initial;
      ex[pa.rec] = e;
    }
    txn_tables_dummies(b2, flat, cur, ex.data(), n_pre_acct, acct.data(), table, keys.data() + B.dig_base, T);
    v.seg_a = T.seg_a.data(), v.seg_b = T.seg_b.data();
  }
  // ---- the sweep's invariant: a node's level exceeds the level of everything it reads ----
  {
    auto lv = [&](uint32_t n) -> int { return (n == NODE_EMPTY || n >= 0x80000000u) ? -1 : (int)level[n]; };
    for (uint32_t n = n_pre_nodes; n < cur.n_nodes; n++) {
      const NodeRec r = nodes[n];
      const uint32_t kind = r.w0 & 0xff;
      int need = -1;
      if (kind == NK_EXT || kind == NK_ROOT) need = lv(r.a1);
      if (kind == NK_LEAF_ACCOUNT) need = lv(accounts[r.a1].storage_src);
      if (kind == NK_BRANCH)
        for (uint32_t j = 0; j < (uint32_t)__builtin_popcount(r.a1 & 0xffff); j++) need = std::max(need, lv(children[r.a0 + j]));
      if ((int)level[n] <= need) {
        printf("%s: node %u (kind %u) has level %u but reads a node of level %d\n", name, n, kind, level[n], need);
        return 1;
      }
    }
  }
  // ---- compare ----
  ArenaRO R1{J1.A.nodes.data(), J1.A.key_pool.data(), J1.A.val_pool.data(), J1.A.hash_pool.data(), J1.A.child_pool.data(), J1.A.accounts.data(), {}};
  ArenaRO R2{nodes.data(), keys.data(), vals.data(), A.hash_pool.data(), children.data(), accounts.data(), {}};
  int bad = 0;
  auto expect = [&](bool ok, uint32_t ti, const char* what) {
    if (!ok && bad++ < 10) printf("%s: txn %u: %s differs\n", name, ti, what);
  };
  expect(b1.irs.size() == T.n_ir, 0, "number of IrDump entries");
  // dummy entries: the tries (roots only), every storage trie in hashed-address order, the roots after
  auto check_dummy = [&](int ir) {
    if (ir < 0 || (size_t)ir >= b1.irs.size()) return;
    IrPlan& p = b1.irs[ir];
    uint32_t q = T.seg_begin[ir];
    auto next_kind = [&](uint32_t kind) {
      while (q < T.seg_end[ir] && v.seg_b[q] != kind) q++;
      return q < T.seg_end[ir] ? q++ : 0xffffffffu;
    };
    uint32_t s0 = next_kind(IR_SEG_ROOT_ONLY), s1 = next_kind(IR_SEG_ROOT_ONLY), s2 = next_kind(IR_SEG_ROOT_ONLY);
    expect(s0 != 0xffffffffu && s2 != 0xffffffffu, ir, "dummy: trie segments");
    if (s2 == 0xffffffffu) return;
    expect(R1.fp(p.state_sub) == R2.fp(v.seg_a[s0]), ir, "dummy: state trie");
    expect(R1.fp(p.txn_sub) == R2.fp(v.seg_a[s1]), ir, "dummy: transactions trie");
    expect(R1.fp(p.receipt_sub) == R2.fp(v.seg_a[s2]), ir, "dummy: receipts trie");
    std::stable_sort(p.storage_subs.begin(), p.storage_subs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    for (size_t k = 0; k < p.storage_subs.size(); k++) {
      uint32_t sq = next_kind(IR_SEG_ROOT_ONLY);
      expect(sq != 0xffffffffu, ir, "dummy: number of storage tries");
      if (sq == 0xffffffffu) return;
      // the hashed address is the tail of the literal segment before it
      const uint32_t lq = sq - 1;
      expect(v.seg_b[lq] == IR_SEG_LIT_DEV && memcmp(T.lit.data() + T.seg_c[lq] + v.seg_a[lq] - 32, p.storage_subs[k].first.b, 32) == 0, ir, "dummy: hashed address order");
      expect(R1.fp(p.storage_subs[k].second) == R2.fp(v.seg_a[sq]), ir, "dummy: a storage trie");
    }
    uint32_t r0 = next_kind(IR_SEG_REF), r1 = next_kind(IR_SEG_REF), r2 = next_kind(IR_SEG_REF);
    expect(r2 != 0xffffffffu, ir, "dummy: root segments");
    if (r2 == 0xffffffffu) return;
    expect(R1.fp(p.root_state) == R2.fp(v.seg_a[r0]), ir, "dummy: state root after");
    expect(R1.fp(p.root_txn) == R2.fp(v.seg_a[r1]), ir, "dummy: transactions root after");
    expect(R1.fp(p.root_receipt) == R2.fp(v.seg_a[r2]), ir, "dummy: receipts root after");
    bool has_wd_seg = false;
    for (uint32_t k = T.seg_begin[ir]; k < T.seg_end[ir]; k++)
      has_wd_seg |= v.seg_b[k] == IR_SEG_FLAT && !b2.withdrawals.empty() && T.seg_c[k] == (uint32_t)(b2.withdrawals[0].first - 4 - flat);
    expect(has_wd_seg == p.has_withdrawals, ir, "dummy: withdrawals");
  };
  check_dummy(T.dummy_initial[0]), check_dummy(T.dummy_initial[1]), check_dummy(T.dummy_final);
  for (uint32_t ti = 0; ti < v.n_txns; ti++) {
    IrPlan& p = b1.irs[T.first_txn_ir + ti];
    const txn::TxnDesc& tx = T.txns[ti];
    expect(R1.fp(p.state_sub) == R2.fp(v.seg_b[tx.seg_tries]), ti, "state trie before the txn");
    expect(R1.fp(p.txn_sub) == R2.fp(v.seg_b[tx.seg_tries + 1]), ti, "transactions trie before the txn");
    expect(R1.fp(p.receipt_sub) == R2.fp(v.seg_b[tx.seg_tries + 2]), ti, "receipts trie before the txn");
    std::stable_sort(p.storage_subs.begin(), p.storage_subs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    const uint32_t ntr = tx.trace_end - tx.trace_begin;
    expect(p.storage_subs.size() == ntr, ti, "number of storage tries");
    for (uint32_t r = 0; r < ntr && r < p.storage_subs.size(); r++) {
      expect(v.seg_b[tx.seg_storage + 2 * r] == IR_SEG_KEY32, ti, "storage segment kind");
      expect(memcmp(keys.data() + v.seg_a[tx.seg_storage + 2 * r], p.storage_subs[r].first.b, 32) == 0, ti, "hashed address order");
      expect(R1.fp(p.storage_subs[r].second) == R2.fp(v.seg_b[tx.seg_storage + 2 * r + 1]), ti, "a storage trie before the txn");
    }
    expect(R1.fp(p.root_state) == R2.fp(v.seg_a[tx.seg_roots]), ti, "state trie after the txn");
    expect(R1.fp(p.root_txn) == R2.fp(v.seg_a[tx.seg_roots + 1]), ti, "transactions trie after the txn");
    expect(R1.fp(p.root_receipt) == R2.fp(v.seg_a[tx.seg_roots + 2]), ti, "receipts trie after the txn");
    std::set<uint64_t> t1, t2;
    for (uint32_t n : p.touched) t1.insert(R1.fp(n));
    for (uint32_t k = T.touched_begin[T.first_txn_ir + ti]; k < T.touched_begin[T.first_txn_ir + ti + 1]; k++)
      if (touched[k] != NODE_EMPTY) t2.insert(R2.fp(touched[k]));
    expect(t1 == t2, ti, "set of touched nodes");
  }
  printf("%s: %u txns, %u traces, %u + %u ops; host arena %zu nodes, device loop %u nodes (%u pre-image): %s\n", name, v.n_txns, n_traces, T.n_ops1,
         T.n_ops2, J1.A.nodes.size(), cur.n_nodes, n_pre_nodes, bad ? "MISMATCH" : "identical");
  return bad ? 1 : 0;
}

}  // namespace

int main(int argc, char** argv) {
  int rc = 0;
  for (int i = 1; i < argc; i++) {
    FILE* f = fopen(argv[i], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> flat(n);
    if (fread(flat.data(), 1, n, f) != (size_t)n) return 2;
    fclose(f);
    try {
      rc |= check_block(flat, argv[i]);
    } catch (const Fail& e) {
      printf("%s: status %d (%s) from the host path\n", argv[i], e.code, e.msg.c_str());
      rc |= 4;
    }
  }
  return rc;
}
