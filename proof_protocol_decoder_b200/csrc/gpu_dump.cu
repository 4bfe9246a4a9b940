// gpu_dump.cu — host side of the IrDump serialisation on the GPU (kernels: ppd_dump.cu).
#include "host_pipeline.h"

namespace ppd {

// ---- step 5 on the GPU (ppd_dump.cu): the host lays out each IR as literals and tries; the device sizes,
// places and writes every trie; the host fills the literals in.  IRs the device flags are serialised by
// dump_ir.  Returns false when the block has to take the host path altogether.
struct LitPool {
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u32(uint32_t v) {
    uint8_t t[4];
    memcpy(t, &v, 4);
    b.insert(b.end(), t, t + 4);
  }
  void raw(const uint8_t* p, size_t n) {
    if (n) b.insert(b.end(), p, p + n);
  }
  void span(Span s) {
    u32(s.n);
    raw(s.p, s.n);
  }
  void u256(uint64_t v) {
    uint8_t be[32];
    memset(be, 0, 32);
    for (int i = 0; i < 8; i++) be[31 - i] = (uint8_t)(v >> (8 * i));
    raw(be, 32);
  }
};

bool gpu_dump_enabled() {
#ifdef PPD_HOSTPROF
  return false;
#else
  static const bool disabled = getenv("PPD_HOST_DUMP") != nullptr;
  return !disabled;
#endif
}

int gpu_dump_block(ppd_ctx* c, Lane* L, Job& J, uint8_t** out, size_t* out_len) {
#ifdef PPD_HOSTPROF
  return DUMP_ON_HOST;
#else
  const bool disabled = !gpu_dump_enabled();
  static const bool verify = getenv("PPD_VERIFY_GPU_DUMP") != nullptr;
  BlockJob& b = J.blocks[0];
  const uint32_t n_ir = (uint32_t)b.irs.size();
  if (disabled || !L->has_last || n_ir == 0 || J.A.nodes.size() == 0) return DUMP_ON_HOST;
  PhaseTimer pt;
  // ---- plan ----
  LitPool lit;
  std::vector<uint32_t> seg_a, seg_b, seg_begin(1, 0), touched_begin(1, 0), lit_at;  // lit_at[seg]: offset into lit.b
  size_t n_touched = 0;
  for (IrPlan& p : b.irs) n_touched += p.touched.size();
  {
    // the two dump kernels cost two read-backs and about a millisecond of launch + CTA latency whatever the size;
    // a small block (config 1: a few thousand touched nodes) is serialised faster by the host threads
    const char* e = getenv("PPD_GPU_DUMP_MIN_TOUCHED");
    const size_t min_touched = e ? (size_t)atoll(e) : 32768;
    if (n_touched < min_touched) return DUMP_ON_HOST;
  }
  auto add_lit_from = [&](size_t from) {  // the bytes appended to lit.b since `from` become (part of) a literal segment
    uint32_t len = (uint32_t)(lit.b.size() - from);
    if (!len) return;
    if (seg_b.size() > seg_begin.back() && seg_b.back() == IR_SEG_LITERAL) {
      seg_a.back() += len;
    } else {
      seg_a.push_back(len), seg_b.push_back(IR_SEG_LITERAL), lit_at.push_back((uint32_t)from);
    }
  };
  auto add_trie = [&](uint32_t root) { seg_a.push_back(0), seg_b.push_back(root), lit_at.push_back(0); };
  auto add_ref = [&](uint32_t node) { seg_a.push_back(node), seg_b.push_back(IR_SEG_REF), lit_at.push_back(0); };
  for (IrPlan& p : b.irs) {
    size_t from = lit.b.size();
    lit.u256(p.txn_before), lit.u256(p.gas_before), lit.u256(p.gas_after);
    lit.u8(p.has_signed_txn);
    lit.span(p.has_signed_txn ? p.signed_txn : Span{});
    if (p.has_withdrawals) {
      lit.u32((uint32_t)b.withdrawals.size());
      for (auto& w : b.withdrawals) lit.raw(w.first, 20), lit.raw(w.second, 32);
    } else {
      lit.u32(0);
    }
    add_lit_from(from);
    add_trie(p.state_sub), add_trie(p.txn_sub), add_trie(p.receipt_sub);
    std::stable_sort(p.storage_subs.begin(), p.storage_subs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    from = lit.b.size();
    lit.u32((uint32_t)p.storage_subs.size());
    for (auto& sub : p.storage_subs) {
      lit.raw(sub.first.b, 32);
      add_lit_from(from);
      add_trie(sub.second);
      from = lit.b.size();
    }
    add_lit_from(from);
    add_ref(p.root_state), add_ref(p.root_txn), add_ref(p.root_receipt);  // TrieRoots: 32 bytes each, straight from the refs in HBM
    from = lit.b.size();
    lit.raw(b.checkpoint, 32);
    lit.u32((uint32_t)p.code.size());
    for (auto& cd : p.code) lit.raw(cd.first.b, 32), lit.span(cd.second);
    lit.span(b.b_meta);
    lit.span(b.b_hashes);
    add_lit_from(from);
    seg_begin.push_back((uint32_t)seg_a.size());
    touched_begin.push_back((uint32_t)(touched_begin.back() + p.touched.size()));
  }
  const uint32_t n_seg = (uint32_t)seg_a.size();
  // one pinned buffer / one device buffer: [touched | touched_begin | seg_a | seg_b | seg_begin | ir_base(u64) |
  //                                        seg_off | ir_size | ir_flag | ir_nuniq | u_node | u_size | u_off]
  auto al = [](size_t x) { return (x + 3) & ~(size_t)3; };  // keep the u64 array 8-byte aligned (counts in u32 words)
  const size_t o_touched = 0, o_tb = al(o_touched + n_touched), o_sa = al(o_tb + n_ir + 1), o_sb = al(o_sa + n_seg), o_sg = al(o_sb + n_seg),
               o_base = al(o_sg + n_ir + 1), o_in_end = al(o_base + 2 * (size_t)n_ir);
  const size_t o_soff = o_in_end, o_isz = al(o_soff + n_seg), o_ifl = al(o_isz + n_ir), o_inu = al(o_ifl + n_ir),
               o_un = al(o_inu + n_ir), o_us = al(o_un + n_touched), o_uo = al(o_us + n_touched), o_end = al(o_uo + n_touched);
  J.plan.resize(o_end);
  uint32_t* h = J.plan.data();
  {
    uint32_t* t = h + o_touched;
    for (uint32_t ir = 0; ir < n_ir; ir++) {
      IrPlan& p = b.irs[ir];
      if (!p.touched.empty()) memcpy(t, p.touched.data(), 4 * p.touched.size());
      t += p.touched.size();
    }
  }
  memcpy(h + o_tb, touched_begin.data(), 4 * (n_ir + 1));
  memcpy(h + o_sa, seg_a.data(), 4 * n_seg);
  memcpy(h + o_sb, seg_b.data(), 4 * n_seg);
  memcpy(h + o_sg, seg_begin.data(), 4 * (n_ir + 1));
  L->d_plan.reserve(4 * o_end);
  uint32_t* d = L->d_plan.as<uint32_t>();
  CUDA_OK(cudaMemcpyAsync(d, h, 4 * o_base, cudaMemcpyHostToDevice, L->st));
  IrDumpPlanView P{};
  P.touched = d + o_touched, P.touched_begin = d + o_tb, P.seg_a = d + o_sa, P.seg_b = d + o_sb, P.seg_begin = d + o_sg;
  P.ir_base = reinterpret_cast<const uint64_t*>(d + o_base);
  P.seg_off = d + o_soff, P.ir_size = d + o_isz, P.ir_flag = d + o_ifl, P.ir_nuniq = d + o_inu;
  P.u_node = d + o_un, P.u_size = d + o_us, P.u_off = d + o_uo;
  pt.lap("  d:plan");
  launch_ir_size(L->last_view, P, n_ir, L->st);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(h + o_soff, d + o_soff, 4 * (o_inu - o_soff), cudaMemcpyDeviceToHost, L->st));  // seg_off, ir_size, ir_flag
  lane_sync(L);
  L->stats.h2d_bytes += 4.0 * o_base, L->stats.d2h_bytes += 4.0 * (o_inu - o_soff), L->stats.kernel_launches += 1;
  pt.lap("  d:size");
  // ---- IRs the device could not lay out: host serialisation ----
  const uint32_t *seg_off = h + o_soff, *ir_size = h + o_isz;
  uint32_t* ir_flag = h + o_ifl;
  std::vector<Out> host_parts(n_ir);
  Stamp st;
  uint64_t* ir_base = reinterpret_cast<uint64_t*>(h + o_base);
  uint64_t total = 8;
  for (uint32_t i = 0; i < n_ir; i++) {
    ir_base[i] = total;
    if (ir_flag[i]) {
      if (st.v.empty()) st.v.assign(J.A.nodes.size(), 0);
      fetch_refs(L, J);
      fetch_pools(L, J);
      dump_ir(J, b, b.irs[i], st, host_parts[i]);
      total += host_parts[i].n;
    } else {
      total += ir_size[i];
    }
  }
  // ---- emit, copy back, fill the literals in ----
  Out o;
  // pool buffers are 8 MiB-granular: a small IrDump (a config-1 block is 85 KB) lands in the lane's staging buffer instead
  uint8_t* pinned = total >= ((size_t)1 << 20) ? out_pool().take(total) : nullptr;
  if (!pinned) o.need(total);
  uint8_t* dst = pinned ? pinned : o.p;
  L->d_out.reserve(total + 64);
  CUDA_OK(cudaMemcpyAsync(d + o_base, h + o_base, 8 * (size_t)n_ir, cudaMemcpyHostToDevice, L->st));
  launch_ir_emit(L->last_view, P, n_ir, L->d_out.as<uint8_t>(), L->st);
  CUDA_OK(cudaGetLastError());
  if (pinned) {
    // straight into the caller's (page-locked) buffer
    cudaError_t e = cudaMemcpyAsync(dst, L->d_out.p, total, cudaMemcpyDeviceToHost, L->st);
    if (e == cudaSuccess) e = cudaEventRecord(L->ev_sync, L->st);
    if (e == cudaSuccess) e = cudaEventSynchronize(L->ev_sync);
    if (e != cudaSuccess) {
      out_pool().give_back(pinned);
      throw Fail{PPD_ERR_CUDA, std::string("IR dump copy: ") + cudaGetErrorString(e)};
    }
  } else {
    // The output is pageable memory: land the copy in a page-locked buffer in chunks (full-rate,
    // truly asynchronous DMA) and move each chunk on while the next one is in flight.
    J.out_stage.resize(total);
    const size_t CH = 8u << 20;
    size_t n_ch = (total + CH - 1) / CH;
    std::vector<cudaEvent_t> evs(n_ch);
    for (size_t k = 0; k < n_ch; k++) {
      size_t at = k * CH, len = std::min(CH, (size_t)total - at);
      CUDA_OK(cudaMemcpyAsync(J.out_stage.data() + at, L->d_out.as<uint8_t>() + at, len, cudaMemcpyDeviceToHost, L->st));
      CUDA_OK(cudaEventCreateWithFlags(&evs[k], cudaEventBlockingSync | cudaEventDisableTiming));
      CUDA_OK(cudaEventRecord(evs[k], L->st));
    }
    for (size_t k = 0; k < n_ch; k++) {
      size_t at = k * CH, len = std::min(CH, (size_t)total - at);
      cudaError_t e = cudaEventSynchronize(evs[k]);
      cudaEventDestroy(evs[k]);
      if (e != cudaSuccess) {
        for (size_t r = k + 1; r < n_ch; r++) cudaEventDestroy(evs[r]);
        throw Fail{PPD_ERR_CUDA, std::string("cudaEventSynchronize: ") + cudaGetErrorString(e)};
      }
      memcpy(o.p + at, J.out_stage.data() + at, len);
    }
  }
  L->stats.d2h_bytes += (double)total, L->stats.kernel_launches += 1;
  o.n = total;
  pt.lap("  d:emit+copy");
  uint32_t hdr[2] = {PPD_IR_DUMP_MAGIC, n_ir};
  memcpy(dst, hdr, 8);
  for (uint32_t i = 0; i < n_ir; i++) {
    uint8_t* base = dst + ir_base[i];
    if (ir_flag[i]) {
      memcpy(base, host_parts[i].p, host_parts[i].n);
      continue;
    }
    for (uint32_t q = seg_begin[i]; q < seg_begin[i + 1]; q++)
      if (seg_b[q] == IR_SEG_LITERAL) memcpy(base + seg_off[q], lit.b.data() + lit_at[q], seg_a[q]);
  }
  pt.lap("  d:literals");
  if (verify) {
    uint8_t* want = nullptr;
    size_t want_len = 0;
    fetch_refs(L, J);
    fetch_pools(L, J);
    dump_blocks(J, &want, &want_len, 1);
    bool same = want_len == o.n && memcmp(want, dst, o.n) == 0;
    size_t at = 0;
    if (!same)
      while (at < want_len && at < o.n && want[at] == dst[at]) at++;
    free(want);
    if (!same) {
      if (pinned) out_pool().give_back(pinned);
      std::lock_guard<std::mutex> g(c->err_mu);
      throw Fail{PPD_ERR_CUDA, "GPU IR dump differs from the host dump at byte " + std::to_string(at) + " (sizes " + std::to_string(o.n) + " / " + std::to_string(want_len) + ")"};
    }
  }
  if (pinned) {
    *out = pinned, *out_len = total;
  } else {
    *out = o.give(out_len);
  }
  return DUMP_DONE;
#endif
}

}  // namespace ppd
