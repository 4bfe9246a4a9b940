// ppd_host.cu — host side of libppd_b200.so: the C ABI (include/ppd_b200.h), the context that
// owns the stream and the HBM buffers, and the block pipeline
//
//   FlatBlock ─► parse witness (compact_prestate_processing.rs:683-875, 387-668)
//             ─► batch-hash every address / slot / code on the GPU      (utils.rs:11-13 call sites)
//             ─► shape all versions of all tries as one persistent DAG   (host_arena.h; no hashing)
//             ─► ONE level-synchronous GPU sweep over the DAG            (ppd_kernels.cu)
//             ─► serialise Vec<GenerationInputs> as an IrDump            (decoding.rs:131-145)
//
// The host never computes a Keccak or a node encoding.  If the CUDA device or the kernels are not
// usable every entry point fails with PPD_ERR_CUDA: there is no CPU fallback.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ppd_b200.h"
#include "arena.h"
#include "devbuf.h"
#include "host_arena.h"
#include "flat_maps.h"
#include "ppd_kernels.h"
#ifdef PPD_HOSTPROF
#include "hostprof_stub.h"  // development-only host profiler build (tools/hostprof); never defined for libppd_b200.so
#endif

using namespace ppd;

namespace {

const uint8_t EMPTY_CODE_HASH[32] = {0xc5, 0xd2, 0x46, 0x01, 0x86, 0xf7, 0x23, 0x3c, 0x92, 0x7e, 0x7d, 0xb2, 0xdc, 0xc7, 0x03, 0xc0,
                                     0xe5, 0x00, 0xb6, 0x53, 0xca, 0x82, 0x27, 0x3b, 0x7b, 0xfa, 0xd8, 0x04, 0x5d, 0x85, 0xa4, 0x70};
const uint8_t EMPTY_TRIE_HASH[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                     0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

}  // namespace

namespace {
struct Job;
void job_delete(Job*);
#ifdef PPD_HOSTPROF
void* pinned_alloc(size_t n) { return malloc(n); }
void pinned_free(void* p) { free(p); }
#else
void* pinned_alloc(size_t n) {
  void* p = nullptr;
  return cudaHostAlloc(&p, n, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void pinned_free(void* p) { cudaFreeHost(p); }
#endif
}  // namespace

// One lane of the block pipeline: a stream, its HBM buffers and the host-side scratch of one block.
// Blocks of a batch are decoded concurrently, one lane per host thread; the lanes' kernels and copies
// overlap on the device.
struct Lane {
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_sync = nullptr;  // blocking-sync event: a waiting host thread sleeps instead of spinning
  ppd_stats stats{};
  DevBuf d_nodes, d_order, d_keys, d_vals, d_hashes, d_children, d_accounts, d_ref, d_ref_len, d_counters;
  DevBuf d_msg, d_msg_off, d_digest;
  DevBuf d_plan, d_out;  // IR dump plan and the serialised IrDump
  DevBuf d_wit, d_pa, d_pb, d_pc;  // witness bytes and the scratch of the three parse phases (ppd_parse.cu)
  DevBuf d_level, d_okeys, d_obins;  // node levels, and the scratch of the (level, class) ordering on the device
  uint32_t* h_parse = nullptr;     // page-locked landing area of the parse result words
  // the launch parameters of the lane's last GPU parse (the witness and all scratch stay resident), for ppd_replay_last_parse
  bool has_last_parse = false;
  ParseBounds last_bounds{};
  ParseEmit last_emit{};
  uint32_t* last_ins_pos = nullptr;
  uint32_t last_n_code = 0;
  size_t last_val_bytes = 0;
  Job* job = nullptr;  // page-locked pools, kept across calls
  // the arena of the lane's last block stays resident so that its hashing can be re-run for measurement
  bool has_last = false;
  ArenaView last_view{};
  std::vector<uint32_t> last_level_start;
  uint32_t last_n_msgs = 0;
};

// Counting semaphore: how many lanes may have their witness upload + parse in flight at once.  All lanes of a
// batch start together; letting every one of them share the copy engine and the SMs makes all of them finish
// their parse late and at the same time, after which all host threads shape their tries at once with the GPU
// idle.  Admitting a few at a time staggers the lanes, so the parse, the host shaping and the IR dump of
// different blocks overlap.
struct Slots {
  std::mutex mu;
  std::condition_variable cv;
  int free_slots;
  explicit Slots(int n) : free_slots(n) {}
  void acquire() {
    std::unique_lock<std::mutex> g(mu);
    cv.wait(g, [&] { return free_slots > 0; });
    free_slots--;
  }
  void release() {
    {
      std::lock_guard<std::mutex> g(mu);
      free_slots++;
    }
    cv.notify_one();
  }
};
struct SlotGuard {
  Slots* s;
  explicit SlotGuard(Slots* s_) : s(s_) {
    if (s) s->acquire();
  }
  void done() {
    if (s) s->release();
    s = nullptr;
  }
  ~SlotGuard() { done(); }
};
int parse_slots() {
  const char* e = getenv("PPD_PARSE_SLOTS");
  int v = e ? atoi(e) : 4;
  return v < 1 ? 1 : v;
}

struct ppd_ctx {
  int device = 0;
  Slots parse_slots_sem{parse_slots()};
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  std::mutex err_mu;
  ppd_stats stats{};
  std::vector<Lane*> lanes;
  size_t last_lanes_used = 0;  // lanes holding a resident block of the last decode call
  // HBM buffers of the non-block entry points, grown on demand and reused across calls
  DevBuf d_keys, d_vals, d_ref, d_ref_len, d_counters;
  DevBuf d_msg, d_msg_off, d_digest;
  DevBuf d_build[12];
};

namespace {

void stats_reset(ppd_ctx* c) { c->stats = ppd_stats{}; }

// wait for everything queued on the lane's stream without spinning (lanes may outnumber cores)
void lane_sync(Lane* l) {
  CUDA_OK(cudaEventRecord(l->ev_sync, l->st));
  CUDA_OK(cudaEventSynchronize(l->ev_sync));
}
// The same by polling, for the short waits a lane makes while it holds a parse slot (at most PPD_PARSE_SLOTS
// threads poll at a time): a sleeping thread is woken late when every core is busy shaping other blocks, and
// everything queued behind the slot waits with it.
void lane_sync_poll(Lane* l) {
  CUDA_OK(cudaEventRecord(l->ev_sync, l->st));
  for (unsigned spins = 0;; spins++) {
    cudaError_t e = cudaEventQuery(l->ev_sync);
    if (e == cudaSuccess) return;
    if (e != cudaErrorNotReady) CUDA_OK(e);
    if (spins > 200) std::this_thread::yield();
  }
}

// ============================================================================================
// Phase I: batched Keccak-256 of byte strings (addresses, slots, code)
// ============================================================================================
struct KeyHasher {
  PVec<uint8_t> data;
  std::vector<uint64_t> off{0};
  PVec<H256> digest;
  std::vector<uint64_t> lens;
  PVec<uint64_t> se;  // (begin, end) pairs
  KeyHasher() {
    data.alloc_fn = pinned_alloc, data.free_fn = pinned_free;
    digest.alloc_fn = pinned_alloc, digest.free_fn = pinned_free;
    se.alloc_fn = pinned_alloc, se.free_fn = pinned_free;
  }
  void reset() {
    data.clear(), digest.clear(), lens.clear(), se.clear();
    off.assign(1, 0);
  }
  uint32_t add(const uint8_t* p, size_t n) {
    size_t at = data.size(), padded = (n + 3) & ~(size_t)3;
    data.resize(at + padded);  // every message starts 4-byte aligned
    memcpy(data.data() + at, p, n);
    memset(data.data() + at + n, 0, padded - n);
    uint32_t idx = (uint32_t)lens.size();
    lens.push_back(n);
    off.push_back(data.size());
    return idx;
  }
  void run(Lane* c) {
    size_t n = lens.size();
    digest.resize(n);
    if (!n) return;
#ifdef PPD_HOSTPROF
    for (size_t i = 0; i < n; i++) hostprof::keccak256(data.data() + off[i], lens[i], digest[i].b);
    return;
#endif
    // messages are padded to 4-byte boundaries, so pass explicit (begin, end) pairs
    se.resize(2 * n);
    for (size_t i = 0; i < n; i++) se[2 * i] = off[i], se[2 * i + 1] = off[i] + lens[i];
    c->d_msg.reserve(data.size() + 16);
    c->d_msg_off.reserve(se.size() * 8);
    c->d_digest.reserve(n * 32);
    CUDA_OK(cudaMemcpyAsync(c->d_msg.p, data.data(), data.size(), cudaMemcpyHostToDevice, c->st));
    CUDA_OK(cudaMemcpyAsync(c->d_msg_off.p, se.data(), se.size() * 8, cudaMemcpyHostToDevice, c->st));
    launch_keccak256_ranges(c->d_msg.as<uint8_t>(), c->d_msg_off.as<uint64_t>(), (uint32_t)n, c->d_digest.as<uint8_t>(), c->st);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(digest.data(), c->d_digest.p, n * 32, cudaMemcpyDeviceToHost, c->st));
    lane_sync(c);
    c->stats.key_hashes += n;
    for (size_t i = 0; i < n; i++) c->stats.key_permutations += lens[i] / 136 + 1;
    c->stats.h2d_bytes += (double)(data.size() + se.size() * 8);
    c->stats.d2h_bytes += (double)(n * 32);
    c->stats.kernel_launches += 1;
  }
};

// ============================================================================================
// Compact witness -> instruction tree (compact_prestate_processing.rs:683-875, 387-668)
// ============================================================================================
struct Span {
  const uint8_t* p = nullptr;
  uint32_t n = 0;
};
// One instruction of the witness, 20 bytes.  Operands are not copied: `pos` points at the first operand
// byte and the (already validated) CBOR heads are re-read when the instruction is used.
struct WNode {
  uint32_t pos;
  uint8_t op, flags;     // flags: the account leaf's flag byte (bit0 code, bit1 storage, bit2 nonce, bit3 balance)
  uint16_t unused = 0;
  // tree links filled by the stack machine
  int32_t first_child;   // branch: first child (ascending nibble order); extension: child; account leaf: storage node
  int32_t next_sibling;  // next child of the same branch
  uint32_t aux;          // branch: the 32-bit mask; account leaf: code node (or ~0)
};

struct WCursor {
  const uint8_t* p;
  size_t n, pos = 0;
  uint8_t read_byte() {
    if (pos >= n) fail(PPD_ERR_UNEXPECTED_END_OF_STREAM, "read_byte at end of stream");
    return p[pos++];
  }
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
  Span cbor_bytes(int err) {
    uint8_t major;
    uint64_t len;
    if (!cbor_head(major, len) || major != 2 || len > n - pos) fail(err, "bad CBOR byte string");
    Span s{p + pos, (uint32_t)len};
    pos += len;
    return s;
  }
  uint64_t cbor_uint(uint64_t max) {
    uint8_t major;
    uint64_t v;
    if (!cbor_head(major, v) || major != 0 || v > max) fail(PPD_ERR_INVALID_BYTES_FOR_TYPE, "bad CBOR unsigned integer");
    return v;
  }
};

// key_bytes_to_nibbles (compact_prestate_processing.rs:1338-1390); appends to `out`, returns count
uint32_t compact_key_nibbles(Span k, uint8_t* out) {
  if (k.n == 0) return 0;
  uint32_t c = 0;
  if (k.n == 1) {
    out[c++] = k.p[0] & 15;
    return c;
  }
  bool odd = k.p[0] & 1;
  uint32_t m = k.n - 1;
  if (2 * m > 64 + 1) fail(PPD_ERR_KEY_ERROR, "compact key longer than 64 nibbles");
  for (uint32_t i = 0; i + 1 < m; i++) {
    out[c++] = k.p[1 + i] >> 4;
    out[c++] = k.p[1 + i] & 15;
  }
  out[c++] = k.p[m] >> 4;
  if (!odd) out[c++] = k.p[m] & 15;
  return c;
}

// key_bytes_to_nibbles runs while the instructions are read (compact_prestate_processing.rs:787-835), so
// a key of more than 64 nibbles is reported in stream order, before any later parse error
void check_key_length(Span k) {
  if (k.n >= 2 && 2 * (k.n - 1) > 64 + 1) fail(PPD_ERR_KEY_ERROR, "compact key longer than 64 nibbles");
}

struct Witness {
  const uint8_t* bytes = nullptr;
  size_t len = 0;
  uint8_t version = 0;
  std::vector<WNode> ins;
  int32_t root = -1;  // -1: header only

  // operand views (the stream was validated by parse_witness)
  Span key(const WNode& x) const {
    WCursor c{bytes, len, x.pos};
    return c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR);
  }
  Span leaf_value(const WNode& x) const {
    WCursor c{bytes, len, x.pos};
    c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR);
    return c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR);
  }
  Span code(const WNode& x) const { return key(x); }
  const uint8_t* hash(const WNode& x) const { return bytes + x.pos; }
  void account(const WNode& x, Span& key_out, uint64_t& nonce, Span& balance) const {
    WCursor c{bytes, len, x.pos};
    key_out = c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR);
    c.pos++;  // flags
    nonce = (x.flags & 4) ? c.cbor_uint(~0ull) : 0;
    balance = (x.flags & 8) ? c.cbor_bytes(PPD_ERR_INVALID_BYTE_VECTOR) : Span{};
  }
};

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
          if (bal.n > 32) fail(PPD_ERR_INVALID_BYTE_VECTOR, "balance wider than 256 bits");
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
struct TrieItem {
  uint32_t koff, klen;
  uint8_t kind;  // 0 value leaf, 1 account leaf, 2 hashed-out subtree
  uint32_t a1, a2;
};

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

// ============================================================================================
// One block being decoded
// ============================================================================================
struct TraceV {
  const uint8_t* addr;
  uint8_t flags;
  const uint8_t *balance = nullptr, *nonce = nullptr;
  uint32_t n_reads = 0, n_writes = 0;
  const uint8_t *reads = nullptr, *writes = nullptr;
  const uint8_t* code_read = nullptr;
  Span code_write;
  // message indices into the key hasher
  uint32_t m_addr = 0, m_reads = 0, m_writes_full = 0, m_writes_min = 0, m_code = 0;
};
struct TxnV {
  std::vector<TraceV> traces;
  Span byte_code, new_txn_node, new_receipt_node;
  uint64_t gas_used = 0;
};
struct FlatReader {
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
  const uint8_t* raw(size_t k) {
    need(k);
    const uint8_t* r = p + pos;
    pos += k;
    return r;
  }
  Span bytes() {
    uint32_t k = u32();
    return Span{raw(k), k};
  }
};

struct IrPlan {
  uint64_t txn_before = 0, gas_before = 0, gas_after = 0;
  bool has_signed_txn = false;
  Span signed_txn;
  bool has_withdrawals = false;
  uint32_t state_sub = NODE_EMPTY, txn_sub = NODE_EMPTY, receipt_sub = NODE_EMPTY;  // roots of the tries the subsets are cut from
  std::vector<std::pair<H256, uint32_t>> storage_subs;
  std::vector<uint32_t> touched;
  std::vector<HostArena::MarkItem> items;  // marking walks left to the device (launch_mark_walk); materialize_touched() runs them on the host
  uint32_t root_state = 0, root_txn = 0, root_receipt = 0;  // NK_ROOT nodes
  std::map<H256, Span> code;
};

struct BlockJob {
  // input views
  Span compact;
  std::vector<TxnV> txns;
  std::unordered_map<H256, Span, H256Hasher> resolved_code;
  std::vector<std::pair<const uint8_t*, const uint8_t*>> withdrawals;
  std::vector<uint32_t> m_withdrawal_addr;
  const uint8_t* checkpoint = nullptr;
  Span b_meta, b_hashes;
  // decoded witness
  Witness wit;
  std::vector<uint32_t> m_inline_code;  // per instruction: message index of an inline Code node, or ~0
  std::map<H256, Span> pre_code;        // WitnessOutput.code
  // tries
  uint32_t state_root = NODE_EMPTY;
  H256Map storage;  // hashed address -> root node
  struct PreAccount {
    H256 haddr;
    uint32_t rec;
    bool storage_nonempty;
  };
  std::vector<PreAccount> pre_accounts;
  H256Map pre_with_storage;  // accounts whose storage root != EMPTY_TRIE_HASH -> record
  H256Map acct_rec;          // hashed address -> the account's current record (what state.get() + rlp::decode gives, decoding.rs:251-254)
  FlatMapU32 root_of;                                               // trie root node -> its NK_ROOT node
  std::unordered_map<int32_t, uint32_t> storage_root_of_instr;      // account leaf instruction -> root of its witnessed storage trie
  bool have_empty_form = false;                                     // a witnessed storage trie whose root is EMPTY_TRIE_HASH
  bool pre_image_on_gpu = false;                                    // gpu_pre_image built the pre-image tries
  uint32_t empty_form = NODE_EMPTY;
  std::vector<IrPlan> irs;
  int status = PPD_OK;
  std::string err;
};

void read_flat_block(const uint8_t* p, size_t n, BlockJob& b) {
  FlatReader r{p, n};
  if (r.u32() != PPD_FLAT_BLOCK_MAGIC || r.u32() != 1) fail(PPD_ERR_BAD_FLAT_INPUT, "bad magic/version");
  if (r.u32() != 0) fail(PPD_PANIC_UNIMPLEMENTED_PRE_IMAGE, "only Combined{compact} pre-images are implemented by the reference");
  b.compact = r.bytes();
  uint32_t nt = r.u32();
  b.txns.resize(nt);
  for (uint32_t t = 0; t < nt; t++) {
    TxnV& tx = b.txns[t];
    uint32_t ntr = r.u32();
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
  for (uint32_t i = 0; i < nc; i++) {
    H256 h;
    memcpy(h.b, r.raw(32), 32);
    b.resolved_code[h] = r.bytes();
  }
  uint32_t nw = r.u32();
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
struct RlpItem {
  bool is_list;
  const uint8_t* payload;
  size_t payload_len, total_len;
};
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

// ============================================================================================
// Job = a batch of blocks sharing one arena, one key-hash launch and one sweep
// ============================================================================================
struct Job {
  HostArena A;
  KeyHasher kh;
  std::vector<BlockJob> blocks;
  PVec<uint8_t> ref, ref_len;  // after the sweep
  PVec<uint32_t> order;
  PVec<uint32_t> plan;  // IR dump plan (inputs, then the outputs read back)
  PVec<uint8_t> out_stage;  // page-locked landing buffer of the serialised IrDump
  bool refs_on_host = false;
  // When the pre-image was built on the GPU (gpu_pre_image) the leading part of every pool is already in
  // the lane's device buffers: the sweep uploads only what the txn loop appended.  The value and hash
  // pools of that part are not copied to the host unless a host-side dump needs them (fetch_pools).
  struct Resident {
    size_t nodes = 0, keys = 0, vals = 0, hashes = 0, children = 0, accounts = 0;
  } dev;
  bool pools_on_host = true;
  PVec<uint32_t> acct_list, code_list;
  PVec<uint8_t> wit_stage;  // page-locked staging of a pageable witness
  std::vector<HostArena::MarkItem> mark_items;  // scratch of the txn loop
  std::vector<HostArena::BatchItem> batch_items;
  std::vector<uint32_t> haddr_keys, haddr_leaves;
  bool device_marks = false;  // this block's subset marking walks run on the device (decode_one decides)
  PVec<H256> code_digest;
  std::vector<uint32_t> stamp;
  uint32_t serial = 0;
  Job() {
    plan.alloc_fn = pinned_alloc, plan.free_fn = pinned_free;
    out_stage.alloc_fn = pinned_alloc, out_stage.free_fn = pinned_free;
    A.set_allocator(pinned_alloc, pinned_free);
    ref.alloc_fn = ref_len.alloc_fn = pinned_alloc, ref.free_fn = ref_len.free_fn = pinned_free;
    order.alloc_fn = pinned_alloc, order.free_fn = pinned_free;
    A.level.alloc_fn = pinned_alloc, A.level.free_fn = pinned_free;
    acct_list.alloc_fn = code_list.alloc_fn = pinned_alloc, acct_list.free_fn = code_list.free_fn = pinned_free;
    code_digest.alloc_fn = pinned_alloc, code_digest.free_fn = pinned_free;
    wit_stage.alloc_fn = pinned_alloc, wit_stage.free_fn = pinned_free;
  }
  void reset(size_t n_blocks) {
    dev = Resident{};
    pools_on_host = true;
    device_marks = false;
    A.clear();
    kh.reset();
    blocks.clear();
    blocks.resize(n_blocks);
    ref.clear(), ref_len.clear(), order.clear();
    serial = 0;
  }
};
void job_delete(Job* j) { delete j; }
Job& job_of(Lane* l, size_t n_blocks) {
  if (!l->job) l->job = new Job();
  l->job->reset(n_blocks);
  return *l->job;
}

Lane* lane_of(ppd_ctx* c, size_t w) {
  while (c->lanes.size() <= w) {
    std::unique_ptr<Lane> l(new Lane());
#ifndef PPD_HOSTPROF
    CUDA_OK(cudaStreamCreateWithFlags(&l->st, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreate(&l->ev0));
    CUDA_OK(cudaEventCreate(&l->ev1));
    CUDA_OK(cudaEventCreateWithFlags(&l->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
#endif
    c->lanes.push_back(l.release());
  }
  return c->lanes[w];
}
void lane_delete(Lane* l) {
#ifndef PPD_HOSTPROF
  DevBuf* bufs[] = {&l->d_nodes, &l->d_order,  &l->d_keys,     &l->d_vals, &l->d_hashes,  &l->d_children, &l->d_accounts,
                    &l->d_ref,   &l->d_ref_len, &l->d_counters, &l->d_msg,  &l->d_msg_off, &l->d_digest,
                    &l->d_plan,  &l->d_out,     &l->d_wit,      &l->d_pa,   &l->d_pb,      &l->d_pc,
                    &l->d_level, &l->d_okeys,   &l->d_obins};
  for (DevBuf* b : bufs) b->release();
  if (l->h_parse) pinned_free(l->h_parse);
  if (l->ev0) cudaEventDestroy(l->ev0);
  if (l->ev1) cudaEventDestroy(l->ev1);
  if (l->ev_sync) cudaEventDestroy(l->ev_sync);
  if (l->st) cudaStreamDestroy(l->st);
#endif
  if (l->job) job_delete(l->job);
  delete l;
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
  b.pre_accounts.push_back({haddr, r, nonempty});
  if (J.device_marks) b.acct_rec[haddr] = r;
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
      b.acct_rec.erase(b.pre_accounts[i].haddr);
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
}

// ---- step 2 on the GPU (ppd_parse.cu): witness bytes -> instruction list -> tree links -> arena ------
// Three device phases with one small read-back each (instruction count; flags and pool sizes; the
// structural half of the arena).  The host keeps only what the txn loop walks (node records, keys,
// child lists, account records, levels); leaf values and the hashed-out subtrees stay in HBM.
// Returns false when the witness is not a well-formed canonical one: the host builder then takes it
// from the start and reports the reference's error, if any.
void verify_gpu_pre_image(Lane* L, Job& J, BlockJob& b);
void fetch_pools(Lane* c, Job& J);
bool gpu_parse_enabled() {
#ifdef PPD_HOSTPROF
  return false;
#else
  return getenv("PPD_HOST_PARSE") == nullptr;  // read per call: the tests compare both builders in one process
#endif
}

struct Carve {
  uint8_t* base;
  size_t off = 0;
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
};

bool gpu_pre_image(Lane* L, Job& J, BlockJob& b, bool check_version = true, Slots* slots = nullptr) {
#ifdef PPD_HOSTPROF
  return false;
#else
  const uint8_t* w = b.compact.p;
  const size_t n = b.compact.n;
  if (n < 2 || n >= 0xfff00000ull) return false;
  // The three phases cost three read-backs and about 45 launches whatever the size: below a few hundred KiB
  // that latency exceeds what the host builder needs for the whole witness (config 4: 1024 blocks of 100 KB
  // each decode at 7.1 k blocks/s with the host builder, 3.3 k with this one), at config-2 size (36 MB) it
  // is 16x faster.  PPD_GPU_PARSE_MIN_BYTES moves the switch (the tests set it to 0).
  {
    const char* e = getenv("PPD_GPU_PARSE_MIN_BYTES");
    const size_t min_bytes = e ? (size_t)atoll(e) : (size_t)512 << 10;
    if (n < min_bytes) return false;
  }
  HostArena& A = J.A;
  cudaStream_t st = L->st;
  if (!L->h_parse) {
    L->h_parse = (uint32_t*)pinned_alloc(4 * PARSE_R_WORDS);
    if (!L->h_parse) fail(PPD_ERR_BAD_ARGUMENT, "out of page-locked memory");
  }
  uint32_t* hr = L->h_parse;
  static const bool timing = getenv("PPD_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[ppd]   %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  SlotGuard slot(slots);
  auto sync_in_slot = [&] { slots ? lane_sync_poll(L) : lane_sync(L); };
  lap("p:slot-wait");
  // ---- phase A: instruction boundaries ----
  L->d_wit.reserve(n + 64);
  {
    // A page-locked caller buffer (ppd_alloc_pinned, cudaHostRegister) is read by the copy engine directly.
    // A pageable one is staged through the lane's page-locked buffer in chunks: concurrent pageable
    // cudaMemcpyAsync calls serialise inside the driver, a plain memcpy per lane does not.
    cudaPointerAttributes at{};
    bool pinned = cudaPointerGetAttributes(&at, w) == cudaSuccess && at.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();
    if (pinned) {
      CUDA_OK(cudaMemcpyAsync(L->d_wit.p, w, n, cudaMemcpyHostToDevice, st));
    } else {
      J.wit_stage.resize(n);
      const size_t CH = 4u << 20;
      for (size_t at0 = 0; at0 < n; at0 += CH) {
        size_t len = std::min(CH, n - at0);
        memcpy(J.wit_stage.data() + at0, w + at0, len);
        CUDA_OK(cudaMemcpyAsync(L->d_wit.as<uint8_t>() + at0, J.wit_stage.data() + at0, len, cudaMemcpyHostToDevice, st));
      }
    }
  }
  CUDA_OK(cudaMemsetAsync(L->d_wit.as<uint8_t>() + n, 0, 64, st));
  L->stats.h2d_bytes += (double)n;
  ParseBounds B{};
  B.wit = L->d_wit.as<uint8_t>();
  B.n = (uint32_t)n;
  B.n_tiles = (uint32_t)((n + PARSE_TILE - 1) / PARSE_TILE);
  B.group_tiles = 8;
  while (B.group_tiles < 1024 && (uint64_t)B.group_tiles * B.group_tiles < B.n_tiles) B.group_tiles *= 2;
  B.n_groups = (B.n_tiles + B.group_tiles - 1) / B.group_tiles;
  auto layout_a = [&](Carve& c) {
    B.result = c.take<uint32_t>(PARSE_R_WORDS);
    B.exit1 = c.take<uint32_t>((size_t)B.n_tiles * PARSE_TILE);  // whole tiles: tile_exit_kernel stores 128-bit rows
    B.step1 = c.take<uint16_t>((size_t)B.n_tiles * PARSE_TILE);
    B.exit2 = c.take<uint32_t>((size_t)B.n_groups * PARSE_TILE);
    B.group_entry = c.take<uint32_t>(B.n_groups);
    B.tile_entry = c.take<uint32_t>(B.n_tiles);
    B.bitmap = c.take<uint32_t>((size_t)B.n_tiles * (PARSE_TILE / 32));
    B.tile_count = c.take<uint32_t>(B.n_tiles + 1);
    B.tile_base = c.take<uint32_t>(B.n_tiles + 1);
    B.scan_tmp = c.take<uint32_t>(parse_scan_tmp_words(B.n_tiles + 1, 1));
  };
  {
    Carve sz{nullptr};
    layout_a(sz);
    L->d_pa.reserve(sz.off + 256);
    Carve c{L->d_pa.as<uint8_t>()};
    layout_a(c);
  }
  auto phase_ms = [&] {
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, L->ev0, L->ev1));
    L->stats.parse_gpu_ms += ms;
  };
  CUDA_OK(cudaMemsetAsync(B.result, 0, 4 * PARSE_R_WORDS, st));
  CUDA_OK(cudaEventRecord(L->ev0, st));
  L->stats.kernel_launches += launch_parse_bounds(B, st);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(L->ev1, st));
  CUDA_OK(cudaMemcpyAsync(hr, B.result, 4 * PARSE_R_WORDS, cudaMemcpyDeviceToHost, st));
  sync_in_slot();
  phase_ms();
  lap("p:upload+A");
  if (hr[PARSE_R_END] != (uint32_t)n) return false;  // a parse error: the host parser reports it
  const uint32_t n_ins = hr[PARSE_R_NINS];
  if (n_ins == 0 || n_ins > n) return false;
  // phase B keeps about 100 bytes per instruction: a stream of one- and two-byte instructions (no real witness:
  // a node that can be a child is at least an opcode and a CBOR head) would ask for more memory than the witness
  // justifies; the host builder takes it
  if ((uint64_t)n_ins * 4 > (uint64_t)n + 256) return false;
  // ---- phase B: tree links, depths, sizes ----
  ParseTree T{};
  T.wit = B.wit, T.n = B.n, T.n_ins = n_ins, T.result = B.result;
  T.cnt_stride = ((size_t)n_ins + 1 + 3) & ~(size_t)3;
  uint32_t* ins_pos = nullptr;
  auto layout_b = [&](Carve& c) {
    const size_t n1 = (size_t)n_ins + 1;
    const size_t n_m1 = (n1 + 63) / 64, n_m2 = (n_m1 + 63) / 64, n_m3 = (n_m2 + 63) / 64;
    ins_pos = c.take<uint32_t>(n_ins);
    T.meta = c.take<uint32_t>(n_ins);
    T.knib = c.take<uint8_t>(n_ins);
    T.delta = c.take<uint32_t>(n1);
    T.hb = c.take<uint32_t>(n1);
    T.h16 = c.take<int16_t>(n1);
    T.m1 = c.take<int16_t>(n_m1);
    T.m2 = c.take<int16_t>(n_m2);
    T.m3 = c.take<int16_t>(n_m3);
    T.parent = c.take<uint32_t>(n_ins);
    T.info = c.take<uint32_t>(n_ins);
    T.aux0 = c.take<uint32_t>(n_ins);
    T.pending = c.take<uint32_t>(n_ins);
    T.lvlmax = c.take<uint32_t>(n_ins);
    T.cnt = c.take<uint32_t>(PARSE_N_CNT * T.cnt_stride);
    T.scn = c.take<uint32_t>(PARSE_N_CNT * T.cnt_stride);
    T.scan_tmp = c.take<uint32_t>(parse_scan_tmp_words(n1, PARSE_N_CNT));
  };
  {
    Carve sz{nullptr};
    layout_b(sz);
    L->d_pb.reserve(sz.off + 256);
    Carve c{L->d_pb.as<uint8_t>()};
    layout_b(c);
  }
  T.ins_pos = ins_pos;
  CUDA_OK(cudaEventRecord(L->ev0, st));
  launch_parse_scatter(B, ins_pos, st);
  L->stats.kernel_launches += 1 + launch_parse_tree(T, st);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(L->ev1, st));
  CUDA_OK(cudaMemcpyAsync(hr, B.result, 4 * PARSE_R_WORDS, cudaMemcpyDeviceToHost, st));
  sync_in_slot();
  phase_ms();
  lap("p:B");
  if (hr[PARSE_R_FLAG] != 0 || hr[PARSE_R_HEIGHT] != 1) return false;
  if (check_version && w[0] != 1) fail(PPD_PANIC_INCOMPATIBLE_HEADER_VERSION, "compact header version is not 1");
  const uint32_t* tot = hr + PARSE_R_TOTALS;
  const size_t n_nodes = tot[PARSE_C_NODE], n_hash = tot[PARSE_C_HASH], key_bytes = tot[PARSE_C_KEY], val_bytes = tot[PARSE_C_VAL],
               n_child = tot[PARSE_C_CHILD], n_acct = tot[PARSE_C_ACCT], n_code = tot[PARSE_C_CODE];
  const uint32_t root_ins = hr[PARSE_R_ROOT];
  if (root_ins >= n_ins || n_hash >= HASH_ID_END - HASH_ID_BASE) return false;
  // ---- phase C: emit the arena into the lane's buffers ----
  ParseEmit E{};
  E.T = T;
  uint16_t* d_level = nullptr;
  uint8_t* d_code_digest = nullptr;
  auto layout_c = [&](Carve& c) {
    E.acct_list = c.take<uint32_t>(5 * n_acct + 1);
    E.code_se = c.take<uint64_t>(2 * n_code + 1);
    E.code_list = c.take<uint32_t>(2 * n_code + 1);
    d_code_digest = c.take<uint8_t>(32 * n_code + 32);
  };
  {
    Carve sz{nullptr};
    layout_c(sz);
    L->d_pc.reserve(sz.off + 256);
    Carve c{L->d_pc.as<uint8_t>()};
    layout_c(c);
  }
  // room for what the txn loop appends, so that the sweep does not have to move the resident part
  L->d_nodes.reserve(16 * (n_nodes + n_nodes / 2) + 4096);
  L->d_level.reserve(2 * (n_nodes + n_nodes / 2) + 4096);
  d_level = L->d_level.as<uint16_t>();
  L->d_keys.reserve(2 * key_bytes + 65536);
  L->d_vals.reserve(2 * val_bytes + 65536);
  L->d_hashes.reserve(32 * n_hash + 32);
  L->d_children.reserve(4 * (n_child + n_child / 2) + 4096);
  L->d_accounts.reserve(sizeof(AccountRec) * (2 * n_acct + 64));
  E.nodes = L->d_nodes.as<NodeRec>();
  E.level = d_level;
  E.key_pool = L->d_keys.as<uint8_t>();
  E.val_pool = L->d_vals.as<uint8_t>();
  E.hash_pool = L->d_hashes.as<uint8_t>();
  E.child_pool = L->d_children.as<uint32_t>();
  E.accounts = L->d_accounts.as<AccountRec>();
  E.code_digest = d_code_digest;
  CUDA_OK(cudaEventRecord(L->ev0, st));
  if (n_code) {
    launch_parse_code_list(E, st);
    launch_keccak256_ranges(B.wit, E.code_se, (uint32_t)n_code, d_code_digest, st);
    L->stats.kernel_launches += 2;
    L->stats.key_hashes += n_code;
  }
  if (val_bytes) CUDA_OK(cudaMemsetAsync(E.val_pool, 0, val_bytes, st));
  launch_parse_emit(E, st);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(L->ev1, st));
  L->stats.kernel_launches += 2;
  slot.done();  // the next lane may start its upload while this one's emit kernels and download run
  A.nodes.resize(n_nodes), A.level.resize(n_nodes), A.key_pool.resize(key_bytes), A.child_pool.resize(n_child), A.accounts.resize(n_acct);
  A.val_pool.resize(val_bytes), A.hash_pool.resize(32 * n_hash);  // contents stay on the device (fetch_pools)
  J.acct_list.resize(5 * n_acct), J.code_list.resize(2 * n_code), J.code_digest.resize(n_code);
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    L->stats.d2h_bytes += (double)bytes;
  };
  down(A.nodes.data(), E.nodes, 16 * n_nodes);
  down(A.level.data(), d_level, 2 * n_nodes);
  down(A.key_pool.data(), E.key_pool, key_bytes);
  down(A.child_pool.data(), E.child_pool, 4 * n_child);
  down(A.accounts.data(), E.accounts, sizeof(AccountRec) * n_acct);
  down(J.acct_list.data(), E.acct_list, 20 * n_acct);
  down(J.code_list.data(), E.code_list, 8 * n_code);
  down(J.code_digest.data(), d_code_digest, 32 * n_code);
  down(hr, B.result, 4 * PARSE_R_WORDS);
  lane_sync(L);
  phase_ms();
  lap("p:C+download");
  for (size_t k = 0; k < n_code; k++) {
    L->stats.key_permutations += J.code_list[2 * k + 1] / 136 + 1;
    b.pre_code[J.code_digest[k]] = Span{w + J.code_list[2 * k], J.code_list[2 * k + 1]};
  }
  J.dev.nodes = n_nodes, J.dev.keys = key_bytes, J.dev.vals = val_bytes, J.dev.hashes = 32 * n_hash, J.dev.children = n_child, J.dev.accounts = n_acct;
  J.pools_on_host = false;
  // ---- the block's per-account tables (compact_to_partial_trie.rs:167-190), as make_account_record builds them ----
  b.wit.version = w[0];
  b.state_root = hr[PARSE_R_ROOT_ID];
  b.storage.reserve(n_acct), b.pre_accounts.reserve(n_acct), b.root_of.reserve(2 * n_acct + 1024);
  if (J.device_marks) b.acct_rec.reserve(n_acct + n_acct / 4);
  b.have_empty_form = false, b.empty_form = NODE_EMPTY;
  const uint32_t* al = J.acct_list.data();
  for (size_t a = 0; a < n_acct; a++)
    if ((al[5 * a + 3] & 1u) && !(al[5 * a + 3] & 2u)) b.have_empty_form = true, b.empty_form = al[5 * a + 1];
  for (size_t a = 0; a < n_acct; a++) {
    const uint32_t leaf = al[5 * a], flags = al[5 * a + 3];
    const bool nonempty = flags & 2u;
    bool has_trie = flags & 1u;
    uint32_t sroot = al[5 * a + 1];
    if (!nonempty) has_trie = b.have_empty_form, sroot = b.empty_form;
    const NodeRec& nr = A.nodes[leaf];
    const uint32_t klen = ((nr.w0 >> 8) & 0xff) + ((nr.w0 >> 16) & 0xff);
    H256 haddr;
    if (klen == 64) {
      memcpy(haddr.b, A.key_pool.data() + nr.a0, 32);
    } else {  // utils.rs:49-59: the nibbles right-aligned in 32 bytes
      memset(haddr.b, 0, 32);
      for (uint32_t k = 0; k < klen; k++) {
        uint32_t posn = 64 - klen + k, nib = A.key_nib(nr.a0, k);
        haddr.b[posn >> 1] |= (uint8_t)((posn & 1) ? nib : (nib << 4));
      }
    }
    if (has_trie) b.storage[haddr] = sroot;
    b.pre_accounts.push_back({haddr, (uint32_t)a, nonempty});
    if (J.device_marks) b.acct_rec.insert({haddr, (uint32_t)a});
    if (nonempty) {
      b.pre_with_storage[haddr] = (uint32_t)a;
      b.root_of.put(al[5 * a + 1], al[5 * a + 2]);
    }
  }
  lap("p:tables");
  b.pre_image_on_gpu = true;
  L->has_last_parse = true, L->last_bounds = B, L->last_emit = E, L->last_ins_pos = ins_pos, L->last_n_code = (uint32_t)n_code, L->last_val_bytes = val_bytes;
  L->stats.witnesses_on_gpu += 1, L->stats.witness_instructions += n_ins, L->stats.witness_bytes += n;
  if (getenv("PPD_VERIFY_GPU_PARSE")) verify_gpu_pre_image(L, J, b);
  return true;
#endif
}

// ---- PPD_VERIFY_GPU_PARSE: the GPU-built pre-image against the host builder's, node by node ----------
struct TrieCmp {
  const HostArena &X, &Y;
  std::string why;
  bool no(const char* what, uint32_t x, uint32_t y) {
    if (why.empty()) why = std::string(what) + " (gpu node " + std::to_string(x) + ", host node " + std::to_string(y) + ")";
    return false;
  }
  bool nibs_eq(uint32_t x, uint32_t y) {
    if (X.nstart(x) != Y.nstart(y) || X.nlen(x) != Y.nlen(y)) return false;
    for (uint32_t k = 0; k < X.nstart(x) + X.nlen(x); k++)  // the whole key up to the end of the node's range
      if (X.key_nib(X.nodes[x].a0, k) != Y.key_nib(Y.nodes[y].a0, k)) return false;
    return true;
  }
  bool eq(uint32_t x, uint32_t y) {
    if (x == NODE_EMPTY || y == NODE_EMPTY) return x == y ? true : no("empty vs non-empty", x, y);
    uint32_t kx = X.kind(x), ky = Y.kind(y);
    if (kx != ky) return no("node kinds differ", x, y);
    if (kx == NK_HASH) return memcmp(X.hash_of(x), Y.hash_of(y), 32) == 0 ? true : no("hashed-out nodes differ", x, y);
    if (X.lvl(x) != Y.lvl(y)) return no("levels differ", x, y);
    const NodeRec &a = X.nodes[x], &b = Y.nodes[y];
    switch (kx) {
      case NK_LEAF:
        if (!nibs_eq(x, y)) return no("leaf keys differ", x, y);
        if (a.a2 != b.a2 || memcmp(X.val_pool.data() + a.a1, Y.val_pool.data() + b.a1, a.a2) != 0) return no("leaf values differ", x, y);
        return true;
      case NK_LEAF_ACCOUNT: {
        if (!nibs_eq(x, y)) return no("account keys differ", x, y);
        const AccountRec &ra = X.accounts[a.a1], &rb = Y.accounts[b.a1];
        if (memcmp(&ra, &rb, 128) != 0) return no("account records differ", x, y);
        if ((ra.storage_src == NODE_EMPTY) != (rb.storage_src == NODE_EMPTY)) return no("account storage sources differ", x, y);
        return ra.storage_src == NODE_EMPTY ? true : eq(ra.storage_src, rb.storage_src);
      }
      case NK_EXT:
        if (!nibs_eq(x, y)) return no("extension keys differ", x, y);
        return eq(a.a1, b.a1);
      case NK_ROOT:
        return eq(a.a1, b.a1);
      case NK_BRANCH: {
        if ((a.a1 & 0xffff) != (b.a1 & 0xffff)) return no("branch masks differ", x, y);
        uint32_t k = (uint32_t)__builtin_popcount(a.a1 & 0xffff);
        for (uint32_t i = 0; i < k; i++)
          if (!eq(X.child_pool[a.a0 + i], Y.child_pool[b.a0 + i])) return false;
        return true;
      }
    }
    return no("unknown node kind", x, y);
  }
};

void verify_gpu_pre_image(Lane* L, Job& J, BlockJob& b) {
  fetch_pools(L, J);
  std::unique_ptr<Job> J2(new Job());
  J2->reset(1);
  BlockJob& b2 = J2->blocks[0];
  b2.compact = b.compact;
  parse_witness(b.compact.p, b.compact.n, b2.wit);
  collect_witness_messages(*J2, b2);
  J2->kh.run(L);
  build_pre_image(*J2, b2);
  auto bad = [&](const std::string& m) { throw Fail{PPD_ERR_CUDA, "GPU pre-image differs from the host builder's: " + m}; };
  TrieCmp cmp{J.A, J2->A};
  if (!cmp.eq(b.state_root, b2.state_root)) bad("state trie: " + cmp.why);
  if (b.storage.size() != b2.storage.size()) bad("storage map sizes " + std::to_string(b.storage.size()) + " / " + std::to_string(b2.storage.size()));
  b2.storage.for_each([&](const H256Map::Entry& s2) {
    auto f = b.storage.find(s2.first);
    if (f == b.storage.end()) bad("storage trie missing for an account");
    if (!cmp.eq(f->second, s2.second)) bad("storage trie: " + cmp.why);
  });
  if (b.pre_accounts.size() != b2.pre_accounts.size()) bad("pre-image account counts");
  for (size_t i = 0; i < b.pre_accounts.size(); i++) {
    const auto &p = b.pre_accounts[i], &q = b2.pre_accounts[i];
    if (!(p.haddr == q.haddr) || p.storage_nonempty != q.storage_nonempty || memcmp(&J.A.accounts[p.rec], &J2->A.accounts[q.rec], 128) != 0)
      bad("pre-image account " + std::to_string(i));
  }
  if (b.pre_with_storage.size() != b2.pre_with_storage.size()) bad("accounts with storage");
  b2.pre_with_storage.for_each([&](const H256Map::Entry& s2) {
    if (!b.pre_with_storage.count(s2.first)) bad("account with storage missing");
  });
  if (b.pre_code.size() != b2.pre_code.size()) bad("code map sizes");
  for (auto& c2 : b2.pre_code) {
    auto f = b.pre_code.find(c2.first);
    if (f == b.pre_code.end() || f->second.p != c2.second.p || f->second.n != c2.second.n) bad("code map entry");
  }
  b.root_of.for_each([&](uint32_t root, uint32_t root_node) {
    if (root_node >= J.A.nodes.size() || J.A.kind(root_node) != NK_ROOT || J.A.nodes[root_node].a1 != root) bad("root_of entry");
  });
}

struct PhaseTimer {
  bool on = getenv("PPD_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[ppd] %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};
// accumulating timer for the sections of the txn loop (PPD_TIMING only)
struct SectionTimer {
  bool on = getenv("PPD_TIMING") != nullptr;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::chrono::steady_clock::time_point t;
  void start() {
    if (on) t = std::chrono::steady_clock::now();
  }
  void stop(int k) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    acc[k] += std::chrono::duration<double, std::milli>(now - t).count();
    t = now;
  }
  void report(const char* const* names, int n) {
    if (!on) return;
    for (int k = 0; k < n; k++) fprintf(stderr, "[ppd]   %-12s %8.3f ms\n", names[k], acc[k]);
  }
};

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

void apply_withdrawals(Job& J, BlockJob& b, uint32_t& state_root) {
  HostArena& A = J.A;
  for (size_t i = 0; i < b.withdrawals.size(); i++) {
    const H256& h = J.kh.digest[b.m_withdrawal_addr[i]];
    uint32_t koff = key_from_digest(J, h);
    uint32_t leaf = A.get(state_root, koff, 64);
    if (leaf == NODE_EMPTY) fail(PPD_ERR_MISSING_WITHDRAWAL_ACCOUNT, "withdrawal to an account that is not in the state trie");
    if (A.kind(leaf) != NK_LEAF_ACCOUNT) fail(PPD_ERR_ACCOUNT_DECODE, "withdrawal account does not decode");
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
  if (!b.pre_image_on_gpu) build_pre_image(J, b);
  sec.stop(0);
  const uint32_t initial_state = b.state_root;
  // the storage tries before the first txn: only the dummy IRs of a block with at most one txn read them (decoding.rs:304-347)
  H256Map initial_storage;
  if (b.txns.size() <= 1) initial_storage = b.storage;
  uint32_t state = b.state_root, txn_trie = NODE_EMPTY, receipt_trie = NODE_EMPTY;
  uint64_t txn_before = 0, gas_before = 0, gas_after = 0;

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
    // every key this txn touches: marked on the device after the sweep's upload (mark_walk_kernel), or here in one
    // interleaved pass (HostArena::mark_many) when the block is dumped by the host or is being redone for its error
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
    bool short_haddr = false;
    for (size_t i = 0; i < tx.traces.size(); i++) {
      TraceV& tr = tx.traces[i];
      const H256& haddr = J.kh.digest[tr.m_addr];
      if (haddr.b[0] == 0) {  // reported after the marks collected so far (they come first in the reference's order)
        short_haddr = true;
        break;
      }
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
    if (J.device_marks) {
      p.items.assign(marks.begin(), marks.end());  // walked by mark_walk_kernel once the arena is resident
    } else {
      p.touched.reserve(marks.size() * 10);
      A.mark_many(marks.data(), marks.size(), p.touched);
      // the marking walk of an address also finds its leaf in the pre-txn state: the state writes below read the account
      // from it (other addresses' writes in between only path-copy branches; the leaf's payload stays)
      for (size_t i = 0; i < tx.traces.size() && i < marks.size(); i++) haddr_leaf[i] = marks[i].leaf;
    }
    if (short_haddr) fail(PPD_PANIC_H256_FROM_SLICE, "H256::from_slice on a short bytes_be()");
    gas_after += tx.gas_used;
    sec.stop(2);

    // ---- apply_deltas_to_trie_state (decoding.rs:219-292) ----
    for (TraceV& tr : tx.traces) {
      const H256& haddr = J.kh.digest[tr.m_addr];
      auto f = b.storage.find(haddr);
      if (f == b.storage.end()) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, "no storage trie for a written account");
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
      // the account as state.get() would give it (decoding.rs:251-254): the leaf the host's marking walk found, or, when
      // the marking walks are left to the device, the record tracked per hashed address
      AccountRec rec;
      const H256Map::Entry* cur = J.device_marks ? b.acct_rec.find(haddr) : nullptr;
      const uint32_t leaf = J.device_marks ? NODE_EMPTY : haddr_leaf[i];
      if (cur) {
        rec = A.accounts[cur->second];
      } else if (leaf != NODE_EMPTY) {
        if (A.kind(leaf) != NK_LEAF_ACCOUNT) fail(PPD_ERR_ACCOUNT_DECODE, "state leaf is not an account");
        rec = A.accounts[A.nodes[leaf].a1];
      } else {
        memset(&rec, 0, sizeof rec);
        memcpy(rec.storage_root, EMPTY_TRIE_HASH, 32);
        memcpy(rec.code_hash, EMPTY_CODE_HASH, 32);
        rec.storage_src = NODE_EMPTY;
      }
      if (storage_change) {
        auto f = b.storage.find(haddr);
        if (f == b.storage.end()) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, "no storage trie for a changed account");
        rec.storage_src = root_node_for(J, b, f->second);
      }
      if (tr.flags & PPD_TR_BALANCE) memcpy(rec.balance, tr.balance, 32);
      if (tr.flags & PPD_TR_NONCE) memcpy(rec.nonce, tr.nonce, 32);
      if (tr.flags & PPD_TR_CODE_READ) memcpy(rec.code_hash, tr.code_read, 32);
      if (tr.flags & PPD_TR_CODE_WRITE) memcpy(rec.code_hash, J.kh.digest[tr.m_code].b, 32);
      uint32_t r = (uint32_t)A.accounts.size();
      A.accounts.push_back(rec);
      if (J.device_marks) b.acct_rec[haddr] = r;
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
      if (!b.storage.erase(haddr)) fail(PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE, "self-destructed account has no storage trie");
      if (J.device_marks) b.acct_rec.erase(haddr);
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

// ---- step 4: the sweep ---------------------------------------------------------------------------
void sweep(Lane* c, Job& J, bool refs_to_host = true) {
  HostArena& A = J.A;
  uint32_t n = (uint32_t)A.nodes.size();
  J.ref.resize(32ull * n);
  J.ref_len.resize(n);
  J.refs_on_host = true;
  if (!n) return;
#ifdef PPD_HOSTPROF
  memset(J.ref.data(), 0, 32ull * n);
  memset(J.ref_len.data(), 32, n);
  return;
#endif
  // node ids counting-sorted by (level, class): inside a level, nodes of one kind and one permutation count are
  // adjacent, so the lanes of a warp do the same work.  The sort itself runs on the device (ppd_kernels.cu:
  // launch_order_by_level_class); the host only needs where every level starts.
  uint32_t n_levels = 0;
  for (uint32_t i = 0; i < n; i++) n_levels = std::max<uint32_t>(n_levels, A.level[i] + 1u);
  std::vector<uint32_t> level_start(n_levels + 1, 0);
  for (uint32_t i = 0; i < n; i++) level_start[A.level[i] + 1u]++;
  for (uint32_t l = 0; l < n_levels; l++) level_start[l + 1] += level_start[l];
  const bool device_order = 64u * n_levels <= ORDER_MAX_BINS;
  PVec<uint32_t>& order = J.order;
  if (!device_order) {  // a trie deeper than 64 levels: the same sort on the host
    auto node_class = [&](uint32_t i) -> uint32_t {
      const NodeRec& r = A.nodes[i];
      uint32_t kind = r.w0 & 0xff;
      if (kind == NK_BRANCH) return 40 + ((uint32_t)__builtin_popcount(r.a1 & 0xffff) - 1 & 15);  // by child count
      if (kind == NK_ROOT) return 56;
      uint32_t perms = 1;
      if (kind == NK_LEAF) {
        uint32_t nl = (r.w0 >> 16) & 0xff;
        perms = ((nl < 2 ? 1 : 2 + (nl >> 1)) + r.a2 + 6) / 136 + 1;  // header bytes over-estimated by at most 3
      }
      return kind * 8 + (perms > 8 ? 7 : perms - 1);
    };
    order.resize(n);
    std::vector<uint8_t> cls(n);
    std::vector<uint32_t> bucket((size_t)n_levels * 64 + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
      cls[i] = (uint8_t)node_class(i);
      bucket[(size_t)A.level[i] * 64 + cls[i] + 1]++;
    }
    for (size_t k = 0; k + 1 < bucket.size(); k++) bucket[k + 1] += bucket[k];
    for (uint32_t i = 0; i < n; i++) order[bucket[(size_t)A.level[i] * 64 + cls[i]]++] = i;
  }
  // the part of every pool that gpu_pre_image left in the lane's buffers stays where it is
  const Job::Resident& R = J.dev;
  c->d_nodes.reserve_keep(16ull * n, 16ull * R.nodes, c->st);
  c->d_order.reserve(4ull * n);
  c->d_keys.reserve_keep(A.key_pool.size() + 16, R.keys, c->st);
  c->d_vals.reserve_keep(A.val_pool.size() + 16, R.vals, c->st);
  c->d_hashes.reserve_keep(A.hash_pool.size() + 32, R.hashes, c->st);
  c->d_children.reserve_keep(4ull * A.child_pool.size() + 16, 4ull * R.children, c->st);
  c->d_accounts.reserve_keep(sizeof(AccountRec) * A.accounts.size() + 16, sizeof(AccountRec) * R.accounts, c->st);
  c->d_ref.reserve(32ull * n);
  c->d_ref_len.reserve(n);
  c->d_counters.reserve(32);
  auto up = [&](DevBuf& d, const void* src, size_t bytes, size_t resident = 0) {
    if (bytes <= resident) return;
    CUDA_OK(cudaMemcpyAsync(d.as<uint8_t>() + resident, (const uint8_t*)src + resident, bytes - resident, cudaMemcpyHostToDevice, c->st));
    c->stats.h2d_bytes += (double)(bytes - resident);
  };
  up(c->d_nodes, A.nodes.data(), 16ull * n, 16ull * R.nodes);
  if (device_order) {
    c->d_level.reserve_keep(2ull * n + 16, 2ull * R.nodes, c->st);
    c->d_okeys.reserve(2ull * n + 16);
    c->d_obins.reserve(4ull * ORDER_MAX_BINS);
    up(c->d_level, A.level.data(), 2ull * n, 2ull * R.nodes);
    CUDA_OK(cudaMemsetAsync(c->d_obins.p, 0, 4ull * 64 * n_levels, c->st));
    launch_order_by_level_class(c->d_nodes.as<NodeRec>(), c->d_level.as<uint16_t>(), n, 64 * n_levels, c->d_okeys.as<uint16_t>(),
                                c->d_obins.as<uint32_t>(), c->d_order.as<uint32_t>(), c->st);
    c->stats.kernel_launches += 3;
  } else {
    up(c->d_order, order.data(), 4ull * n);
  }
  up(c->d_keys, A.key_pool.data(), A.key_pool.size(), R.keys);
  up(c->d_vals, A.val_pool.data(), A.val_pool.size(), R.vals);
  up(c->d_hashes, A.hash_pool.data(), A.hash_pool.size(), R.hashes);
  up(c->d_children, A.child_pool.data(), 4ull * A.child_pool.size(), 4ull * R.children);
  up(c->d_accounts, A.accounts.data(), sizeof(AccountRec) * A.accounts.size(), sizeof(AccountRec) * R.accounts);
  CUDA_OK(cudaMemsetAsync(c->d_counters.p, 0, 32, c->st));
  ArenaView V;
  V.nodes = c->d_nodes.as<NodeRec>();
  V.key_pool = c->d_keys.as<uint8_t>();
  V.val_pool = c->d_vals.as<uint8_t>();
  V.hash_pool = c->d_hashes.as<uint8_t>();
  V.child_pool = c->d_children.as<uint32_t>();
  V.accounts = c->d_accounts.as<AccountRec>();
  V.ref = c->d_ref.as<uint8_t>();
  V.ref_len = c->d_ref_len.as<uint8_t>();
  V.counters = c->d_counters.as<unsigned long long>();
  CUDA_OK(cudaEventRecord(c->ev0, c->st));
  for (uint32_t l = 0; l < n_levels; l++) {
    launch_hash_level(V, c->d_order.as<uint32_t>(), level_start[l], level_start[l + 1], c->st);
    c->stats.kernel_launches++;
    c->stats.level_launches++;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(c->ev1, c->st));
  unsigned long long counters[3] = {0, 0, 0};
  c->has_last = true;
  c->last_view = V;
  c->last_level_start = level_start;
  c->last_n_msgs = (uint32_t)J.kh.lens.size();
  J.refs_on_host = refs_to_host;
  if (refs_to_host) {
    CUDA_OK(cudaMemcpyAsync(J.ref.data(), c->d_ref.p, 32ull * n, cudaMemcpyDeviceToHost, c->st));
    CUDA_OK(cudaMemcpyAsync(J.ref_len.data(), c->d_ref_len.p, n, cudaMemcpyDeviceToHost, c->st));
  }
  CUDA_OK(cudaMemcpyAsync(counters, c->d_counters.p, 24, cudaMemcpyDeviceToHost, c->st));
  lane_sync(c);
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->stats.gpu_ms += ms;
  c->stats.nodes_hashed += counters[0];
  c->stats.node_permutations += counters[1];
  c->stats.node_bytes += counters[2];
  c->stats.arena_nodes += n;
  c->stats.levels += n_levels;
  if (refs_to_host) c->stats.d2h_bytes += 33.0 * n;
}

// the refs of the last sweep, for the host paths that need them (host IR serialisation)
void fetch_refs(Lane* c, Job& J) {
  if (J.refs_on_host) return;
  size_t n = J.A.nodes.size();
  CUDA_OK(cudaMemcpyAsync(J.ref.data(), c->d_ref.p, 32ull * n, cudaMemcpyDeviceToHost, c->st));
  CUDA_OK(cudaMemcpyAsync(J.ref_len.data(), c->d_ref_len.p, n, cudaMemcpyDeviceToHost, c->st));
  lane_sync(c);
  c->stats.d2h_bytes += 33.0 * n;
  J.refs_on_host = true;
}

// the value and hash pools of a GPU-built pre-image, for the host paths that read them (host IR serialisation)
void fetch_pools(Lane* c, Job& J) {
  if (J.pools_on_host) return;
  if (J.dev.vals) CUDA_OK(cudaMemcpyAsync(J.A.val_pool.data(), c->d_vals.p, J.dev.vals, cudaMemcpyDeviceToHost, c->st));
  if (J.dev.hashes) CUDA_OK(cudaMemcpyAsync(J.A.hash_pool.data(), c->d_hashes.p, J.dev.hashes, cudaMemcpyDeviceToHost, c->st));
  lane_sync(c);
  c->stats.d2h_bytes += (double)(J.dev.vals + J.dev.hashes);
  J.pools_on_host = true;
}

// ---- step 5: IrDump ------------------------------------------------------------------------------
// Growable byte buffer with unchecked-after-need() writes; give() hands the malloc'ed storage to the caller.
struct Out {
  uint8_t* p = nullptr;
  size_t n = 0, cap = 0;
  Out() {}
  Out(const Out&) = delete;
  Out& operator=(const Out&) = delete;
  Out(Out&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr, o.n = o.cap = 0; }
  ~Out() { free(p); }
  void need(size_t k) {
    if (n + k <= cap) return;
    size_t nc = cap ? cap * 2 : 4096;
    while (nc < n + k) nc *= 2;
    uint8_t* q;
    if (nc >= (8u << 20)) {
      // large output buffers: 2 MiB-aligned and advised for transparent huge pages, so that first-touch
      // costs a few dozen page faults instead of thousands (free() releases it like any malloc block)
      nc = (nc + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
      q = (uint8_t*)aligned_alloc(2u << 20, nc);
      if (q) {
        madvise(q, nc, MADV_HUGEPAGE);
        if (n) memcpy(q, p, n);
        free(p);
      }
    } else {
      q = (uint8_t*)realloc(p, nc);
    }
    if (!q) fail(PPD_ERR_BAD_ARGUMENT, "out of host memory");
    p = q, cap = nc;
  }
  void u8(uint8_t v) {
    need(1);
    p[n++] = v;
  }
  void u32(uint32_t v) {
    need(4);
    memcpy(p + n, &v, 4);
    n += 4;
  }
  void u64(uint64_t v) {
    need(8);
    memcpy(p + n, &v, 8);
    n += 8;
  }
  void raw(const uint8_t* q, size_t k) {
    need(k);
    if (k) memcpy(p + n, q, k);
    n += k;
  }
  void span(Span s) {
    u32(s.n);
    raw(s.p, s.n);
  }
  void u256(uint64_t v) {
    need(32);
    memset(p + n, 0, 24);
    for (int i = 0; i < 8; i++) p[n + 31 - i] = (uint8_t)(v >> (8 * i));
    n += 32;
  }
  uint8_t* give(size_t* len) {
    uint8_t* r = p ? p : (uint8_t*)malloc(1);
    *len = n;
    p = nullptr, n = cap = 0;
    return r;
  }
};

// per-thread marks of the nodes a subset keeps expanded
struct Stamp {
  std::vector<uint32_t> v;
  uint32_t serial = 0;
};

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

// the marking walks of an IR that were left to the device, run on the host instead (host serialisation of the IR)
void materialize_touched(const Job& J, IrPlan& p) {
  if (p.items.empty()) return;
  p.touched.reserve(p.touched.size() + p.items.size() * 10);
  J.A.mark_many(p.items.data(), p.items.size(), p.touched);
  p.items.clear();
}

void dump_ir(const Job& J, const BlockJob& b, IrPlan& p, Stamp& st, Out& o) {
  materialize_touched(J, p);
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

// Runs f(item, worker) for every item in [0, n) on up to `workers` threads (the caller's included).
template <class F>
void parallel_for(size_t n, unsigned workers, F f) {
  if (workers > n) workers = (unsigned)n;
  if (workers <= 1) {
    for (size_t i = 0; i < n; i++) f(i, 0u);
    return;
  }
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  Fail first{PPD_OK, ""};
  std::mutex mu;
  auto body = [&](unsigned w) {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n || failed.load()) return;
      try {
        f(i, w);
      } catch (const Fail& e) {
        std::lock_guard<std::mutex> g(mu);
        if (!failed.exchange(true)) first = e;
      }
    }
  };
  std::vector<std::thread> th;
  for (unsigned w = 1; w < workers; w++) th.emplace_back(body, w);
  body(0);
  for (auto& t : th) t.join();
  if (failed.load()) throw first;
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
  for (BlockJob& b : J.blocks)
    if (b.status == PPD_OK)
      for (IrPlan& p : b.irs) materialize_touched(J, p);  // (throws what the reference's marking pass reports)
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

template <class F>
int guarded(ppd_ctx* c, F f) {
  if (!c) return PPD_ERR_BAD_ARGUMENT;
  try {
#ifndef PPD_HOSTPROF
    CUDA_OK(cudaSetDevice(c->device));
#endif
    f();
    return PPD_OK;
  } catch (const Fail& e) {
    c->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    c->err = e.what();
    return PPD_ERR_BAD_ARGUMENT;
  }
}

// ---- page-locked output buffers ---------------------------------------------------------------------
// The IrDump of a block is ~50 MB that the caller owns until ppd_free().  Handing out page-locked
// buffers from a process-wide pool lets the device write the result straight into the caller's buffer
// (no bounce copy, no first-touch page faults); ppd_free() returns the buffer to the pool.  The pool is
// capped (PPD_PINNED_OUT_MB, default 12288): beyond the cap outputs are ordinary malloc blocks.
struct OutPool {
  struct Entry {
    uint8_t* p;
    size_t cap;
    bool in_use;
  };
  std::mutex mu;
  std::vector<Entry> entries;
  size_t total = 0;
  size_t limit() {
    static size_t v = [] {
      const char* e = getenv("PPD_PINNED_OUT_MB");
      return (size_t)(e ? atoll(e) : 12288) << 20;
    }();
    return v;
  }
  uint8_t* take(size_t n) {
#ifdef PPD_HOSTPROF
    return nullptr;
#else
    std::lock_guard<std::mutex> g(mu);
    Entry* best = nullptr;
    for (Entry& e : entries)
      if (!e.in_use && e.cap >= n && (!best || e.cap < best->cap)) best = &e;
    if (best) {
      best->in_use = true;
      return best->p;
    }
    size_t cap = (n + (n >> 3) + (8u << 20) - 1) & ~(size_t)((8u << 20) - 1);
    if (total + cap > limit()) {
      // make room by releasing idle buffers that were too small
      for (size_t i = 0; i < entries.size() && total + cap > limit();)
        if (!entries[i].in_use) {
          cudaFreeHost(entries[i].p);
          total -= entries[i].cap;
          entries.erase(entries.begin() + i);
        } else {
          i++;
        }
      if (total + cap > limit()) return nullptr;
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    entries.push_back({(uint8_t*)p, cap, true});
    total += cap;
    return (uint8_t*)p;
#endif
  }
  bool give_back(void* p) {
    std::lock_guard<std::mutex> g(mu);
    for (Entry& e : entries)
      if (e.p == p) {
        e.in_use = false;
        return true;
      }
    return false;
  }
};
OutPool& out_pool() {
  static OutPool* p = new OutPool();  // never destroyed: buffers may outlive every context
  return *p;
}

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

enum { DUMP_ON_HOST = 0, DUMP_DONE = 1, DUMP_REDO_HOST_MARKS = 2 };
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
  size_t n_touched = 0, n_items = 0;
  for (IrPlan& p : b.irs) n_touched += p.touched.size() + MARK_SLOTS * p.items.size(), n_items += p.items.size();
  {
    // the two dump kernels cost two read-backs and about a millisecond of launch + CTA latency whatever the size;
    // a small block (config 1: a few thousand touched nodes) is serialised faster by the host threads
    const char* e = getenv("PPD_GPU_DUMP_MIN_TOUCHED");
    const size_t min_touched = e ? (size_t)atoll(e) : 32768;
    if (n_touched - (MARK_SLOTS - 8) * n_items < min_touched) return DUMP_ON_HOST;  // a marking walk touches about eight nodes
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
    touched_begin.push_back((uint32_t)(touched_begin.back() + p.touched.size() + MARK_SLOTS * p.items.size()));
  }
  const uint32_t n_seg = (uint32_t)seg_a.size();
  // one pinned buffer / one device buffer: [touched | touched_begin | seg_a | seg_b | seg_begin | ir_base(u64) |
  //                                        seg_off | ir_size | ir_flag | ir_nuniq | u_node | u_size | u_off]
  auto al = [](size_t x) { return (x + 3) & ~(size_t)3; };  // keep the u64 array 8-byte aligned (counts in u32 words)
  // (items4: the marking walks left to the device, 4 words each; mark_flags: [n_ir + 1] written by mark_walk_kernel)
  const size_t o_touched = 0, o_tb = al(o_touched + n_touched), o_sa = al(o_tb + n_ir + 1), o_sb = al(o_sa + n_seg), o_sg = al(o_sb + n_seg),
               o_it = al(o_sg + n_ir + 1), o_base = al(o_it + 4 * n_items), o_in_end = al(o_base + 2 * (size_t)n_ir);
  const size_t o_soff = o_in_end, o_isz = al(o_soff + n_seg), o_ifl = al(o_isz + n_ir), o_mf = al(o_ifl + n_ir), o_inu = al(o_mf + n_ir + 1),
               o_un = al(o_inu + n_ir), o_us = al(o_un + n_touched), o_uo = al(o_us + n_touched), o_end = al(o_uo + n_touched);
  J.plan.resize(o_end);
  uint32_t* h = J.plan.data();
  {
    uint32_t* t = h + o_touched;
    uint32_t* it4 = h + o_it;
    for (uint32_t ir = 0; ir < n_ir; ir++) {
      IrPlan& p = b.irs[ir];
      if (!p.touched.empty()) memcpy(t, p.touched.data(), 4 * p.touched.size());
      t += p.touched.size();
      if (!p.items.empty()) {
        memset(t, 0xff, 4 * MARK_SLOTS * p.items.size());  // NODE_EMPTY: slots a walk does not reach
        for (const HostArena::MarkItem& m : p.items) {
          it4[0] = m.root, it4[1] = m.koff, it4[2] = (m.klen & 0xffu) | (ir << 8), it4[3] = (uint32_t)(t - (h + o_touched));
          it4 += 4, t += MARK_SLOTS;
        }
      }
    }
  }
  memcpy(h + o_tb, touched_begin.data(), 4 * (n_ir + 1));
  memcpy(h + o_sa, seg_a.data(), 4 * n_seg);
  memcpy(h + o_sb, seg_b.data(), 4 * n_seg);
  memcpy(h + o_sg, seg_begin.data(), 4 * (n_ir + 1));
  L->d_plan.reserve(4 * o_end);
  uint32_t* d = L->d_plan.as<uint32_t>();
  CUDA_OK(cudaMemcpyAsync(d, h, 4 * o_base, cudaMemcpyHostToDevice, L->st));
  IrDumpPlanView P;
  P.touched = d + o_touched, P.touched_begin = d + o_tb, P.seg_a = d + o_sa, P.seg_b = d + o_sb, P.seg_begin = d + o_sg;
  P.ir_base = reinterpret_cast<const uint64_t*>(d + o_base);
  P.seg_off = d + o_soff, P.ir_size = d + o_isz, P.ir_flag = d + o_ifl, P.ir_nuniq = d + o_inu;
  P.u_node = d + o_un, P.u_size = d + o_us, P.u_off = d + o_uo;
  pt.lap("  d:plan");
  if (n_items) {
    CUDA_OK(cudaMemsetAsync(d + o_mf, 0, 4 * (n_ir + 1), L->st));
    launch_mark_walk(L->last_view, d + o_it, (uint32_t)n_items, n_ir, d + o_touched, d + o_mf, L->st);
    L->stats.kernel_launches += 1;
    L->stats.marks_on_gpu += n_items;
  }
  launch_ir_size(L->last_view, P, n_ir, L->st);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(h + o_soff, d + o_soff, 4 * (o_inu - o_soff), cudaMemcpyDeviceToHost, L->st));  // seg_off, ir_size, ir_flag
  lane_sync(L);
  L->stats.h2d_bytes += 4.0 * o_base, L->stats.d2h_bytes += 4.0 * (o_inu - o_soff), L->stats.kernel_launches += 1;
  pt.lap("  d:size");
  // ---- IRs the device could not lay out: host serialisation ----
  const uint32_t *seg_off = h + o_soff, *ir_size = h + o_isz;
  uint32_t* ir_flag = h + o_ifl;
  if (n_items) {
    const uint32_t* mark_flags = h + o_mf;
    // a key ran into a hashed-out node: the reference reports MissingKeysCreatingSubPartialTrie, possibly after other
    // errors of earlier txns; the block is redone with the host's marking pass, which keeps the reference's order
    if (mark_flags[n_ir]) return DUMP_REDO_HOST_MARKS;
    for (uint32_t i = 0; i < n_ir; i++) ir_flag[i] |= mark_flags[i];  // a walk longer than its slots: the host serialises that IR
  }
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

// One block on one lane: parse, key hashes, shaping, sweep, dump.  A failure that is the block's own
// (bad input, a reference panic site) is reported through *status; a CUDA failure is thrown.
// Whether the block's subset marking walks are left to the device: only when the device also serialises the IRs (same
// threshold as gpu_dump_block; a marking walk touches about eight nodes).
bool device_marks_wanted(const BlockJob& b) {
  // Opt-in (PPD_DEVICE_MARKS): measured on the 16-core box the host's own marking pass is the better choice today,
  // because its interleaved, prefetched walks are what brings the paths into the cache that the inserts of the same
  // txn then copy; without it the inserts take those misses one by one (state writes 7 -> 28 ms per block, 410 -> 296
  // blocks/s).  It pays once the inserts themselves run on the device.
  if (!gpu_dump_enabled() || !getenv("PPD_DEVICE_MARKS")) return false;
  size_t n_marks = 0;
  for (const TxnV& tx : b.txns) {
    n_marks += 2;
    for (const TraceV& tr : tx.traces) n_marks += 1 + tr.n_reads + tr.n_writes;
  }
  const char* e = getenv("PPD_GPU_DUMP_MIN_TOUCHED");
  const size_t min_touched = e ? (size_t)atoll(e) : 32768;
  return n_marks > 0 && 8 * n_marks >= min_touched;
}

void decode_one(ppd_ctx* c, Lane* L, const uint8_t* flat, size_t len, uint8_t** out, size_t* out_len, int* status, unsigned dump_workers) {
  *out = nullptr, *out_len = 0;
  const ppd_stats stats0 = L->stats;
  // Second attempt: only after a first one with device-side marking walks that met an error.  With the marks
  // deferred to the device the host cannot tell whether an earlier txn's marking pass would have failed first, so the
  // block is redone with the host's own marking pass, which reports errors in the reference's order.
  for (int attempt = 0; attempt < 2; attempt++) {
    PhaseTimer pt;
    Job& J = job_of(L, 1);
    BlockJob& b = J.blocks[0];
    bool redo = false;
    try {
      read_flat_block(flat, len, b);
      J.device_marks = attempt == 0 && device_marks_wanted(b);
      pt.lap("read-flat");
      if (gpu_parse_enabled()) gpu_pre_image(L, J, b, true, &c->parse_slots_sem);
      collect_messages(J, b);
      pt.lap("parse");
      J.kh.run(L);
      pt.lap("keyhash");
      shape_block(J, b);
      pt.lap("shape");
      sweep(L, J, /*refs_to_host=*/!gpu_dump_enabled());
      pt.lap("sweep");
      const int r = gpu_dump_block(c, L, J, out, out_len);
      if (r == DUMP_REDO_HOST_MARKS) {
        redo = true;
      } else if (r == DUMP_ON_HOST) {
        fetch_refs(L, J);
        fetch_pools(L, J);
        dump_blocks(J, out, out_len, dump_workers);
      }
      pt.lap("dump");
    } catch (const Fail& e) {
      if (e.code == PPD_ERR_CUDA) throw;
      if (J.device_marks) {
        redo = true;
      } else {
        *status = e.code;
        std::lock_guard<std::mutex> g(c->err_mu);
        c->err = e.msg;
        return;
      }
    }
    if (!redo) {
      *status = PPD_OK;
      return;
    }
    L->stats = stats0;
  }
}

void add_stats(ppd_stats& a, const ppd_stats& b) {
  a.nodes_hashed += b.nodes_hashed, a.node_permutations += b.node_permutations, a.key_hashes += b.key_hashes;
  a.key_permutations += b.key_permutations, a.node_bytes += b.node_bytes, a.arena_nodes += b.arena_nodes;
  a.levels = std::max(a.levels, b.levels);
  a.gpu_ms += b.gpu_ms, a.h2d_bytes += b.h2d_bytes, a.d2h_bytes += b.d2h_bytes, a.kernel_launches += b.kernel_launches;
  a.witnesses_on_gpu += b.witnesses_on_gpu, a.witness_instructions += b.witness_instructions, a.witness_bytes += b.witness_bytes;
  a.parse_gpu_ms += b.parse_gpu_ms, a.level_launches += b.level_launches, a.marks_on_gpu += b.marks_on_gpu;
}

// Blocks are independent (each BlockTrace carries its own pre-image, trace_protocol.rs:40-48): every
// host thread takes blocks on its own lane, so parsing / shaping of one block overlaps the copies and
// kernels of the others.
static const size_t MAX_LANES = 64;

void decode_blocks(ppd_ctx* c, const uint8_t* const* flats, const size_t* lens, size_t n, uint8_t** outs, size_t* out_lens, int* statuses) {
  stats_reset(c);
  if (!n) return;
  // block i runs on lane i mod n_lanes; a host thread takes whole lanes, so a lane never runs two
  // blocks at once and the last block of every lane stays resident in HBM (ppd_replay_last_hashing)
  const size_t n_lanes = std::min(n, MAX_LANES);
  const unsigned workers = (unsigned)std::min<size_t>(host_threads(), n_lanes);
  for (size_t w = 0; w < n_lanes; w++) {
    Lane* L = lane_of(c, w);
    L->stats = ppd_stats{};
    L->has_last = false;
    L->has_last_parse = false;
  }
  const unsigned dump_workers = std::max(1u, host_threads() / workers);
  for (size_t i = 0; i < n; i++) outs[i] = nullptr, out_lens[i] = 0, statuses[i] = PPD_OK;
  parallel_for(n_lanes, workers, [&](size_t lane, unsigned) {
#ifndef PPD_HOSTPROF
    CUDA_OK(cudaSetDevice(c->device));
#endif
    for (size_t i = lane; i < n; i += n_lanes) decode_one(c, c->lanes[lane], flats[i], lens[i], &outs[i], &out_lens[i], &statuses[i], dump_workers);
  });
  c->last_lanes_used = n_lanes;
  for (size_t w = 0; w < n_lanes; w++) add_stats(c->stats, c->lanes[w]->stats);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int ppd_ctx_create(int device, ppd_ctx** out) {
  if (!out) return PPD_ERR_BAD_ARGUMENT;
  *out = nullptr;
#ifdef PPD_HOSTPROF
  *out = new ppd_ctx();
  return PPD_OK;
#endif
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return PPD_ERR_CUDA;
  ppd_ctx* c = new ppd_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    delete c;
    return PPD_ERR_CUDA;
  }
  *out = c;
  return PPD_OK;
}

void ppd_ctx_destroy(ppd_ctx* c) {
  if (!c) return;
#ifdef PPD_HOSTPROF
  for (Lane* l : c->lanes) lane_delete(l);
  delete c;
  return;
#endif
  cudaSetDevice(c->device);
  for (Lane* l : c->lanes) lane_delete(l);
  DevBuf* bufs[] = {&c->d_keys, &c->d_vals, &c->d_ref, &c->d_ref_len, &c->d_counters, &c->d_msg, &c->d_msg_off, &c->d_digest};
  for (DevBuf* b : bufs) b->release();
  for (DevBuf& b : c->d_build) b.release();
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
}

const char* ppd_last_error(const ppd_ctx* c) { return c ? c->err.c_str() : "null context"; }
void ppd_last_stats(const ppd_ctx* c, ppd_stats* out) {
  if (c && out) *out = c->stats;
}
void ppd_free(void* p) {
  if (p && !out_pool().give_back(p)) free(p);
}
void* ppd_alloc_pinned(size_t n) {
  void* p = out_pool().take(n ? n : 1);
  return p ? p : malloc(n ? n : 1);
}

int ppd_keccak256_batch(ppd_ctx* c, const uint8_t* data, const uint64_t* offsets, size_t n, uint8_t* out32n) {
  return guarded(c, [&] {
    stats_reset(c);
    if (!n) return;
    size_t bytes = offsets[n];
    c->d_msg.reserve(bytes + 16);
    c->d_msg_off.reserve((n + 1) * 8);
    c->d_digest.reserve(n * 32);
    CUDA_OK(cudaMemcpyAsync(c->d_msg.p, data, bytes, cudaMemcpyHostToDevice, c->st));
    CUDA_OK(cudaMemcpyAsync(c->d_msg_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, c->st));
    CUDA_OK(cudaEventRecord(c->ev0, c->st));
    launch_keccak256_batch(c->d_msg.as<uint8_t>(), c->d_msg_off.as<uint64_t>(), (uint32_t)n, c->d_digest.as<uint8_t>(), c->st);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(c->ev1, c->st));
    CUDA_OK(cudaMemcpyAsync(out32n, c->d_digest.p, n * 32, cudaMemcpyDeviceToHost, c->st));
    CUDA_OK(cudaStreamSynchronize(c->st));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.gpu_ms = ms;
    c->stats.key_hashes = n;
    for (size_t i = 0; i < n; i++) c->stats.key_permutations += (offsets[i + 1] - offsets[i]) / 136 + 1;
    c->stats.h2d_bytes = (double)(bytes + (n + 1) * 8);
    c->stats.d2h_bytes = (double)(n * 32);
    c->stats.kernel_launches = 1;
  });
}

// Re-run every kernel of the last ppd_block_decode / ppd_blocks_decode_batch call on the arena and
// key messages that are still resident in HBM (no host work, no copies) and return its device time.
int ppd_replay_last_hashing(ppd_ctx* c, double* gpu_ms_out) {
  return guarded(c, [&] {
    size_t used = 0;
    for (size_t w = 0; w < c->last_lanes_used && w < c->lanes.size(); w++) used += c->lanes[w]->has_last;
    if (!used) fail(PPD_ERR_BAD_ARGUMENT, "no block job is resident");
    // all lanes start after ev0 on the main stream; the main stream then waits for every lane
    CUDA_OK(cudaEventRecord(c->ev0, c->st));
    for (size_t w = 0; w < c->last_lanes_used; w++) {
      Lane* L = c->lanes[w];
      if (!L->has_last) continue;
      CUDA_OK(cudaStreamWaitEvent(L->st, c->ev0, 0));
      CUDA_OK(cudaMemsetAsync(L->d_counters.p, 0, 32, L->st));
      launch_keccak256_ranges(L->d_msg.as<uint8_t>(), L->d_msg_off.as<uint64_t>(), L->last_n_msgs, L->d_digest.as<uint8_t>(), L->st);
      size_t nl = L->last_level_start.size() - 1;
      for (size_t l = 0; l < nl; l++)
        launch_hash_level(L->last_view, L->d_order.as<uint32_t>(), L->last_level_start[l], L->last_level_start[l + 1], L->st);
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaEventRecord(L->ev1, L->st));
      CUDA_OK(cudaStreamWaitEvent(c->st, L->ev1, 0));
    }
    CUDA_OK(cudaEventRecord(c->ev1, c->st));
    CUDA_OK(cudaStreamSynchronize(c->st));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (gpu_ms_out) *gpu_ms_out = ms;
  });
}

// The same for the witness parse / pre-image arena kernels (ppd_parse.cu) of the last call: every lane replays
// its three phases back to back on the witness still resident in HBM (the sizes the host read back between the
// phases are those of the first run, so nothing is copied or synchronised in between).
int ppd_replay_last_parse(ppd_ctx* c, double* gpu_ms_out) {
  return guarded(c, [&] {
    size_t used = 0;
    for (size_t w = 0; w < c->last_lanes_used && w < c->lanes.size(); w++) used += c->lanes[w]->has_last_parse;
    if (!used) fail(PPD_ERR_BAD_ARGUMENT, "no GPU-parsed witness is resident");
    CUDA_OK(cudaEventRecord(c->ev0, c->st));
    for (size_t w = 0; w < c->last_lanes_used; w++) {
      Lane* L = c->lanes[w];
      if (!L->has_last_parse) continue;
      CUDA_OK(cudaStreamWaitEvent(L->st, c->ev0, 0));
      ParseEmit E = L->last_emit;
      // the sweep may have moved the pools to larger buffers since
      E.nodes = L->d_nodes.as<NodeRec>(), E.key_pool = L->d_keys.as<uint8_t>(), E.val_pool = L->d_vals.as<uint8_t>();
      E.hash_pool = L->d_hashes.as<uint8_t>(), E.child_pool = L->d_children.as<uint32_t>(), E.accounts = L->d_accounts.as<AccountRec>();
      E.level = L->d_level.as<uint16_t>();
      CUDA_OK(cudaMemsetAsync(L->last_bounds.result, 0, 4 * PARSE_R_WORDS, L->st));
      launch_parse_bounds(L->last_bounds, L->st);
      launch_parse_scatter(L->last_bounds, L->last_ins_pos, L->st);
      launch_parse_tree(E.T, L->st);
      if (L->last_n_code) {
        launch_parse_code_list(E, L->st);
        launch_keccak256_ranges(E.T.wit, E.code_se, L->last_n_code, const_cast<uint8_t*>(E.code_digest), L->st);
      }
      if (L->last_val_bytes) CUDA_OK(cudaMemsetAsync(E.val_pool, 0, L->last_val_bytes, L->st));
      launch_parse_emit(E, L->st);
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaEventRecord(L->ev1, L->st));
      CUDA_OK(cudaStreamWaitEvent(c->st, L->ev1, 0));
    }
    CUDA_OK(cudaEventRecord(c->ev1, c->st));
    CUDA_OK(cudaStreamSynchronize(c->st));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (gpu_ms_out) *gpu_ms_out = ms;
  });
}

// Measurement hook (ppd_microbench.cu): variant 0 = dependent-free LOP3/SHF issue rate, 1.. = register
// resident keccak-f variants.  units_out = ALU instructions (variant 0) or permutations executed.
int ppd_microbench(ppd_ctx* c, int variant, uint32_t blocks_per_sm, uint32_t iters, double* gpu_ms_out, double* units_out, uint32_t* digest_out) {
  return guarded(c, [&] {
    c->d_counters.reserve(64);
    CUDA_OK(cudaMemsetAsync(c->d_counters.p, 0, 64, c->st));
    uint32_t bt = 0;
    double per = 0;
    // warm-up launch, then the timed one
    if (!launch_microbench(variant, c->d_counters.as<uint32_t>(), blocks_per_sm, iters, &bt, &per, c->st)) fail(PPD_ERR_BAD_ARGUMENT, "unknown microbench variant");
    CUDA_OK(cudaEventRecord(c->ev0, c->st));
    launch_microbench(variant, c->d_counters.as<uint32_t>(), blocks_per_sm, iters, &bt, &per, c->st);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(c->ev1, c->st));
    uint32_t out[4];
    CUDA_OK(cudaMemcpyAsync(out, c->d_counters.p, 16, cudaMemcpyDeviceToHost, c->st));
    CUDA_OK(cudaStreamSynchronize(c->st));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (gpu_ms_out) *gpu_ms_out = ms;
    if (units_out) *units_out = per * (double)iters * (double)bt * 148.0 * (double)blocks_per_sm;
    if (digest_out) digest_out[0] = out[2], digest_out[1] = out[3];
  });
}

int ppd_compact_decode(ppd_ctx* c, const uint8_t* witness, size_t len, uint8_t** out, size_t* out_len) {
  return guarded(c, [&] {
    stats_reset(c);
    Lane* L = lane_of(c, 0);
    L->stats = ppd_stats{};
    Job& J = job_of(L, 1);
    BlockJob& b = J.blocks[0];
    if (len >= 0xffffffffull) fail(PPD_ERR_BAD_ARGUMENT, "witness larger than 4 GiB");
    b.compact = Span{witness, (uint32_t)len};
    if (!(gpu_parse_enabled() && gpu_pre_image(L, J, b, /*check_version=*/false))) {
      parse_witness(witness, len, b.wit);
      collect_witness_messages(J, b);
      J.kh.run(L);
      build_pre_image(J, b);
    }
    uint32_t sr = root_node_for(J, b, b.state_root);
    std::map<H256, uint32_t> storage_roots;
    b.storage.for_each([&](const H256Map::Entry& s) { storage_roots[s.first] = root_node_for(J, b, s.second); });
    sweep(L, J);
    c->stats = L->stats;
    Out o;
    o.u32(PPD_PRE_IMAGE_MAGIC);
    o.u8(b.wit.version);
    o.raw(J.ref.data() + 32ull * sr, 32);
    o.u32((uint32_t)storage_roots.size());
    for (auto& s : storage_roots) {
      o.raw(s.first.b, 32);
      o.raw(J.ref.data() + 32ull * s.second, 32);
    }
    o.u32((uint32_t)b.pre_code.size());
    for (auto& cd : b.pre_code) {
      o.raw(cd.first.b, 32);
      o.u32(cd.second.n);
    }
    o.u64(c->stats.nodes_hashed);
    o.u64(c->stats.node_permutations);
    *out = o.give(out_len);
  });
}

int ppd_blocks_decode_batch(ppd_ctx* c, const uint8_t* const* flats, const size_t* lens, size_t n, uint8_t** outs, size_t* out_lens,
                            int* statuses) {
  return guarded(c, [&] { decode_blocks(c, flats, lens, n, outs, out_lens, statuses); });
}

int ppd_block_decode(ppd_ctx* c, const uint8_t* flat, size_t len, uint8_t** out, size_t* out_len) {
  int status = PPD_OK;
  int rc = guarded(c, [&] { decode_blocks(c, &flat, &len, 1, out, out_len, &status); });
  return rc != PPD_OK ? rc : status;
}

}  // extern "C"

// ---- trie root over sorted leaves: structure built and hashed on the GPU (ppd_build.cu) ----------
namespace {

void trie_root_sorted_dev(ppd_ctx* c, const uint8_t* d_keys, const uint64_t* d_val_off, const uint8_t* d_vals, size_t n_, uint8_t root_out[32]) {
  if (n_ == 0) {
    memcpy(root_out, EMPTY_TRIE_HASH, 32);  // HashedPartialTrie::default().hash(): keccak(0x80), types.rs:30-34
    return;
  }
  if (n_ >= 0x7fffffffull) fail(PPD_ERR_BAD_ARGUMENT, "at most 2^31 - 2 leaves per call");
  const uint32_t n = (uint32_t)n_;
  cudaStream_t st = c->st;
  const uint32_t n1 = (n + 1 + 63) / 64, n2 = (n1 + 63) / 64, n3 = (n2 + 63) / 64;
  enum { B_L, B_M, B_LINKA, B_LINKB, B_BIDX, B_TMP, B_SMALL, B_DEPTH, B_REP, B_CHILD, B_ORDER, B_NCHILD };
  DevBuf* D = c->d_build;
  D[B_L].reserve((size_t)n + 1 + 64);
  D[B_M].reserve((size_t)n1 + n2 + n3 + 192);
  D[B_LINKA].reserve(4ull * (n + 1));
  D[B_LINKB].reserve(4ull * (n + 1));
  D[B_BIDX].reserve(4ull * (n + 1));
  D[B_TMP].reserve(4ull * scan_tmp_words((size_t)n + 1));
  D[B_SMALL].reserve(16384);
  int8_t* L = D[B_L].as<int8_t>();
  int8_t* m1 = D[B_M].as<int8_t>();
  int8_t* m2 = m1 + ((n1 + 63) & ~63u);
  int8_t* m3 = m2 + ((n2 + 63) & ~63u);
  uint32_t* small = D[B_SMALL].as<uint32_t>();  // [0] error flags, [1] root id, [2..7] counters(u64 x 3), [8..15] root, [1024..2047] hist, [2048..3071] cursor
  CUDA_OK(cudaMemsetAsync(small, 0, 16384, st));
  CUDA_OK(cudaEventRecord(c->ev0, st));
  launch_lcp(d_keys, n, L, small + 0, st);
  launch_min64(L, n + 1, m1, n1, st);
  launch_min64(m1, n1, m2, n2, st);
  launch_min64(m2, n2, m3, n3, st);
  uint32_t* leader = D[B_LINKA].as<uint32_t>();
  uint32_t* flag = D[B_LINKB].as<uint32_t>();
  launch_leaders(L, m1, m2, m3, n, leader, flag, flag, st);
  uint32_t* bidx = D[B_BIDX].as<uint32_t>();
  exclusive_scan_u32(flag, bidx, n + 1, D[B_TMP].as<uint32_t>(), st);
  CUDA_OK(cudaGetLastError());
  uint32_t h_flags = 0, nb = 0;
  CUDA_OK(cudaMemcpyAsync(&h_flags, small, 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(&nb, bidx + n, 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  if (h_flags & 1) fail(PPD_ERR_UNSORTED_KEYS, "keys are not strictly ascending");
  c->stats.kernel_launches += 11 + 3;

  D[B_DEPTH].reserve(2ull * nb + 64);
  D[B_REP].reserve(4ull * nb + 64);
  D[B_CHILD].reserve(64ull * nb + 64);
  D[B_ORDER].reserve(4ull * nb + 64);
  c->d_ref.reserve(32ull * ((size_t)n + nb));
  c->d_ref_len.reserve((size_t)n + nb);
  BuildView V;
  V.keys = d_keys;
  V.val_off = d_val_off;
  V.vals = d_vals;
  V.n = n;
  V.P = Pyramid{L, m1, m2, m3};
  V.leader = leader;
  V.bidx = bidx;
  V.depth = D[B_DEPTH].as<uint8_t>();
  V.ext_start = V.depth + nb + 32;
  V.rep = D[B_REP].as<uint32_t>();
  V.child = D[B_CHILD].as<uint32_t>();
  V.root_id = small + 1;
  V.ref = c->d_ref.as<uint8_t>();
  V.ref_len = c->d_ref_len.as<uint8_t>();
  V.root_out = reinterpret_cast<uint8_t*>(small + 8);
  V.counters = reinterpret_cast<unsigned long long*>(small + 2);
  D[B_NCHILD].reserve(4ull * nb + 64);
  uint32_t* nchild = D[B_NCHILD].as<uint32_t>();
  if (nb) CUDA_OK(cudaMemsetAsync(V.child, 0xff, 64ull * nb, st));
  if (nb) CUDA_OK(cudaMemsetAsync(nchild, 0, 4ull * nb, st));
  launch_branch_info(V, st);
  launch_child_count(leader, bidx, n, nchild, st);
  // counting sort by (depth, number of children): the lanes of a warp walk the same number of children
  // and run the same number of permutations
  launch_depth_hist(V.depth, nchild, nb, small + 1024, st);
  std::vector<uint32_t> hist(1024), cursor(1024), start(1024);
  CUDA_OK(cudaMemcpyAsync(hist.data(), small + 1024, 4096, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  // deepest level first; inside a depth, child counts ascending
  uint32_t acc = 0;
  for (int d = 63; d >= 0; d--)
    for (int k = 0; k < 16; k++) {
      cursor[16 * d + k] = start[16 * d + k] = acc;
      acc += hist[16 * d + k];
    }
  if (acc != nb) fail(PPD_ERR_CUDA, "branch histogram does not add up");
  CUDA_OK(cudaMemcpyAsync(small + 2048, cursor.data(), 4096, cudaMemcpyHostToDevice, st));
  uint32_t* order = D[B_ORDER].as<uint32_t>();
  launch_branch_scatter(V.depth, nchild, nb, small + 2048, order, st);
  launch_hash_sorted_leaves(V, st);
  c->stats.kernel_launches += 5;
  uint32_t levels = 1;
  for (int d = 63; d >= 0; d--) {
    uint32_t cnt = 0;
    for (int k = 0; k < 16; k++) cnt += hist[16 * d + k];
    if (cnt) {
      launch_hash_branch_level(V, order, start[16 * d], start[16 * d] + cnt, st);
      c->stats.kernel_launches++;
      levels++;
    }
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(c->ev1, st));
  uint32_t out[16];
  CUDA_OK(cudaMemcpyAsync(out, small, 64, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  memcpy(root_out, out + 8, 32);
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  unsigned long long cnt[3];
  memcpy(cnt, out + 2, 24);
  c->stats.gpu_ms += ms;
  c->stats.nodes_hashed += cnt[0];
  c->stats.node_permutations += cnt[1];
  c->stats.node_bytes += cnt[2];
  c->stats.arena_nodes += (uint64_t)n + nb;
  c->stats.levels += levels;
  c->stats.d2h_bytes += 64 + 4096 + 8;
}

}  // namespace

extern "C" {

int ppd_trie_root_sorted_leaves_dev(ppd_ctx* c, const uint8_t* d_keys32, const uint64_t* d_val_off, const uint8_t* d_vals, size_t n,
                                    size_t vals_bytes, uint8_t root_out[32]) {
  (void)vals_bytes;
  return guarded(c, [&] {
    stats_reset(c);
    trie_root_sorted_dev(c, d_keys32, d_val_off, d_vals, n, root_out);
  });
}

int ppd_trie_root_sorted_leaves(ppd_ctx* c, const uint8_t* keys32, const uint64_t* val_off, const uint8_t* vals, size_t n,
                                uint8_t root_out[32]) {
  return guarded(c, [&] {
    stats_reset(c);
    if (!n) {
      trie_root_sorted_dev(c, nullptr, nullptr, nullptr, 0, root_out);
      return;
    }
    size_t vb = val_off[n];
    c->d_keys.reserve(32 * n);
    c->d_msg_off.reserve(8 * (n + 1));
    c->d_vals.reserve(vb + 64);
    CUDA_OK(cudaMemcpyAsync(c->d_keys.p, keys32, 32 * n, cudaMemcpyHostToDevice, c->st));
    CUDA_OK(cudaMemcpyAsync(c->d_msg_off.p, val_off, 8 * (n + 1), cudaMemcpyHostToDevice, c->st));
    CUDA_OK(cudaMemcpyAsync(c->d_vals.p, vals, vb, cudaMemcpyHostToDevice, c->st));
    c->stats.h2d_bytes += (double)(32 * n + 8 * (n + 1) + vb);
    trie_root_sorted_dev(c, c->d_keys.as<uint8_t>(), c->d_msg_off.as<uint64_t>(), c->d_vals.as<uint8_t>(), n, root_out);
  });
}

}  // extern "C"
