// host_arena_check.cpp — CPU-only checks of the host-side trie shaping (csrc/host_arena.h): the batched
// operations the txn loop uses must give exactly what the one-by-one operations give.
//   insert_many(sorted keys)  ==  insert, key by key (same canonical trie, node for node)
//   mark_many                 ==  mark, key by key (same touched nodes, same leaves)
//   branch_with               ==  rebuilding the branch from its 16 slots
//   flat_maps.h               ==  std::map under a random mix of operations
// Built and run by tests/test_host_cpu.py with g++ (no CUDA needed: the headers only shape tries).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../proof_protocol_decoder_b200/csrc/flat_maps.h"
#include "../../proof_protocol_decoder_b200/csrc/host_arena.h"
#include <map>

using namespace ppd;

static int failures = 0;
#define CHECK(cond, ...)                      \
  do {                                        \
    if (!(cond)) {                            \
      failures++;                             \
      fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);           \
      fprintf(stderr, "\n");                  \
    }                                         \
  } while (0)

static bool same(const HostArena& A, uint32_t x, uint32_t y) {
  if (x == NODE_EMPTY || y == NODE_EMPTY) return x == y;
  if (A.kind(x) != A.kind(y)) return false;
  if (is_hash_id(x)) return memcmp(A.hash_of(x), A.hash_of(y), 32) == 0;
  const NodeRec &a = A.nodes[x], &b = A.nodes[y];
  auto nibs = [&](uint32_t u, uint32_t v) {
    if (A.nstart(u) != A.nstart(v) || A.nlen(u) != A.nlen(v)) return false;
    for (uint32_t k = 0; k < A.nlen(u); k++)
      if (A.key_nib(A.nodes[u].a0, A.nstart(u) + k) != A.key_nib(A.nodes[v].a0, A.nstart(v) + k)) return false;
    return true;
  };
  switch (A.kind(x)) {
    case NK_LEAF:
      return nibs(x, y) && a.a2 == b.a2 && memcmp(A.val_pool.data() + a.a1, A.val_pool.data() + b.a1, a.a2) == 0;
    case NK_LEAF_ACCOUNT:
      return nibs(x, y) && a.a1 == b.a1;
    case NK_EXT:
      return nibs(x, y) && same(A, a.a1, b.a1);
    case NK_BRANCH: {
      if ((a.a1 & 0xffff) != (b.a1 & 0xffff)) return false;
      for (uint32_t nib = 0; nib < 16; nib++)
        if (!same(A, A.child_at(x, nib), A.child_at(y, nib))) return false;
      return true;
    }
    default:
      return same(A, a.a1, b.a1);
  }
}

// every node's level must exceed the level of everything it reads (the bottom-up sweep's only requirement)
static bool levels_ok(const HostArena& A, uint32_t x) {
  if (x == NODE_EMPTY || is_hash_id(x)) return true;
  const NodeRec& a = A.nodes[x];
  switch (A.kind(x)) {
    case NK_EXT:
      return (is_hash_id(a.a1) ? A.lvl(x) >= 1 : A.lvl(x) > A.lvl(a.a1)) && levels_ok(A, a.a1);
    case NK_BRANCH:
      for (uint32_t nib = 0; nib < 16; nib++) {
        uint32_t c = A.child_at(x, nib);
        if (c == NODE_EMPTY) continue;
        if (!(A.lvl(x) > A.lvl(c) || (is_hash_id(c) && A.lvl(x) >= 1))) return false;
        if (!levels_ok(A, c)) return false;
      }
      return true;
    default:
      return true;
  }
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 40;
  std::mt19937_64 rng(12345);
  for (int round = 0; round < rounds; round++) {
    HostArena A;
    A.accounts.resize(1);
    memset(&A.accounts[0], 0, sizeof(AccountRec));
    A.accounts[0].storage_src = NODE_EMPTY;
    // a base trie of n0 keys; a share of the keys get a common prefix so that extensions and deep splits occur
    const size_t n0 = 1 + rng() % 400, n1 = 1 + rng() % 120;
    const uint32_t shared_nibbles = (uint32_t)(rng() % 40);
    auto new_key = [&]() {
      uint8_t k[32];
      for (auto& b : k) b = (uint8_t)rng();
      if (rng() % 3 == 0)
        for (uint32_t i = 0; i < shared_nibbles / 2; i++) k[i] = 0xab;
      return A.add_key_bytes(k, 32);
    };
    auto payload = [&](uint32_t tag) {
      uint8_t v[8];
      memcpy(v, &tag, 4), memcpy(v + 4, &tag, 4);
      return HostArena::Payload{false, A.add_val(v, 8), 8};
    };
    uint32_t root = NODE_EMPTY;
    std::vector<uint32_t> base_keys;
    for (size_t i = 0; i < n0; i++) {
      uint32_t k = new_key();
      base_keys.push_back(k);
      root = A.insert(root, k, 64, 0, payload((uint32_t)i));
    }
    std::vector<HostArena::BatchItem> items;
    for (size_t i = 0; i < n1; i++) {
      // half new keys, half overwrites of existing ones
      uint32_t k = (rng() & 1) ? new_key() : base_keys[rng() % base_keys.size()];
      bool dup = false;
      for (auto& it : items) dup |= memcmp(A.key_pool.data() + it.koff, A.key_pool.data() + k, 32) == 0;
      if (dup) continue;
      items.push_back({k, 64, payload(1000000u + (uint32_t)i)});
    }
    // one by one, in the given order
    uint32_t seq = root;
    for (auto& it : items) seq = A.insert(seq, it.koff, it.klen, 0, it.payload);
    // batched, sorted by key
    std::vector<HostArena::BatchItem> sorted = items;
    std::sort(sorted.begin(), sorted.end(), [&](const HostArena::BatchItem& x, const HostArena::BatchItem& y) {
      return memcmp(A.key_pool.data() + x.koff, A.key_pool.data() + y.koff, 32) < 0;
    });
    const size_t nodes_before = A.nodes.size();
    uint32_t bat = A.insert_many(root, sorted.data(), 0, sorted.size(), 0);
    const size_t nodes_batched = A.nodes.size() - nodes_before;
    CHECK(same(A, seq, bat), "round %d: insert_many differs from sequential inserts (%zu base keys, %zu items)", round, n0, items.size());
    CHECK(levels_ok(A, bat), "round %d: a level of the batched trie is not above its children's", round);
    CHECK(levels_ok(A, seq), "round %d: a level of the sequential trie is not above its children's", round);
    CHECK(nodes_batched <= 70 * items.size() + 70, "round %d: batched insert created %zu nodes for %zu keys", round, nodes_batched, items.size());
    // from an empty trie too
    uint32_t seq0 = NODE_EMPTY;
    for (auto& it : items) seq0 = A.insert(seq0, it.koff, it.klen, 0, it.payload);
    CHECK(same(A, seq0, A.insert_many(NODE_EMPTY, sorted.data(), 0, sorted.size(), 0)), "round %d: insert_many into an empty trie", round);
    // marks: present keys, absent keys
    std::vector<HostArena::MarkItem> marks;
    std::vector<uint32_t> touched_seq, touched_many, leaves_seq;
    for (size_t i = 0; i < 150; i++) {
      uint32_t k = (rng() & 1) ? new_key() : sorted[rng() % sorted.size()].koff;
      marks.push_back({bat, k, 64, NODE_EMPTY});
      leaves_seq.push_back(A.mark(bat, k, 64, touched_seq));
    }
    marks.push_back({NODE_EMPTY, sorted[0].koff, 64, 7});  // an empty trie: nothing touched, no leaf
    leaves_seq.push_back(NODE_EMPTY);
    A.mark_many(marks.data(), marks.size(), touched_many);
    for (size_t i = 0; i < marks.size(); i++) {
      CHECK(marks[i].leaf == leaves_seq[i], "round %d: mark_many leaf %zu", round, i);
      CHECK(marks[i].leaf == A.get(marks[i].root, marks[i].koff, 64), "round %d: mark leaf != get, key %zu", round, i);
    }
    std::sort(touched_seq.begin(), touched_seq.end());
    std::sort(touched_many.begin(), touched_many.end());
    CHECK(touched_seq == touched_many, "round %d: mark_many touched %zu nodes, mark %zu", round, touched_many.size(), touched_seq.size());
    // branch_with against a rebuild from the 16 slots
    for (uint32_t n = 0; n < A.nodes.size() && n < 4000; n++) {
      if (A.kind(n) != NK_BRANCH) continue;
      uint32_t nib = (uint32_t)(rng() % 16), child = (rng() % 4 == 0) ? NODE_EMPTY : (uint32_t)(rng() % A.nodes.size());
      uint32_t kids[16], k = 0, mask = 0;
      for (uint32_t i = 0; i < 16; i++) {
        uint32_t c = i == nib ? child : A.child_at(n, i);
        if (c != NODE_EMPTY) kids[k++] = c, mask |= 1u << i;
      }
      if (k == 0) continue;
      uint32_t want = A.new_branch(mask, kids), got = A.branch_with(n, nib, child);
      bool eq = (A.nodes[want].a1 & 0xffff) == (A.nodes[got].a1 & 0xffff);
      for (uint32_t i = 0; i < 16 && eq; i++) eq = A.child_at(want, i) == A.child_at(got, i);
      CHECK(eq, "round %d: branch_with(%u, %u) differs from the rebuilt branch", round, n, nib);
      CHECK(A.lvl(got) >= A.lvl(want), "round %d: branch_with level %u below the exact level %u", round, A.lvl(got), A.lvl(want));
    }
    // removals after the batch: sequential semantics unchanged
    uint32_t r1 = bat, r2 = seq;
    for (size_t i = 0; i < sorted.size(); i += 3) {
      uint32_t a = A.remove(r1, sorted[i].koff, 64, 0), b = A.remove(r2, sorted[i].koff, 64, 0);
      CHECK((a == UNCHANGED) == (b == UNCHANGED), "round %d: remove presence", round);
      if (a != UNCHANGED) r1 = a;
      if (b != UNCHANGED) r2 = b;
    }
    CHECK(same(A, r1, r2), "round %d: tries differ after removals", round);
  }
  // ---- flat_maps.h against std::map under a random mix of inserts, overwrites, erasures and look-ups ----
  for (int round = 0; round < rounds; round++) {
    FlatMapU32 fm;
    std::map<uint32_t, uint32_t> ref;
    H256Map hm;
    std::map<std::vector<uint8_t>, uint32_t> href;
    if (round & 1) fm.reserve(100), hm.reserve(100);
    const size_t ops = 200 + rng() % 20000;
    auto key_of = [&](uint32_t i) {
      H256 k;
      uint64_t x = i * 0x9E3779B97F4A7C15ull + 1;
      for (int w = 0; w < 4; w++) {
        x ^= x >> 29, x *= 0xBF58476D1CE4E5B9ull;
        memcpy(k.b + 8 * w, &x, 8);
      }
      if (i % 7 == 0) memset(k.b, 0, 8);  // colliding hash prefixes
      return k;
    };
    for (size_t op = 0; op < ops; op++) {
      const uint32_t k = (uint32_t)(rng() % 3000), v = (uint32_t)(rng() % 1000000);
      const uint32_t key32 = (k % 11 == 0) ? 0xffffffffu - (k % 5) : k;  // NODE_EMPTY and its neighbours are valid keys
      const H256 hk = key_of(k);
      const std::vector<uint8_t> hv(hk.b, hk.b + 32);
      switch (rng() % 4) {
        case 0:
          fm.put(key32, v), ref[key32] = v;
          hm[hk] = v, href[hv] = v;
          break;
        case 1: {
          auto ins = hm.insert({hk, v});
          auto rins = href.insert({hv, v});
          CHECK(ins.second == rins.second && ins.first->second == rins.first->second, "H256Map::insert");
          break;
        }
        case 2:
          CHECK(hm.erase(hk) == href.erase(hv), "H256Map::erase");
          break;
        default: {
          const uint32_t* f = fm.find(key32);
          auto r = ref.find(key32);
          CHECK((f != nullptr) == (r != ref.end()) && (!f || *f == r->second), "FlatMapU32::find");
          auto hf = hm.find(hk);
          auto hr = href.find(hv);
          CHECK((hf != hm.end()) == (hr != href.end()) && (hf == hm.end() || hf->second == hr->second), "H256Map::find");
        }
      }
    }
    CHECK(fm.count == ref.size(), "FlatMapU32 size %zu / %zu", fm.count, ref.size());
    CHECK(hm.size() == href.size(), "H256Map size %zu / %zu", hm.size(), href.size());
    size_t seen = 0;
    hm.for_each([&](const H256Map::Entry& e) {
      seen++;
      auto hr = href.find(std::vector<uint8_t>(e.first.b, e.first.b + 32));
      CHECK(hr != href.end() && hr->second == e.second, "H256Map::for_each entry");
    });
    CHECK(seen == href.size(), "H256Map::for_each count");
    fm.erase_if([&](uint32_t key, uint32_t) { return key % 2 == 0; });
    for (auto it = ref.begin(); it != ref.end();) it = it->first % 2 == 0 ? ref.erase(it) : std::next(it);
    size_t left = 0;
    fm.for_each([&](uint32_t key, uint32_t val) {
      left++;
      CHECK(ref.count(key) && ref[key] == val, "FlatMapU32 after erase_if");
    });
    CHECK(left == ref.size(), "FlatMapU32::erase_if count");
    H256Map copy = hm;  // the storage-map snapshot a dummy IR reads
    hm.erase(key_of(1));
    CHECK(copy.size() == href.size(), "H256Map copy is independent");
  }
  if (failures) {
    fprintf(stderr, "%d check(s) failed\n", failures);
    return 1;
  }
  printf("host_arena_check: %d rounds ok\n", rounds);
  return 0;
}
