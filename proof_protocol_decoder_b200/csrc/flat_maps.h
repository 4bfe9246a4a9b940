// flat_maps.h — the host pipeline's hash maps: open addressing, no heap allocation per entry.
//
// The txn loop (decoding.rs:80-177 as shaped by ppd_host.cu) looks up an account's storage trie once per touched
// account per txn and asks for the NK_ROOT node of about a hundred new trie versions per txn.  With node-based
// std::unordered_map that is one malloc per entry, on every host thread at once; these two maps keep keys and values
// inline.  Pure C++ (unit-checked on the CPU by tests/cpp/host_arena_check.cpp).
#pragma once
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace ppd {

struct H256 {
  uint8_t b[32];
  bool operator==(const H256& o) const { return memcmp(b, o.b, 32) == 0; }
  bool operator<(const H256& o) const { return memcmp(b, o.b, 32) < 0; }
};
struct H256Hasher {
  size_t operator()(const H256& h) const {
    size_t v;
    memcpy(&v, h.b + 8, sizeof v);
    return v;
  }
};

// uint32 -> uint32 map without a heap allocation per entry (open addressing, linear probing).  The txn loop asks for
// the NK_ROOT node of about a hundred new trie versions per txn; with node-based maps that is one malloc each, on
// every host thread at once.
struct FlatMapU32 {
  static constexpr uint64_t EMPTY_SLOT = ~0ull;  // (key 0xffffffff, value 0xffffffff): values are node ids, never 0xffffffff
  std::vector<uint64_t> slots;
  size_t count = 0;
  static size_t hash(uint32_t k) { return (size_t)(k * 0x9E3779B1u); }
  void clear() {
    slots.clear();
    count = 0;
  }
  void rehash(size_t cap) {  // cap: a power of two
    std::vector<uint64_t> old;
    old.swap(slots);
    slots.assign(cap, EMPTY_SLOT);
    count = 0;
    for (uint64_t e : old)
      if (e != EMPTY_SLOT) put((uint32_t)(e >> 32), (uint32_t)e);
  }
  void reserve(size_t n) {
    size_t cap = 64;
    while (cap < 2 * n) cap <<= 1;
    if (cap > slots.size()) rehash(cap);
  }
  const uint32_t* find(uint32_t key) const {
    if (slots.empty()) return nullptr;
    const size_t m = slots.size() - 1;
    for (size_t i = hash(key) & m;; i = (i + 1) & m) {
      const uint64_t e = slots[i];
      if (e == EMPTY_SLOT) return nullptr;
      if ((uint32_t)(e >> 32) == key) return reinterpret_cast<const uint32_t*>(&slots[i]);  // little-endian: the value is the low word
    }
  }
  void put(uint32_t key, uint32_t val) {
    if (2 * (count + 1) > slots.size()) rehash(slots.empty() ? 64 : slots.size() * 2);
    const size_t m = slots.size() - 1;
    for (size_t i = hash(key) & m;; i = (i + 1) & m) {
      const uint64_t e = slots[i];
      if (e == EMPTY_SLOT || (uint32_t)(e >> 32) == key) {
        count += e == EMPTY_SLOT;
        slots[i] = ((uint64_t)key << 32) | val;
        return;
      }
    }
  }
  template <class Pred>
  void erase_if(Pred pred) {  // rare (a witness rebuilt from its items): rebuild without the matching entries
    std::vector<uint64_t> old;
    old.swap(slots);
    slots.assign(old.size(), EMPTY_SLOT);
    count = 0;
    for (uint64_t e : old)
      if (e != EMPTY_SLOT && !pred((uint32_t)(e >> 32), (uint32_t)e)) put((uint32_t)(e >> 32), (uint32_t)e);
  }
  template <class F>
  void for_each(F f) const {
    for (uint64_t e : slots)
      if (e != EMPTY_SLOT) f((uint32_t)(e >> 32), (uint32_t)e);
  }
};

// H256 -> uint32 map with inline keys (open addressing, linear probing, tombstones).  Keys are Keccak digests, so
// their first eight bytes are the hash.  find() / end() / insert() / erase() / operator[] follow std::unordered_map
// closely enough for the call sites; iteration goes through for_each().
struct H256Map {
  struct Entry {
    H256 first;
    uint32_t second;
    uint32_t state;  // 0 empty, 1 full, 2 deleted
  };
  std::vector<Entry> slots;
  size_t n_full = 0, n_used = 0;  // n_used counts deleted slots too
  static size_t hash(const H256& k) {
    uint64_t h;
    memcpy(&h, k.b, 8);
    return (size_t)(h * 0x9E3779B97F4A7C15ull >> 17);
  }
  size_t size() const { return n_full; }
  Entry* end() const { return nullptr; }
  void rehash(size_t cap) {
    std::vector<Entry> old;
    old.swap(slots);
    slots.assign(cap, Entry{H256{}, 0, 0});
    n_full = n_used = 0;
    for (const Entry& e : old)
      if (e.state == 1) insert({e.first, e.second});
  }
  void reserve(size_t n) {
    size_t cap = 64;
    while (cap < 2 * n) cap <<= 1;
    if (cap > slots.size()) rehash(cap);
  }
  Entry* find(const H256& k) const {
    if (slots.empty()) return nullptr;
    const size_t m = slots.size() - 1;
    for (size_t i = hash(k) & m;; i = (i + 1) & m) {
      const Entry& e = slots[i];
      if (e.state == 0) return nullptr;
      if (e.state == 1 && e.first == k) return const_cast<Entry*>(&e);
    }
  }
  size_t count(const H256& k) const { return find(k) ? 1 : 0; }
  std::pair<Entry*, bool> insert(const std::pair<H256, uint32_t>& kv) {
    if (Entry* f = find(kv.first)) return {f, false};
    if (2 * (n_used + 1) > slots.size()) rehash(slots.empty() ? 64 : (n_full * 4 > slots.size() ? slots.size() * 2 : slots.size()));
    const size_t m = slots.size() - 1;
    for (size_t i = hash(kv.first) & m;; i = (i + 1) & m) {
      Entry& e = slots[i];
      if (e.state != 1) {
        n_used += e.state == 0;
        e.first = kv.first, e.second = kv.second, e.state = 1;
        n_full++;
        return {&e, true};
      }
    }
  }
  uint32_t& operator[](const H256& k) { return insert({k, 0u}).first->second; }
  size_t erase(const H256& k) {
    Entry* f = find(k);
    if (!f) return 0;
    f->state = 2;
    n_full--;
    return 1;
  }
  template <class F>
  void for_each(F f) const {
    for (const Entry& e : slots)
      if (e.state == 1) f(e);
  }
};

}  // namespace ppd
