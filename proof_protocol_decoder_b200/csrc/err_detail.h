// err_detail.h — what ppd_last_error says for the TraceParsingError statuses (21-25): a sentence, then "; " and the
// variant's payload as key=value words, so that the Rust shim can rebuild the
// reference's error VALUE, not only its variant (decoding.rs:31-49):
//   21 AccountDecode(hex bytes, rlp error)                 ; bytes=<hex>          (the shim re-derives the rlp error text)
//   22 MissingAccountStorageTrie(HashedAccountAddr)        ; hashed_addr=<64 hex>
//   24 MissingKeysCreatingSubPartialTrie(TrieType)         ; trie_type=State|Storage|Receipt|Txn
//   25 MissingWithdrawalAccount(Address, hashed, U256)     ; addr=<40 hex> hashed_addr=<64 hex> amount=<64 hex>
// Host only, no dependencies: tests/cpp/err_detail_check.cpp compiles it alone and compares with the oracle's wording.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace ppd {

inline std::string hex_of(const uint8_t* p, size_t n) {
  static const char d[] = "0123456789abcdef";
  std::string s(2 * n, '0');
  for (size_t i = 0; i < n; i++) s[2 * i] = d[p[i] >> 4], s[2 * i + 1] = d[p[i] & 15];
  return s;
}

// decoding.rs:51-57: the order the reference declares TrieType in
enum TrieTypeName { TRIE_STATE = 0, TRIE_STORAGE = 1, TRIE_RECEIPT = 2, TRIE_TXN = 3 };
inline const char* trie_type_name(int t) {
  static const char* const names[] = {"State", "Storage", "Receipt", "Txn"};
  return names[t & 3];
}

inline std::string detail_account_decode(const char* what, const uint8_t* bytes, size_t n) {
  return std::string(what) + "; bytes=" + hex_of(bytes, n);
}
inline std::string detail_missing_storage_trie(const char* what, const uint8_t hashed_addr[32]) {
  return std::string(what) + "; hashed_addr=" + hex_of(hashed_addr, 32);
}
inline std::string detail_missing_keys(int trie_type) {
  return std::string("subset key runs into a hashed-out node; trie_type=") + trie_type_name(trie_type);
}
inline std::string detail_missing_withdrawal_account(const uint8_t addr[20], const uint8_t hashed_addr[32], const uint8_t amount_be[32]) {
  return "withdrawal to an account that is not in the state trie; addr=" + hex_of(addr, 20) + " hashed_addr=" + hex_of(hashed_addr, 32) +
         " amount=" + hex_of(amount_be, 32);
}

}  // namespace ppd
