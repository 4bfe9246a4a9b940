// hostprof_stub.h — DEVELOPMENT ONLY.  Included by ppd_host.cu solely when it is compiled with
// -DPPD_HOSTPROF for tools/hostprof (a host-phase profiler that runs where there is no GPU).  The
// shipped libppd_b200.so is never built with that macro: it has no CPU hashing of any kind.
// In this mode node refs are NOT computed (all zero), only the byte strings whose hashes shape the
// tries (addresses, slots, code) are hashed, with this scalar Keccak-256.
#pragma once
#include <cstdint>
#include <cstring>

namespace hostprof {
inline uint64_t rol(uint64_t x, int s) { return s ? (x << s) | (x >> (64 - s)) : x; }
inline void f1600(uint64_t* a) {
  static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
                                  0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                                  0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
                                  0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int R[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int r = 0; r < 24; r++) {
    uint64_t c[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) {
      uint64_t d = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
      for (int y = 0; y < 25; y += 5) a[y + x] ^= d;
    }
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], R[x + 5 * y]);
    for (int y = 0; y < 25; y += 5)
      for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
    a[0] ^= RC[r];
  }
}
inline void keccak256(const uint8_t* p, size_t n, uint8_t out[32]) {
  uint64_t a[25] = {0};
  uint8_t blk[136];
  while (n >= 136) {
    for (int i = 0; i < 17; i++) {
      uint64_t w;
      memcpy(&w, p + 8 * i, 8);
      a[i] ^= w;
    }
    f1600(a);
    p += 136, n -= 136;
  }
  memset(blk, 0, 136);
  memcpy(blk, p, n);
  blk[n] ^= 0x01;
  blk[135] ^= 0x80;
  for (int i = 0; i < 17; i++) {
    uint64_t w;
    memcpy(&w, blk + 8 * i, 8);
    a[i] ^= w;
  }
  f1600(a);
  memcpy(out, a, 32);
}
}  // namespace hostprof
