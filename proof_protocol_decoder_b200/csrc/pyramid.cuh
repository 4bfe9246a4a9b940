// pyramid.cuh — nearest-smaller-value queries over an integer array (int8 `Pyramid`, int16 `Pyramid16`) through a 3-level min-pyramid
// (64 / 4 096 / 262 144 positions per cell).  Shared by the sorted-leaf trie builder (ppd_build.cu:
// branch runs of the LCP array) and the witness parser (ppd_parse.cu: the parent of a post-order
// instruction is the next instruction whose stack height is not larger).
#pragma once
#include <cstdint>

#include "ppd_kernels.h"

namespace ppd {

// largest p < q with L[p] < thr (thr >= 0; L[0] = -1 guarantees termination)
template <class PyramidT>
static __device__ __forceinline__ uint32_t scan_left(const PyramidT& P, uint32_t q, int thr) {
  uint32_t p = q - 1;
  for (;;) {
    if ((p & 63u) == 63u) {
      if ((p & 4095u) == 4095u) {
        if ((p & 262143u) == 262143u && P.m3[p >> 18] >= thr) {
          p -= 262144u;
          continue;
        }
        if (P.m2[p >> 12] >= thr) {
          p -= 4096u;
          continue;
        }
      }
      if (P.m1[p >> 6] >= thr) {
        p -= 64u;
        continue;
      }
    }
    if (P.L[p] < thr) return p;
    p--;
  }
}
// smallest r > q with L[r] < thr (L[N] = -1 guarantees termination)
template <class PyramidT>
static __device__ __forceinline__ uint32_t scan_right(const PyramidT& P, uint32_t q, int thr) {
  uint32_t p = q + 1;
  for (;;) {
    if ((p & 63u) == 0u) {
      if ((p & 4095u) == 0u) {
        if ((p & 262143u) == 0u && P.m3[p >> 18] >= thr) {
          p += 262144u;
          continue;
        }
        if (P.m2[p >> 12] >= thr) {
          p += 4096u;
          continue;
        }
      }
      if (P.m1[p >> 6] >= thr) {
        p += 64u;
        continue;
      }
    }
    if (P.L[p] < thr) return p;
    p++;
  }
}

}  // namespace ppd
