// ppd_build.cu — trie construction ON THE GPU from leaves sorted by key, and the root computation
// over the resulting structure-of-arrays arena (config 5: full state-trie rehash; also the shape
// of every large storage trie).
//
// Replaces: building a HashedPartialTrie by repeated `insert` (eth_trie_utils trie_ops.rs; call
// sites compact_to_partial_trie.rs:105,125) followed by the recursive `hash()`.
//
// With the keys sorted, the Patricia trie is determined by the array L of longest common nibble
// prefixes of adjacent keys (L[i] = lcp(key[i-1], key[i]); L[0] = L[N] = -1):
//   * a branch node at depth d  <=>  a maximal run of positions q with L[q] >= d that contains at
//     least one L[q] == d; its "leader" is the leftmost position of the run with L == d;
//   * the parent of a leaf i is the branch owning position i or i+1, whichever has the larger L;
//   * the parent of a branch spanning items [l, r] is the branch owning position l or r+1,
//     whichever has the larger L; if that depth Dp < d-1 the d-Dp-1 nibbles in between form an
//     extension node, which is encoded and hashed by the same thread right after its branch.
// "Nearest position to the left/right with a smaller L" queries run on a 3-level min-pyramid
// (64 / 4096 / 262144 positions per cell), so every thread does O(64 * levels) byte reads at worst.
//
// Arena (structure of arrays, all in HBM):
//   leaves   id i in [0, N):        keys32[i], vals[val_off[i]..val_off[i+1])
//   branches id N + b, b in [0, B): depth[b], child[b][16] (ids or NODE_EMPTY), ext_start[b],
//                                   rep[b] (an item below it: its key spells the extension nibbles)
//   ref[id][32] + ref_len[id]: what the parent embeds.
// Branches are counting-sorted by depth; one launch per non-empty depth, deepest first.
#include <cstdint>

#include "arena.h"
#include "keccak.cuh"
#include "ppd_kernels.h"

namespace ppd {

// ------------------------------------------------------------------ helpers (as ppd_kernels.cu)
namespace b {

template <int B>
__device__ __forceinline__ void emit_len_prefix(Stage<B>& s, uint32_t len, uint32_t short_base, uint32_t long_base) {
  if (len < 56) {
    s.put_byte(short_base + len);
  } else if (len < 256) {
    s.put_byte(long_base + 1);
    s.put_byte(len);
  } else if (len < 65536) {
    s.put_byte(long_base + 2);
    s.put_byte(len >> 8);
    s.put_byte(len & 255);
  } else {
    s.put_byte(long_base + 3);
    s.put_byte(len >> 16);
    s.put_byte((len >> 8) & 255);
    s.put_byte(len & 255);
  }
}
__device__ __forceinline__ uint32_t len_prefix_size(uint32_t len) { return len < 56 ? 1 : len < 256 ? 2 : len < 65536 ? 3 : 4; }
__device__ __forceinline__ uint32_t hex_prefix_str_size(uint32_t n) { return n < 2 ? 1 : 2 + (n >> 1); }

// rlp_str(hex_prefix(nibbles [start, start+n) of a 32-byte key held in 8 big-endian-packed words))
template <int B>
__device__ __forceinline__ void emit_hex_prefix_str(Stage<B>& s, const uint8_t* key, uint32_t start, uint32_t n, uint32_t is_leaf) {
  if (n >= 2) s.put_byte(0x80 + 1 + (n >> 1));
  uint32_t flag = (is_leaf ? 2u : 0u) + (n & 1);
  uint32_t j = start, end = start + n;
  if (n & 1) {
    uint32_t bb = __ldg(key + (j >> 1));
    s.put_byte((flag << 4) | ((j & 1) ? (bb & 15) : (bb >> 4)));
    j++;
  } else {
    s.put_byte(flag << 4);
  }
  if ((j & 1) == 0) {
    for (; j < end; j += 2) s.put_byte(__ldg(key + (j >> 1)));
  } else {
    uint32_t prev = __ldg(key + (j >> 1));
    for (; j < end; j += 2) {
      uint32_t next = __ldg(key + (j >> 1) + 1);
      s.put_byte(((prev & 15) << 4) | (next >> 4));
      prev = next;
    }
  }
}

// up to 32 bytes from any global address
template <int B>
__device__ __forceinline__ void emit_chunk(Stage<B>& s, const uint8_t* p, uint32_t len) {
  if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    uint32_t nw = len >> 2;
    for (uint32_t i = 0; i < nw; i++) s.put_word(__ldg(w + i));
    uint32_t rem = len & 3;
    if (rem) {
      uint32_t x = 0;
      for (uint32_t k = 0; k < rem; k++) x |= (uint32_t)__ldg(p + 4 * nw + k) << (8 * k);
      s.put_partial(x, rem);
    }
  } else {
    for (uint32_t i = 0; i < len; i++) s.put_byte(__ldg(p + i));
  }
}

template <int B>
__device__ __forceinline__ void emit_ref_words(Stage<B>& s, const uint32_t (&w)[8], uint32_t len) {
  if (len == 32) {
    s.put_byte(0xa0);
#pragma unroll
    for (int i = 0; i < 8; i++) s.put_word(w[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t take = len > 4u * i ? min(4u, len - 4u * i) : 0u;
      s.put_partial(w[i], take);
    }
  }
}

}  // namespace b

// ------------------------------------------------------------------ structure kernels ---------

__device__ __forceinline__ int lcp_nibbles(const uint4* a, const uint4* bq, int* cmp) {
  // 32-byte big-endian keys; returns common nibble prefix and sign of (a - b) through *cmp
  const uint32_t* x = reinterpret_cast<const uint32_t*>(a);
  const uint32_t* y = reinterpret_cast<const uint32_t*>(bq);
  for (int w = 0; w < 8; w++) {
    uint32_t xv = __byte_perm(x[w], 0, 0x0123), yv = __byte_perm(y[w], 0, 0x0123);  // to big-endian order
    uint32_t diff = xv ^ yv;
    if (diff) {
      *cmp = xv < yv ? -1 : 1;
      return 8 * w + (__clz(diff) >> 2);
    }
  }
  *cmp = 0;
  return 64;
}

// L[0] = L[N] = -1;  L[i] = lcp(key[i-1], key[i]).  flags[0] |= 1 if not strictly ascending.
__global__ void lcp_kernel(const uint8_t* __restrict__ keys, uint32_t n, int8_t* __restrict__ L, uint32_t* __restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == 0 || i == n) {
    L[i] = -1;
    return;
  }
  uint4 ka[2], kb[2];
  const uint4* pa = reinterpret_cast<const uint4*>(keys + 32ull * (i - 1));
  const uint4* pb = reinterpret_cast<const uint4*>(keys + 32ull * i);
  ka[0] = __ldg(pa), ka[1] = __ldg(pa + 1), kb[0] = __ldg(pb), kb[1] = __ldg(pb + 1);
  int cmp;
  int l = lcp_nibbles(ka, kb, &cmp);
  if (cmp >= 0) atomicOr(flags, 1u);
  L[i] = (int8_t)(l > 63 ? 63 : l);
}

// out[j] = min(in[64 j .. 64 j + 63]) over the valid entries
__global__ void min64_kernel(const int8_t* __restrict__ in, uint32_t n_in, int8_t* __restrict__ out, uint32_t n_out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  int m = 127;
  uint64_t base = 64ull * j;
  for (uint32_t k = 0; k < 64 && base + k < n_in; k++) m = min(m, (int)in[base + k]);
  out[j] = (int8_t)m;
}

// largest p < q with L[p] < thr (thr >= 0; L[0] = -1 guarantees termination)
__device__ __forceinline__ uint32_t scan_left(const Pyramid& P, uint32_t q, int thr) {
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
__device__ __forceinline__ uint32_t scan_right(const Pyramid& P, uint32_t q, int thr) {
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

// link[q] = nearest position to the left in the same branch run with the same depth, else q
__global__ void link_kernel(Pyramid P, uint32_t n, uint32_t* __restrict__ link) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0 || q >= n) {
    if (q == 0 || q == n) link[q] = q;
    return;
  }
  int d = P.L[q];
  uint32_t p = scan_left(P, q, d + 1);  // nearest with L <= d
  link[q] = (P.L[p] == d) ? p : q;
}
__global__ void jump_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n_plus_1) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_plus_1) return;
  out[q] = in[in[q]];
}
__global__ void leader_flag_kernel(const uint32_t* __restrict__ leader, uint32_t n, uint32_t* __restrict__ flag) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > n) return;
  flag[q] = (q >= 1 && q < n && leader[q] == q) ? 1u : 0u;
}

// ---- exclusive prefix sum of uint32 (3-phase, recursive on the block sums) ----------------------
static constexpr int SCAN_B = 256, SCAN_ITEMS = 4;  // 1024 elements per block

__global__ void scan_block_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t warp_sums[SCAN_B / 32];
  uint32_t base = blockIdx.x * (SCAN_B * SCAN_ITEMS) + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = base + k < n ? in[base + k] : 0u;
    sum += v[k];
  }
  uint32_t incl = sum;
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
    if ((threadIdx.x & 31) >= off) incl += t;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = threadIdx.x < SCAN_B / 32 ? warp_sums[threadIdx.x] : 0u;
    uint32_t wi = w;
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, off);
      if (threadIdx.x >= off) wi += t;
    }
    if (threadIdx.x < SCAN_B / 32) warp_sums[threadIdx.x] = wi - w;  // exclusive
    if (threadIdx.x == SCAN_B / 32 - 1 && block_sums) block_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t excl = incl - sum + warp_sums[threadIdx.x >> 5];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
}
__global__ void scan_add_kernel(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ block_offsets) {
  uint32_t i = blockIdx.x * (SCAN_B * SCAN_ITEMS) + threadIdx.x;
  uint32_t add = block_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    uint32_t j = i + k * SCAN_B;
    if (j < n) out[j] += add;
  }
}

// tmp must hold at least scan_tmp_words(n) uint32
size_t scan_tmp_words(size_t n) {
  size_t total = 0;
  while (n > 1) {
    n = (n + SCAN_B * SCAN_ITEMS - 1) / (SCAN_B * SCAN_ITEMS);
    total += n + 1;
  }
  return total + 2;
}
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t st) {
  if (!n) return;
  uint32_t nb = (n + SCAN_B * SCAN_ITEMS - 1) / (SCAN_B * SCAN_ITEMS);
  scan_block_kernel<<<nb, SCAN_B, 0, st>>>(in, out, n, nb > 1 ? tmp : nullptr);
  if (nb > 1) {
    exclusive_scan_u32(tmp, tmp, nb, tmp + nb + 1, st);
    scan_add_kernel<<<nb, SCAN_B, 0, st>>>(out, n, tmp);
  }
}


// per leader: depth, extension, parent link
__global__ void branch_info_kernel(BuildView V) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0 || q >= V.n) return;
  if (V.leader[q] != q) return;
  const Pyramid& P = V.P;
  uint32_t bi = V.bidx[q];
  int d = P.L[q];
  uint32_t l = scan_left(P, q, d);    // L[l] < d : item l is the first item of the branch
  uint32_t r1 = scan_right(P, q, d);  // L[r1] < d : item r1 - 1 is the last
  int dl = P.L[l], dr = P.L[r1];
  int dp = max(dl, dr);
  V.depth[bi] = (uint8_t)d;
  V.rep[bi] = l;
  V.ext_start[bi] = (uint8_t)(dp + 1);
  if (dp < 0) {
    *V.root_id = V.n + bi;
  } else {
    uint32_t pl = (dl >= dr) ? l : r1;
    uint32_t pb = V.bidx[V.leader[pl]];
    uint32_t b = __ldg(V.keys + 32ull * l + (dp >> 1));
    uint32_t nib = (dp & 1) ? (b & 15u) : (b >> 4);
    V.child[16ull * pb + nib] = V.n + bi;
  }
}

// hist[d] = number of branches at depth d (block-local histogram, one global atomic per bin and block)
__global__ void depth_hist_kernel(const uint8_t* __restrict__ depth, uint32_t nb, uint32_t* __restrict__ hist_out) {
  __shared__ uint32_t hist[64];
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb) atomicAdd(&hist[depth[b]], 1u);
  __syncthreads();
  if (threadIdx.x < 64 && hist[threadIdx.x]) atomicAdd(hist_out + threadIdx.x, hist[threadIdx.x]);
}

// counting sort of the branches by depth: order[level_start[d] + k] = b
__global__ void branch_scatter_kernel(const uint8_t* __restrict__ depth, uint32_t nb, uint32_t* __restrict__ cursor,
                                      uint32_t* __restrict__ order) {
  __shared__ uint32_t hist[64], base[64];
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t d = 0, rank = 0;
  if (b < nb) {
    d = depth[b];
    rank = atomicAdd(&hist[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 64 && hist[threadIdx.x]) base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, hist[threadIdx.x]);
  __syncthreads();
  if (b < nb) order[base[d] + rank] = b;
}

// ------------------------------------------------------------------ hashing kernels -----------

template <int B>
__global__ void __launch_bounds__(B) hash_sorted_leaves_kernel(BuildView V) {
  __shared__ uint32_t smem[PPD_STAGE_WORDS * B];
  uint32_t i = blockIdx.x * B + threadIdx.x;
  uint32_t hashed = 0, perms = 0, enc_bytes = 0;
  if (i < V.n) {
    const Pyramid& P = V.P;
    int dl = P.L[i], dr = P.L[i + 1];
    int dp = max(dl, dr);
    const uint8_t* key = V.keys + 32ull * i;
    const bool is_root = dp < 0;
    if (is_root) {
      *V.root_id = i;
    } else {
      uint32_t pl = (dl >= dr) ? i : i + 1;
      uint32_t pb = V.bidx[V.leader[pl]];
      uint32_t kb = __ldg(key + (dp >> 1));
      uint32_t nib = (dp & 1) ? (kb & 15u) : (kb >> 4);
      V.child[16ull * pb + nib] = i;
    }
    const uint32_t nib_start = (uint32_t)(dp + 1), nib_len = 64 - nib_start;
    const uint64_t vo = V.val_off[i];
    const uint8_t* val = V.vals + vo;
    const uint32_t vlen = (uint32_t)(V.val_off[i + 1] - vo);
    const uint32_t vhdr = (vlen == 1 && __ldg(val) < 0x80) ? 0 : b::len_prefix_size(vlen);
    const uint32_t payload = b::hex_prefix_str_size(nib_len) + vhdr + vlen;
    const uint32_t total = b::len_prefix_size(payload) + payload;
    const uint32_t nseg = 2 + ((vlen + 31) >> 5);
    const bool inline_ref = !is_root && total < 32;
    Stage<B> s;
    s.init(smem);
    uint64_t a[25];
#pragma unroll
    for (int k = 0; k < 25; k++) a[k] = 0;
    uint32_t seg = 0;
    bool done = false;
    while (!done) {
      while (s.bytes() < 136 && seg < nseg) {
        if (seg == 0) {
          b::emit_len_prefix(s, payload, 0xc0, 0xf7);
          b::emit_hex_prefix_str(s, key, nib_start, nib_len, 1);
        } else if (seg == 1) {
          if (vhdr) b::emit_len_prefix(s, vlen, 0x80, 0xb7);
        } else {
          uint32_t off = (seg - 2) << 5;
          b::emit_chunk(s, val + off, min(32u, vlen - off));
        }
        seg++;
      }
      if (inline_ref) break;
      if (s.bytes() < 136) {
        s.pad();
        done = true;
      }
      absorb_stage<B>(a, s.w);
      keccak_f1600(a);
      perms++;
      if (!done) s.consume_block();
    }
    uint4* o = reinterpret_cast<uint4*>(V.ref + 32ull * i);
    if (inline_ref) {
      s.flush_partial();
      uint32_t nw = (total + 3) >> 2;
      uint32_t h[8];
#pragma unroll
      for (int k = 0; k < 8; k++) h[k] = (uint32_t)k < nw ? s.w[k * B] : 0u;
      uint32_t tail = total & 3;
      if (tail) h[nw - 1] &= (1u << (8 * tail)) - 1;
      o[0] = make_uint4(h[0], h[1], h[2], h[3]);
      o[1] = make_uint4(h[4], h[5], h[6], h[7]);
      V.ref_len[i] = (uint8_t)total;
    } else {
      uint4 x = make_uint4((uint32_t)a[0], (uint32_t)(a[0] >> 32), (uint32_t)a[1], (uint32_t)(a[1] >> 32));
      uint4 y = make_uint4((uint32_t)a[2], (uint32_t)(a[2] >> 32), (uint32_t)a[3], (uint32_t)(a[3] >> 32));
      o[0] = x, o[1] = y;
      V.ref_len[i] = 32;
      hashed = 1;
      enc_bytes = total;
      if (is_root) {
        uint4* ro = reinterpret_cast<uint4*>(V.root_out);
        ro[0] = x, ro[1] = y;
      }
    }
  }
  if (V.counters) {
    for (int off = 16; off > 0; off >>= 1) {
      hashed += __shfl_down_sync(0xffffffffu, hashed, off);
      perms += __shfl_down_sync(0xffffffffu, perms, off);
      enc_bytes += __shfl_down_sync(0xffffffffu, enc_bytes, off);
    }
    if ((threadIdx.x & 31) == 0 && hashed) {
      atomicAdd(V.counters + 0, (unsigned long long)hashed);
      atomicAdd(V.counters + 1, (unsigned long long)perms);
      atomicAdd(V.counters + 2, (unsigned long long)enc_bytes);
    }
  }
}

// One level of branches (all at the same depth), each followed by its extension node if it has one.
template <int B>
__global__ void __launch_bounds__(B) hash_branch_level_kernel(BuildView V, const uint32_t* __restrict__ order, uint32_t begin, uint32_t end) {
  __shared__ uint32_t smem[PPD_STAGE_WORDS * B];
  uint32_t slot = begin + blockIdx.x * B + threadIdx.x;
  uint32_t hashed = 0, perms = 0, enc_bytes = 0;
  if (slot < end) {
    const uint32_t bi = __ldg(order + slot);
    const uint32_t id = V.n + bi;
    const uint32_t d = V.depth[bi], es = V.ext_start[bi];
    const uint32_t ext_len = d - es;
    const bool is_root = (*V.root_id == id);
    uint32_t kid[16];
    {
      const uint4* ct = reinterpret_cast<const uint4*>(V.child + 16ull * bi);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint4 t = __ldcg(ct + k);
        kid[4 * k] = t.x, kid[4 * k + 1] = t.y, kid[4 * k + 2] = t.z, kid[4 * k + 3] = t.w;
      }
    }
    uint32_t payload = 1;  // the empty branch value
#pragma unroll
    for (int k = 0; k < 16; k++) {
      uint32_t cl = kid[k] == NODE_EMPTY ? 0u : (uint32_t)V.ref_len[kid[k]];
      payload += kid[k] == NODE_EMPTY ? 1u : (cl == 32 ? 33u : cl);
    }
    uint32_t total = b::len_prefix_size(payload) + payload;
    Stage<B> s;
    s.init(smem);
    uint64_t a[25];
#pragma unroll
    for (int k = 0; k < 25; k++) a[k] = 0;
    uint32_t seg = 0, nseg = 18, phase = 0;
    uint32_t rw[8];  // ref of the message just finished
    uint32_t rlen = 0;
    for (;;) {
      if (phase == 0) {
        while (s.bytes() < 136 && seg < nseg) {
          if (seg == 0) {
            b::emit_len_prefix(s, payload, 0xc0, 0xf7);
          } else if (seg == 17) {
            s.put_byte(0x80);
          } else {
            uint32_t c = kid[0];
#pragma unroll
            for (int k = 1; k < 16; k++) c = (seg - 1 == (uint32_t)k) ? kid[k] : c;
            if (c == NODE_EMPTY) {
              s.put_byte(0x80);
            } else {
              const uint4* q = reinterpret_cast<const uint4*>(V.ref + 32ull * c);
              uint4 x = __ldcg(q), y = __ldcg(q + 1);
              uint32_t w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
              b::emit_ref_words(s, w, V.ref_len[c]);
            }
          }
          seg++;
        }
      }
      const bool final_msg = (phase == 1) || ext_len == 0;
      const bool want_inline = total < 32 && !(final_msg && is_root);
      bool last = true;
      if (!want_inline) {
        last = s.bytes() < 136;
        if (last) s.pad();
        absorb_stage<B>(a, s.w);
        keccak_f1600(a);
        perms++;
        if (!last) {
          s.consume_block();
          continue;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) rw[2 * k] = (uint32_t)a[k], rw[2 * k + 1] = (uint32_t)(a[k] >> 32);
        rlen = 32;
        hashed++;
        enc_bytes += total;
      } else {
        s.flush_partial();
        uint32_t nw = (total + 3) >> 2;
#pragma unroll
        for (int k = 0; k < 8; k++) rw[k] = (uint32_t)k < nw ? s.w[k * B] : 0u;
        uint32_t tail = total & 3;
        if (tail) rw[nw - 1] &= (1u << (8 * tail)) - 1;
        rlen = total;
      }
      if (final_msg) break;
      // the extension node above this branch: rlp[ hex_prefix(nibbles, false), ref ]
      phase = 1;
#pragma unroll
      for (int k = 0; k < 25; k++) a[k] = 0;
      s.init(smem);
      payload = b::hex_prefix_str_size(ext_len) + (rlen == 32 ? 33u : rlen);
      total = b::len_prefix_size(payload) + payload;
      b::emit_len_prefix(s, payload, 0xc0, 0xf7);
      b::emit_hex_prefix_str(s, V.keys + 32ull * V.rep[bi], es, ext_len, 0);
      b::emit_ref_words(s, rw, rlen);
    }
    uint4* o = reinterpret_cast<uint4*>(V.ref + 32ull * id);
    uint4 x = make_uint4(rw[0], rw[1], rw[2], rw[3]), y = make_uint4(rw[4], rw[5], rw[6], rw[7]);
    o[0] = x, o[1] = y;
    V.ref_len[id] = (uint8_t)rlen;
    if (is_root) {
      uint4* ro = reinterpret_cast<uint4*>(V.root_out);
      ro[0] = x, ro[1] = y;
    }
  }
  if (V.counters) {
    for (int off = 16; off > 0; off >>= 1) {
      hashed += __shfl_down_sync(0xffffffffu, hashed, off);
      perms += __shfl_down_sync(0xffffffffu, perms, off);
      enc_bytes += __shfl_down_sync(0xffffffffu, enc_bytes, off);
    }
    if ((threadIdx.x & 31) == 0 && hashed) {
      atomicAdd(V.counters + 0, (unsigned long long)hashed);
      atomicAdd(V.counters + 1, (unsigned long long)perms);
      atomicAdd(V.counters + 2, (unsigned long long)enc_bytes);
    }
  }
}

// ------------------------------------------------------------------ launchers -----------------

static constexpr int HB = 128;
static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

void launch_lcp(const uint8_t* keys, uint32_t n, int8_t* L, uint32_t* flags, cudaStream_t st) {
  lcp_kernel<<<cdiv((uint64_t)n + 1, 256), 256, 0, st>>>(keys, n, L, flags);
}
void launch_min64(const int8_t* in, uint32_t n_in, int8_t* out, uint32_t n_out, cudaStream_t st) {
  min64_kernel<<<cdiv(n_out, 256), 256, 0, st>>>(in, n_in, out, n_out);
}
void launch_leaders(const int8_t* L, const int8_t* m1, const int8_t* m2, const int8_t* m3, uint32_t n, uint32_t* link_a, uint32_t* link_b,
                    uint32_t* flag, cudaStream_t st) {
  Pyramid P{L, m1, m2, m3};
  uint32_t g = cdiv((uint64_t)n + 1, 256);
  link_kernel<<<g, 256, 0, st>>>(P, n, link_a);
  // a branch has at most 16 children, so a chain has at most 15 links: 4 doublings
  jump_kernel<<<g, 256, 0, st>>>(link_a, link_b, n + 1);
  jump_kernel<<<g, 256, 0, st>>>(link_b, link_a, n + 1);
  jump_kernel<<<g, 256, 0, st>>>(link_a, link_b, n + 1);
  jump_kernel<<<g, 256, 0, st>>>(link_b, link_a, n + 1);
  leader_flag_kernel<<<g, 256, 0, st>>>(link_a, n, flag);
}
void launch_branch_info(const BuildView& V, cudaStream_t st) {
  if (V.n < 2) return;
  branch_info_kernel<<<cdiv(V.n, 256), 256, 0, st>>>(V);
}
void launch_depth_hist(const uint8_t* depth, uint32_t nb, uint32_t* hist, cudaStream_t st) {
  if (!nb) return;
  depth_hist_kernel<<<cdiv(nb, 256), 256, 0, st>>>(depth, nb, hist);
}
void launch_branch_scatter(const uint8_t* depth, uint32_t nb, uint32_t* cursor, uint32_t* order, cudaStream_t st) {
  if (!nb) return;
  branch_scatter_kernel<<<cdiv(nb, 256), 256, 0, st>>>(depth, nb, cursor, order);
}
void launch_hash_sorted_leaves(const BuildView& V, cudaStream_t st) {
  if (!V.n) return;
  hash_sorted_leaves_kernel<HB><<<cdiv(V.n, HB), HB, 0, st>>>(V);
}
void launch_hash_branch_level(const BuildView& V, const uint32_t* order, uint32_t begin, uint32_t end, cudaStream_t st) {
  if (end <= begin) return;
  hash_branch_level_kernel<HB><<<cdiv(end - begin, HB), HB, 0, st>>>(V, order, begin, end);
}

}  // namespace ppd
