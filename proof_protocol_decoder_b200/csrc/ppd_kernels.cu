// ppd_kernels.cu — sm_100a kernels of the hot path.
//
//   keccak256_batch_kernel   utils.rs:11-13 `hash` over a batch of messages (account addresses,
//                            storage slots, contract code: processed_block_trace.rs:219,234,277;
//                            decoding.rs:235,411; compact_to_partial_trie.rs:87,147)
//   hash_level_kernel        one level of the bottom-up root computation: per node, RLP-encode it
//                            straight into a Keccak sponge (eth_trie_utils Node::hash_intern,
//                            SURVEY.md 3.3) and store its "ref" for the parent level.
//
// Both are thread-per-sponge: the state stays in registers, the 136-byte rate block is staged in
// a conflict-free shared-memory slice owned by the thread (keccak.cuh).  Each kernel has exactly
// one keccak_f1600 site inside a loop over rate blocks: filling the next block is cheap and may
// diverge, the permutation (98 % of the instructions) is executed convergently by the warp.
#include <cstdint>

#include "arena.h"
#include "keccak.cuh"
#include "ppd_kernels.h"

namespace ppd {

// ------------------------------------------------------------------ byte-stream helpers ------

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

__device__ __forceinline__ uint32_t key_nibble(const uint8_t* key, uint32_t i) {
  uint32_t b = __ldg(key + (i >> 1));
  return (i & 1) ? (b & 15) : (b >> 4);
}
// size of rlp_str(hex_prefix(nibbles)) for n nibbles: the one-byte case is always < 0x80
__device__ __forceinline__ uint32_t hex_prefix_str_size(uint32_t n) { return n < 2 ? 1 : 2 + (n >> 1); }

// rlp_str(hex_prefix(key[start .. start+n), is_leaf)): at most 34 bytes
template <int B>
__device__ __forceinline__ void emit_hex_prefix_str(Stage<B>& s, const uint8_t* key, uint32_t start, uint32_t n, uint32_t is_leaf) {
  if (n >= 2) s.put_byte(0x80 + 1 + (n >> 1));
  uint32_t flag = (is_leaf ? 2u : 0u) + (n & 1);
  uint32_t j = start, end = start + n;
  if (n & 1) {
    s.put_byte((flag << 4) | key_nibble(key, j));
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

// up to 32 bytes at a 4-byte aligned global address
template <int B>
__device__ __forceinline__ void emit_chunk_aligned(Stage<B>& s, const uint8_t* p, uint32_t len) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
  uint32_t nw = len >> 2;
  for (uint32_t i = 0; i < nw; i++) s.put_word(__ldg(w + i));
  uint32_t rem = len & 3;
  if (rem) s.put_partial(__ldg(w + nw), rem);
}

// a child reference inside a parent: 0xa0 || hash, or the child's raw RLP when shorter than 32 bytes
template <int B>
__device__ __forceinline__ void emit_child_ref(Stage<B>& s, const uint8_t* ref, uint32_t len) {
  const uint4* q = reinterpret_cast<const uint4*>(ref);
  uint4 x = __ldcg(q), y = __ldcg(q + 1);
  if (len == 32) {
    s.put_byte(0xa0);
    s.put_word(x.x), s.put_word(x.y), s.put_word(x.z), s.put_word(x.w);
    s.put_word(y.x), s.put_word(y.y), s.put_word(y.z), s.put_word(y.w);
  } else {
    uint32_t w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t take = len > 4u * i ? min(4u, len - 4u * i) : 0u;
      s.put_partial(w[i], take);
    }
  }
}

__device__ __forceinline__ uint32_t u256_sig_bytes(const uint8_t* be32) {
  uint32_t i = 0;
  while (i < 32 && be32[i] == 0) i++;
  return 32 - i;
}
__device__ __forceinline__ uint32_t u256_str_size(const uint8_t* be32, uint32_t nbytes) {
  if (nbytes == 0) return 1;
  if (nbytes == 1 && be32[31] < 0x80) return 1;
  return 1 + nbytes;
}
// rlp of a U256 as a minimal big-endian string: at most 33 bytes
template <int B>
__device__ __forceinline__ void emit_u256_str(Stage<B>& s, const uint8_t* be32, uint32_t nbytes) {
  if (nbytes == 0) {
    s.put_byte(0x80);
    return;
  }
  if (!(nbytes == 1 && be32[31] < 0x80)) s.put_byte(0x80 + nbytes);
  for (uint32_t i = 32 - nbytes; i < 32; i++) s.put_byte(be32[i]);
}

__device__ __forceinline__ void store_digest(uint8_t* out32, const uint64_t (&a)[25]) {
  uint4* o = reinterpret_cast<uint4*>(out32);
  o[0] = make_uint4((uint32_t)a[0], (uint32_t)(a[0] >> 32), (uint32_t)a[1], (uint32_t)(a[1] >> 32));
  o[1] = make_uint4((uint32_t)a[2], (uint32_t)(a[2] >> 32), (uint32_t)a[3], (uint32_t)(a[3] >> 32));
}

// ------------------------------------------------------------------ keccak batch -------------
// Message i = data[offsets[i * stride] .. offsets[i * stride + 1]): stride 1 for a packed batch
// (n + 1 offsets), stride 2 for explicit (begin, end) pairs.  Segment = 32 bytes of the message.

template <int B>
__global__ void __launch_bounds__(B) keccak256_batch_kernel(const uint8_t* __restrict__ data, const uint64_t* __restrict__ offsets,
                                                            uint32_t stride, uint32_t n, uint8_t* __restrict__ out) {
  __shared__ uint32_t smem[PPD_STAGE_WORDS * B];
  uint32_t i = blockIdx.x * B + threadIdx.x;
  if (i >= n) return;
  Stage<B> s;
  s.init(smem);
  uint64_t a[25];
#pragma unroll
  for (int k = 0; k < 25; k++) a[k] = 0;
  const uint64_t m_begin = offsets[(uint64_t)i * stride], m_end = offsets[(uint64_t)i * stride + 1];
  const uint8_t* p = data + m_begin;
  uint64_t left = m_end - m_begin;
  // bring p to a 4-byte boundary
  while (left && (reinterpret_cast<uintptr_t>(p) & 3)) {
    s.put_byte(__ldg(p));
    p++;
    left--;
  }
  bool done = false;
  while (!done) {
    while (s.bytes() < 136 && left) {
      uint32_t take = left < 32 ? (uint32_t)left : 32u;
      emit_chunk_aligned(s, p, take);
      p += take;
      left -= take;
    }
    if (s.bytes() < 136) {
      s.pad();
      done = true;
    }
    absorb_stage<B>(a, s.w);
    keccak_f1600(a);
    if (!done) s.consume_block();
  }
  store_digest(out + 32ull * i, a);
}

// ------------------------------------------------------------------ trie level hashing -------

// order == nullptr: thread i handles node begin + i; else node order[begin + i]
template <int B>
__global__ void __launch_bounds__(B) hash_level_kernel(ArenaView A, const uint32_t* __restrict__ order, uint32_t begin, uint32_t end) {
  __shared__ uint32_t smem[PPD_STAGE_WORDS * B];
  uint32_t slot = begin + blockIdx.x * B + threadIdx.x;
  uint32_t hashed = 0, perms = 0, enc_bytes = 0;
  if (slot < end) {
    uint32_t id = order ? __ldg(order + slot) : slot;
    uint4 rec = __ldg(reinterpret_cast<const uint4*>(A.nodes) + id);
    const uint32_t kind = rec.x & 0xff, nib_start = (rec.x >> 8) & 0xff, nib_len = (rec.x >> 16) & 0xff;
    uint8_t* my_ref = A.ref + 32ull * id;

    // ---- per-kind set-up: payload length, segment count -------------------------------------
    uint32_t payload = 0, nseg = 0, total = 0;
    bool force_hash = false, copy_only = false;
    const uint8_t* copy_src = nullptr;
    const uint8_t* key = A.key_pool + rec.y;
    const uint8_t* val = nullptr;
    uint32_t vlen = 0, vhdr = 0;            // leaf value
    const AccountRec* acc = nullptr;        // account leaf
    uint32_t nn = 0, nb = 0, apayload = 0;  //   nonce / balance significant bytes, rlp(AccountRlp) payload
    const uint32_t* ch = nullptr;           // branch
    uint32_t mask = 0;
    switch (kind) {
      case NK_HASH:
        copy_only = true;
        copy_src = A.hash_pool + 32ull * rec.y;
        break;
      case NK_LEAF: {
        val = A.val_pool + rec.z;
        vlen = rec.w;
        vhdr = (vlen == 1 && __ldg(val) < 0x80) ? 0 : len_prefix_size(vlen);
        payload = hex_prefix_str_size(nib_len) + vhdr + vlen;
        nseg = 2 + ((vlen + 31) >> 5);
        break;
      }
      case NK_LEAF_ACCOUNT: {
        acc = A.accounts + rec.z;
        nn = u256_sig_bytes(acc->nonce);
        nb = u256_sig_bytes(acc->balance);
        apayload = u256_str_size(acc->nonce, nn) + u256_str_size(acc->balance, nb) + 66;  // >= 68
        vlen = len_prefix_size(apayload) + apayload;                                      // rlp(AccountRlp) >= 70
        payload = hex_prefix_str_size(nib_len) + len_prefix_size(vlen) + vlen;
        nseg = 5;
        break;
      }
      case NK_EXT: {
        uint32_t clen = A.ref_len[rec.z];
        payload = hex_prefix_str_size(nib_len) + (clen == 32 ? 33 : clen);
        nseg = 2;
        break;
      }
      case NK_BRANCH: {
        ch = A.child_pool + rec.y;
        mask = rec.z & 0xffff;
        payload = 17 - __popc(mask);  // empty children and the empty value: 0x80 each
        uint32_t k = 0;
        for (uint32_t m = mask; m; m &= m - 1, k++) {
          uint32_t clen = A.ref_len[__ldg(ch + k)];
          payload += clen == 32 ? 33 : clen;
        }
        nseg = 18;
        break;
      }
      case NK_ROOT: {
        force_hash = true;
        if (rec.z == NODE_EMPTY) {
          nseg = 1;
        } else if (A.ref_len[rec.z] == 32) {
          copy_only = true;
          copy_src = A.ref + 32ull * rec.z;
        } else {
          nseg = 1;
        }
        break;
      }
      default:
        copy_only = true;
        copy_src = my_ref;
        break;
    }
    if (copy_only) {
      const uint4* src = reinterpret_cast<const uint4*>(copy_src);
      uint4 x = __ldcg(src), y = __ldcg(src + 1);
      uint4* o = reinterpret_cast<uint4*>(my_ref);
      o[0] = x;
      o[1] = y;
      A.ref_len[id] = 32;
    } else {
      total = kind == NK_ROOT ? 0 : len_prefix_size(payload) + payload;
      Stage<B> s;
      s.init(smem);
      uint64_t a[25];
#pragma unroll
      for (int k = 0; k < 25; k++) a[k] = 0;
      uint32_t seg = 0;
      bool done = false;
      const bool inline_ref = !force_hash && total < 32;
      while (!done) {
        // ---- fill: append whole segments (each <= 48 bytes) until a rate block is complete ----
        while (s.bytes() < 136 && seg < nseg) {
          switch (kind) {
            case NK_LEAF:
              if (seg == 0) {
                emit_len_prefix(s, payload, 0xc0, 0xf7);
                emit_hex_prefix_str(s, key, nib_start, nib_len, 1);
              } else if (seg == 1) {
                if (vhdr) emit_len_prefix(s, vlen, 0x80, 0xb7);
              } else {
                uint32_t off = (seg - 2) << 5;
                emit_chunk_aligned(s, val + off, min(32u, vlen - off));
              }
              break;
            case NK_LEAF_ACCOUNT:
              if (seg == 0) {
                emit_len_prefix(s, payload, 0xc0, 0xf7);
                emit_hex_prefix_str(s, key, nib_start, nib_len, 1);
              } else if (seg == 1) {
                emit_len_prefix(s, vlen, 0x80, 0xb7);
                emit_len_prefix(s, apayload, 0xc0, 0xf7);
                emit_u256_str(s, acc->nonce, nn);
              } else if (seg == 2) {
                emit_u256_str(s, acc->balance, nb);
              } else if (seg == 3) {
                uint32_t src = acc->storage_src;
                emit_child_ref(s, src == NODE_EMPTY ? acc->storage_root : A.ref + 32ull * src, 32);
              } else {
                emit_child_ref(s, acc->code_hash, 32);
              }
              break;
            case NK_EXT:
              if (seg == 0) {
                emit_len_prefix(s, payload, 0xc0, 0xf7);
                emit_hex_prefix_str(s, key, nib_start, nib_len, 0);
              } else {
                emit_child_ref(s, A.ref + 32ull * rec.z, A.ref_len[rec.z]);
              }
              break;
            case NK_BRANCH:
              if (seg == 0) {
                emit_len_prefix(s, payload, 0xc0, 0xf7);
              } else if (seg == 17) {
                s.put_byte(0x80);
              } else {
                uint32_t bit = 1u << (seg - 1);
                if (mask & bit) {
                  uint32_t c = __ldg(ch + __popc(mask & (bit - 1)));
                  emit_child_ref(s, A.ref + 32ull * c, A.ref_len[c]);
                } else {
                  s.put_byte(0x80);
                }
              }
              break;
            default:  // NK_ROOT over Node::Empty or over a node whose encoding is < 32 bytes
              if (rec.z == NODE_EMPTY)
                s.put_byte(0x80);
              else
                emit_child_ref(s, A.ref + 32ull * rec.z, A.ref_len[rec.z]);
              break;
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
      if (inline_ref) {
        // the whole encoding (< 32 bytes) sits at the front of the stage: it is the ref
        s.flush_partial();
        uint32_t nw = (total + 3) >> 2;
        uint32_t h[8];
#pragma unroll
        for (int k = 0; k < 8; k++) h[k] = (uint32_t)k < nw ? s.w[k * B] : 0u;
        uint32_t tail = total & 3;
        if (tail) h[nw - 1] &= (1u << (8 * tail)) - 1;
        uint4* o = reinterpret_cast<uint4*>(my_ref);
        o[0] = make_uint4(h[0], h[1], h[2], h[3]);
        o[1] = make_uint4(h[4], h[5], h[6], h[7]);
        A.ref_len[id] = (uint8_t)total;
      } else {
        store_digest(my_ref, a);
        A.ref_len[id] = 32;
        hashed = 1;
        enc_bytes = kind == NK_ROOT ? (rec.z == NODE_EMPTY ? 1u : (uint32_t)A.ref_len[rec.z]) : total;
      }
    }
  }
  // work counters: one atomic pair per warp
  if (A.counters) {
    for (int off = 16; off > 0; off >>= 1) {
      hashed += __shfl_down_sync(0xffffffffu, hashed, off);
      perms += __shfl_down_sync(0xffffffffu, perms, off);
      enc_bytes += __shfl_down_sync(0xffffffffu, enc_bytes, off);
    }
    if ((threadIdx.x & 31) == 0 && hashed) {
      atomicAdd(A.counters + 0, (unsigned long long)hashed);
      atomicAdd(A.counters + 1, (unsigned long long)perms);
      atomicAdd(A.counters + 2, (unsigned long long)enc_bytes);
    }
  }
}

// ------------------------------------------------------------------ launchers ----------------

static constexpr int KB = 128;

void launch_keccak256_batch(const uint8_t* data, const uint64_t* offsets, uint32_t n, uint8_t* out, cudaStream_t st) {
  if (!n) return;
  keccak256_batch_kernel<KB><<<(n + KB - 1) / KB, KB, 0, st>>>(data, offsets, 1, n, out);
}
void launch_keccak256_ranges(const uint8_t* data, const uint64_t* begin_end, uint32_t n, uint8_t* out, cudaStream_t st) {
  if (!n) return;
  keccak256_batch_kernel<KB><<<(n + KB - 1) / KB, KB, 0, st>>>(data, begin_end, 2, n, out);
}
void launch_hash_level(const ArenaView& A, const uint32_t* order, uint32_t begin, uint32_t end, cudaStream_t st) {
  if (end <= begin) return;
  uint32_t n = end - begin;
  hash_level_kernel<KB><<<(n + KB - 1) / KB, KB, 0, st>>>(A, order, begin, end);
}

}  // namespace ppd
