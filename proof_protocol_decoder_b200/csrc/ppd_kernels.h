// ppd_kernels.h — launchers of the sm_100a kernels (ppd_kernels.cu, ppd_build.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "arena.h"

namespace ppd {

void launch_keccak256_batch(const uint8_t* data, const uint64_t* offsets, uint32_t n, uint8_t* out, cudaStream_t st);
void launch_keccak256_ranges(const uint8_t* data, const uint64_t* begin_end, uint32_t n, uint8_t* out, cudaStream_t st);
void launch_hash_level(const ArenaView& A, const uint32_t* order, uint32_t begin, uint32_t end, cudaStream_t st);


// ---- ppd_build.cu: trie construction from sorted leaves ----
struct Pyramid {
  const int8_t *L, *m1, *m2, *m3;
};

struct BuildView {
  const uint8_t* keys;       // [N][32]
  const uint64_t* val_off;   // [N+1]
  const uint8_t* vals;
  uint32_t n;
  Pyramid P;
  const uint32_t* leader;    // [N+1]
  const uint32_t* bidx;      // [N+1] exclusive scan of leader flags
  // branches
  uint8_t* depth;            // [B]
  uint8_t* ext_start;        // [B]  first nibble of the extension above the branch (== depth when none)
  uint32_t* rep;             // [B]  an item below the branch
  uint32_t* child;           // [B][16]
  uint32_t* root_id;         // id of the root node (leaf or branch)
  // results
  uint8_t* ref;              // [N + B][32]
  uint8_t* ref_len;          // [N + B]
  uint8_t* root_out;         // [32]
  unsigned long long* counters;
};

size_t scan_tmp_words(size_t n);
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t st);
void launch_lcp(const uint8_t* keys, uint32_t n, int8_t* L, uint32_t* flags, cudaStream_t st);
void launch_min64(const int8_t* in, uint32_t n_in, int8_t* out, uint32_t n_out, cudaStream_t st);
void launch_leaders(const int8_t* L, const int8_t* m1, const int8_t* m2, const int8_t* m3, uint32_t n, uint32_t* link_a, uint32_t* link_b,
                    uint32_t* flag, cudaStream_t st);
void launch_branch_info(const BuildView& V, cudaStream_t st);
void launch_child_count(const uint32_t* leader, const uint32_t* bidx, uint32_t n, uint32_t* nchild, cudaStream_t st);
void launch_depth_hist(const uint8_t* depth, const uint32_t* nchild, uint32_t nb, uint32_t* hist, cudaStream_t st);
void launch_branch_scatter(const uint8_t* depth, const uint32_t* nchild, uint32_t nb, uint32_t* cursor, uint32_t* order, cudaStream_t st);
void launch_hash_sorted_leaves(const BuildView& V, cudaStream_t st);
void launch_hash_branch_level(const BuildView& V, const uint32_t* order, uint32_t begin, uint32_t end, cudaStream_t st);

// ---- ppd_dump.cu: the per-txn sub-tries serialised on the GPU ----
static const uint32_t IR_SEG_LITERAL = 0xfffffffeu;  // seg_b of a literal segment (seg_a = its length); else seg_b = root of a trie
static const uint32_t IR_SEG_REF = 0xfffffffdu;      // seg_b of a "32 bytes of ref[seg_a]" segment (a trie root after the txn)
struct IrDumpPlanView {
  // inputs (per block): touched node ids of every IR, and every IR's segments in output order
  const uint32_t* touched;        // concatenated
  const uint32_t* touched_begin;  // [n_ir + 1]
  const uint32_t* seg_a;
  const uint32_t* seg_b;
  const uint32_t* seg_begin;      // [n_ir + 1]
  const uint64_t* ir_base;        // [n_ir] byte offset of the IR in the output (emit only)
  // outputs of ir_size_kernel
  uint32_t* seg_off;              // offset of every segment inside its IR
  uint32_t* ir_size;
  uint32_t* ir_flag;              // 1: the host must serialise this IR
  uint32_t* ir_nuniq;
  uint32_t* u_node;               // [touched_begin[ir] + k]: k-th unique touched node of the IR, its size, its offset
  uint32_t* u_size;
  uint32_t* u_off;
};
void launch_ir_size(const ArenaView& A, const IrDumpPlanView& P, uint32_t n_ir, cudaStream_t st);
void launch_ir_emit(const ArenaView& A, const IrDumpPlanView& P, uint32_t n_ir, uint8_t* out, cudaStream_t st);

// ---- ppd_microbench.cu ----
bool launch_microbench(int variant, uint32_t* out, uint32_t blocks_per_sm, uint32_t iters, uint32_t* block_threads, double* units_per_thread_iter,
                       cudaStream_t st);

}  // namespace ppd
