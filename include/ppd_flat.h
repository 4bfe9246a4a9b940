/* Flat wire formats crossing the C ABI.  All integers little-endian, all
 * U256 / H256 / Address values big-endian byte strings (as the reference's
 * ethereum-types serialise them).  These layouts are defined by this repo; the
 * reference has no serialised form of its own for this path.
 *
 * ---------------------------------------------------------------- input ---
 * FlatBlock  ==  BlockTrace + resolved code + OtherBlockData
 *                (protocol_decoder/src/trace_protocol.rs:40-205, types.rs:50-64)
 *
 *   u32 magic = PPD_FLAT_BLOCK_MAGIC, u32 version = 1
 *   u32 pre_image_kind      0 = Combined{compact} (the only variant the reference implements end to end)
 *                           2 = Separate{state: Direct, storage: MultipleTries{hashed address -> Direct}}
 *                               (trace_protocol.rs:58-108).  The reference takes the Direct state trie as it is
 *                               (processed_block_trace.rs:143-148) and leaves the storage half as todo!() (:164-168);
 *                               kind 2 completes it: every entry is the trie it holds, no code mappings (:139).
 *                           any other value: PPD_PANIC_UNIMPLEMENTED_PRE_IMAGE (the reference's todo!() variants)
 *   u32 pre_image_len, u8 pre_image[pre_image_len]
 *                           kind 0: TrieCompact bytes
 *                           kind 2: DirectPreImage := Trie state_trie
 *                                                     u32 n_storage; n x { u8 hashed_addr[32]; Trie }
 *                                   (Trie as in the IrDump below; state leaves hold rlp(AccountRlp), storage leaves
 *                                   rlp(value), exactly the bytes the tries hash)
 *   u32 n_txns
 *   n_txns x TxnInfo:
 *     u32 n_traces
 *     n_traces x { u8 addr[20]; u8 flags;                     TxnTrace, trace_protocol.rs:152-183
 *         if flags&PPD_TR_BALANCE      u8 balance[32]
 *         if flags&PPD_TR_NONCE        u8 nonce[32]
 *         if flags&PPD_TR_STORAGE_READ    u32 n; n x u8 key[32]
 *         if flags&PPD_TR_STORAGE_WRITTEN u32 n; n x { u8 key[32]; u8 value[32] }
 *         if flags&PPD_TR_CODE_READ    u8 code_hash[32]
 *         if flags&PPD_TR_CODE_WRITE   u32 len; u8 code[len]
 *       }                                                      PPD_TR_SELF_DESTRUCTED = Some(true)
 *     u32 byte_code_len, bytes                                 TxnMeta.byte_code
 *     u32 new_txn_trie_node_len, bytes                         TxnMeta.new_txn_trie_node_byte (never read by the reference)
 *     u32 new_receipt_trie_node_len, bytes                     TxnMeta.new_receipt_trie_node_byte
 *     u64 gas_used
 *   u32 n_code;  n_code x { u8 hash[32]; u32 len; bytes }      results of the CodeHashResolveFunc callback, resolved by the shim up front
 *   u32 n_withdrawals; n x { u8 addr[20]; u8 amount[32] }      BlockLevelData.withdrawals
 *   u8  checkpoint_state_trie_root[32]
 *   u32 b_meta_len, bytes                                      BlockMetadata, opaque to this path (copied into every IR)
 *   u32 b_hashes_len, bytes                                    BlockHashes, opaque to this path
 *
 * --------------------------------------------------------------- output ---
 * IrDump  ==  Vec<GenerationInputs> in a normalised, byte-comparable form
 *             (decoding.rs:131-145, 507-519).  Hash-map ordered content is
 *             sorted (SURVEY.md 8c hazard 1): storage_tries by hashed address,
 *             contract_code by code hash.
 *
 *   u32 magic = PPD_IR_DUMP_MAGIC, u32 n_ir
 *   n_ir x {
 *     u8 txn_number_before[32], gas_used_before[32], gas_used_after[32]
 *     u8 has_signed_txn; u32 len; bytes
 *     u32 n_withdrawals; n x { u8 addr[20]; u8 amount[32] }
 *     Trie state_trie, transactions_trie, receipts_trie
 *     u32 n_storage; n x { u8 hashed_addr[32]; Trie }
 *     u8 state_root_after[32], transactions_root_after[32], receipts_root_after[32]
 *     u8 checkpoint_state_trie_root[32]
 *     u32 n_code; n x { u8 hash[32]; u32 len; bytes }
 *     u32 b_meta_len, bytes;  u32 b_hashes_len, bytes
 *   }
 *   Trie := Node (pre-order)
 *   Node := u8 kind
 *           PPD_NODE_EMPTY
 *         | PPD_NODE_HASH      u8 hash[32]
 *         | PPD_NODE_BRANCH    16 x Node; u32 value_len; bytes
 *         | PPD_NODE_EXTENSION u8 n_nibbles; u8 nibble[n]; Node
 *         | PPD_NODE_LEAF      u8 n_nibbles; u8 nibble[n]; u32 value_len; bytes
 *
 * ------------------------------------------------------ pre-image dump ---
 * PreImageDump == ProcessedCompactOutput (compact_prestate_processing.rs:1250-1253)
 *   u32 magic = PPD_PRE_IMAGE_MAGIC, u8 header_version
 *   u8 state_root[32]
 *   u32 n_storage; n x { u8 hashed_addr[32]; u8 storage_root[32] }   sorted by hashed_addr
 *   u32 n_code;    n x { u8 code_hash[32]; u32 len }                  sorted by code hash
 *   u64 nodes_hashed, u64 keccak_permutations                        work counters of the root computation
 */
#ifndef PPD_FLAT_H
#define PPD_FLAT_H

#define PPD_FLAT_BLOCK_MAGIC 0x42445050u /* "PPDB" */
#define PPD_IR_DUMP_MAGIC 0x49445050u    /* "PPDI" */
#define PPD_PRE_IMAGE_MAGIC 0x50445050u  /* "PPDP" */

#define PPD_TR_BALANCE 0x01
#define PPD_TR_NONCE 0x02
#define PPD_TR_STORAGE_READ 0x04
#define PPD_TR_STORAGE_WRITTEN 0x08
#define PPD_TR_CODE_READ 0x10
#define PPD_TR_CODE_WRITE 0x20
#define PPD_TR_SELF_DESTRUCTED 0x40

#define PPD_NODE_EMPTY 0
#define PPD_NODE_HASH 1
#define PPD_NODE_BRANCH 2
#define PPD_NODE_EXTENSION 3
#define PPD_NODE_LEAF 4

/* compact witness opcodes, compact_prestate_processing.rs:129-138 */
#define PPD_OP_LEAF 0x00
#define PPD_OP_EXTENSION 0x01
#define PPD_OP_BRANCH 0x02
#define PPD_OP_HASH 0x03
#define PPD_OP_CODE 0x04
#define PPD_OP_ACCOUNT_LEAF 0x05
#define PPD_OP_EMPTY_ROOT 0x06

#endif
