/* Status codes shared by the CUDA library (libppd_b200.so) and the CPU oracle.
 *
 * One code per variant of the reference's two error enums, plus one code per
 * place where the reference panics instead of returning an error.  The Rust
 * shim maps them back to `CompactParsingError` /
 * `TraceParsingError` values or re-raises the panic.
 *
 *   CompactParsingError  protocol_decoder/src/compact/compact_prestate_processing.rs:49-93
 *   TraceParsingError    protocol_decoder/src/decoding.rs:31-49
 *   panic sites          protocol_decoder/src/processed_block_trace.rs:91,144,161,167,172,175,340
 *                        protocol_decoder/src/decoding.rs:202,228-230
 */
#ifndef PPD_STATUS_H
#define PPD_STATUS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ppd_status {
  PPD_OK = 0,

  /* CompactParsingError, in declaration order */
  PPD_ERR_MISSING_HEADER = 1,
  PPD_ERR_INVALID_OPERATOR = 2,
  PPD_ERR_UNEXPECTED_END_OF_STREAM = 3,
  PPD_ERR_INVALID_BYTE_VECTOR = 4,
  PPD_ERR_INVALID_BYTES_FOR_TYPE = 5,
  PPD_ERR_INVALID_WITNESS_FORMAT = 6,
  PPD_ERR_NON_SINGLE_ENTRY_AFTER_PROCESSING = 7,
  PPD_ERR_INCORRECT_NUMBER_OF_NODES_PRECEDING_BRANCH = 8,
  PPD_ERR_MISSING_EXPECTED_NODES_PRECEDING_BRANCH = 9,
  PPD_ERR_PRECEDING_NON_NODE_ENTRY = 10,
  PPD_ERR_KEY_ERROR = 11,

  /* TraceParsingError, in declaration order */
  PPD_ERR_ACCOUNT_DECODE = 21,
  PPD_ERR_MISSING_ACCOUNT_STORAGE_TRIE = 22,
  PPD_ERR_NON_EXISTENT_TRIE_ENTRY = 23,
  PPD_ERR_MISSING_KEYS_CREATING_SUB_PARTIAL_TRIE = 24,
  PPD_ERR_MISSING_WITHDRAWAL_ACCOUNT = 25,

  /* places where the reference panics */
  PPD_PANIC_INCOMPATIBLE_HEADER_VERSION = 40, /* processed_block_trace.rs:175 */
  PPD_PANIC_INSERT_INTO_HASH_NODE = 41,       /* eth_trie_utils insert through Node::Hash */
  PPD_PANIC_H256_FROM_SLICE = 42,             /* decoding.rs:202,228-230 with a short bytes_be() */
  PPD_PANIC_RECEIPT_DECODE = 43,              /* processed_block_trace.rs:340 */
  PPD_PANIC_PRE_IMAGE_ACCOUNT_DECODE = 44,    /* processed_block_trace.rs:91, compact_to_partial_trie.rs:176 */
  PPD_PANIC_UNIMPLEMENTED_PRE_IMAGE = 45,     /* todo!() at processed_block_trace.rs:144,161,167 */
  PPD_PANIC_KEY_IS_PREFIX_OF_KEY = 46,        /* Nibbles::get_nibble(0) on an empty postfix */
  PPD_PANIC_U256_FROM_BIG_ENDIAN = 47,        /* compact_prestate_processing.rs read_cbor_u256: U256::from_big_endian of more than 32 bytes (an account leaf's balance) */

  /* this library's own failures */
  PPD_ERR_BAD_FLAT_INPUT = 60,
  PPD_ERR_UNRESOLVED_CODE_HASH = 61,
  PPD_ERR_BAD_ARGUMENT = 62,
  PPD_ERR_UNSORTED_KEYS = 63,
  PPD_ERR_CUDA = 100
} ppd_status;

#ifdef __cplusplus
}
#endif
#endif
