// protocol_decoder/src/b200/status.rs — include/ppd_status.h -> what the reference does at the same place.
// Source only: this image has no rustc.
//
//  1-11   CompactParsingError variants (compact_prestate_processing.rs:49-93).  The reference unwrap()s them at
//         processed_block_trace.rs:172, i.e. it panics; so does the shim, naming the variant.  Their payloads
//         (cursor positions, witness entries) are Debug output of parser-internal types and are not rebuilt:
//         the panic text carries the library's own description instead.
//  21-25  TraceParsingError variants (decoding.rs:31-49): RETURNED, with the reference's payload.  ppd_last_error
//         spells the payload after "; " as key=value words (csrc/err_detail.h).
//  40-47  the reference's own panic sites: the shim panics too.
//  60-100 failures with no counterpart in the reference (bad flat input, CUDA).
use std::collections::HashMap;
use std::str::FromStr;

use ethereum_types::{Address, H256, U256};
use plonky2_evm::generation::mpt::AccountRlp;

use crate::decoding::{TraceParsingError, TrieType};

const COMPACT_VARIANTS: [&str; 11] = [
    "MissingHeader", "InvalidOperator", "UnexpectedEndOfStream", "InvalidByteVector", "InvalidBytesForType",
    "InvalidWitnessFormat", "NonSingleEntryAfterProcessing", "IncorrectNumberOfNodesPrecedingBranch",
    "MissingExpectedNodesPrecedingBranch", "PrecedingNonNodeEntryFoundWhenProcessingRule", "KeyError",
];

/// "sentence; k1=v1 k2=v2" -> {k1: v1, k2: v2}
fn payload(detail: &str) -> HashMap<&str, &str> {
    detail.split_once("; ").map(|(_, words)| words.split(' ').filter_map(|w| w.split_once('=')).collect()).unwrap_or_default()
}
fn h256(p: &HashMap<&str, &str>, k: &str) -> H256 { H256::from_str(p[k]).expect("64 hex digits from the library") }

pub fn status_to_error((code, detail): (i32, String)) -> TraceParsingError {
    let p = payload(&detail);
    match code {
        1..=11 => panic!("called `Result::unwrap()` on an `Err` value: {}: {detail}", COMPACT_VARIANTS[code as usize - 1]),
        21 => {
            // AccountDecode(hex of the bytes, the rlp crate's error text): the text is re-derived by decoding the same
            // bytes with the same crate (decoding.rs:604-607).  A value that was not resident on the host has no bytes.
            let hex_bytes = p.get("bytes").copied().unwrap_or("");
            let rlp_err = hex::decode(hex_bytes).ok()
                .and_then(|b| rlp::decode::<AccountRlp>(&b).err())
                .map(|e| e.to_string()).unwrap_or_else(|| detail.clone());
            TraceParsingError::AccountDecode(hex_bytes.to_string(), rlp_err)
        }
        22 => TraceParsingError::MissingAccountStorageTrie(h256(&p, "hashed_addr")),
        // 23 NonExistentTrieEntry is declared by the reference but constructed nowhere (decoding.rs:40); the library never returns it
        24 => TraceParsingError::MissingKeysCreatingSubPartialTrie(match p["trie_type"] {
            "State" => TrieType::State,
            "Storage" => TrieType::Storage,
            "Receipt" => TrieType::Receipt,
            _ => TrieType::Txn,
        }),
        25 => TraceParsingError::MissingWithdrawalAccount(
            Address::from_str(p["addr"]).expect("40 hex digits from the library"),
            h256(&p, "hashed_addr"),
            U256::from_str_radix(p["amount"], 16).expect("64 hex digits from the library"),
        ),
        40 => panic!("TODO: Make this into a result... (compact header version)"),            // processed_block_trace.rs:175
        41 => panic!("Found a `Hash` node during an insert in a `PartialTrie`"),                // eth_trie_utils insert
        42 => panic!("H256::from_slice: bytes_be() shorter than 32 bytes"),                    // decoding.rs:202,228-230
        43 => panic!("receipt node bytes do not decode"),                                      // processed_block_trace.rs:340
        44 => panic!("pre-image account bytes do not decode"),                                 // processed_block_trace.rs:91
        45 => unimplemented!(),                                                                // processed_block_trace.rs:144,161,167
        46 => panic!("Nibbles::get_nibble(0) on an empty postfix"),
        47 => panic!("U256::from_big_endian: more than 32 bytes"),                             // read_cbor_u256
        60 | 62 | 63 => panic!("libppd_b200: the shim handed the library malformed input (status {code}): {detail}"),
        61 => panic!("libppd_b200: a ContractCodeUsage::Read hash was not resolved before the call: {detail}"),
        100 => panic!("libppd_b200: CUDA failure: {detail}"),
        _ => panic!("libppd_b200: status {code}: {detail}"),
    }
}
