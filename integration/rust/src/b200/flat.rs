// protocol_decoder/src/b200/flat.rs — BlockTrace -> FlatBlock and IrDump -> Vec<GenerationInputs> (include/ppd_flat.h).
// The Rust twin of proof_protocol_decoder_b200/flat.py (encode_flat_block, parse_ir_dump).  All integers little-endian,
// byte strings `u32 length || bytes`, U256 / H256 as 32 big-endian bytes.  Source only: this image has no rustc.
use std::collections::HashMap;
use eth_trie_utils::{nibbles::Nibbles, partial_trie::{HashedPartialTrie, Node, PartialTrie}};
use ethereum_types::{Address, H256, U256};
use plonky2_evm::generation::{GenerationInputs, TrieInputs};
use plonky2_evm::proof::TrieRoots;
use crate::trace_protocol::{BlockTrace, BlockTraceTriePreImages, CombinedPreImages, ContractCodeUsage,
                            SeparateStorageTriesPreImage, SeparateTriePreImage, SeparateTriePreImages, TrieCompact,
                            TrieDirect, TxnInfo};
use crate::types::OtherBlockData;

const FLAT_BLOCK_MAGIC: u32 = 0x4244_5050;
const IR_DUMP_MAGIC: u32 = 0x4944_5050;
const TR_BALANCE: u8 = 0x01;
const TR_NONCE: u8 = 0x02;
const TR_STORAGE_READ: u8 = 0x04;
const TR_STORAGE_WRITTEN: u8 = 0x08;
const TR_CODE_READ: u8 = 0x10;
const TR_CODE_WRITE: u8 = 0x20;
const TR_SELF_DESTRUCTED: u8 = 0x40;

fn put_u32(o: &mut Vec<u8>, v: u32) { o.extend_from_slice(&v.to_le_bytes()); }
fn put_bytes(o: &mut Vec<u8>, b: &[u8]) { put_u32(o, b.len() as u32); o.extend_from_slice(b); }
fn put_u256(o: &mut Vec<u8>, v: U256) { let mut b = [0u8; 32]; v.to_big_endian(&mut b); o.extend_from_slice(&b); }

/// Trie := Node in pre-order (include/ppd_flat.h); the same form `decode_ir_dump` reads back.
fn put_trie(o: &mut Vec<u8>, t: &HashedPartialTrie) {
    match &**t {   // PartialTrie: Deref<Target = Node<Self>>
        Node::Empty => o.push(0),
        Node::Hash(h) => { o.push(1); o.extend_from_slice(h.as_bytes()); }
        Node::Branch { children, value } => { o.push(2); for c in children { put_trie(o, c); } put_bytes(o, value); }
        Node::Extension { nibbles, child } => { o.push(3); put_nibbles(o, nibbles); put_trie(o, child); }
        Node::Leaf { nibbles, value } => { o.push(4); put_nibbles(o, nibbles); put_bytes(o, value); }
    }
}
fn put_nibbles(o: &mut Vec<u8>, n: &Nibbles) { o.push(n.count as u8); for i in 0..n.count { o.push(n.get_nibble(i)); } }
/// DirectPreImage: the state trie, then a trie per hashed address (any order; the library indexes them by address).
/// Tries not in `tries` stay hashed out: such an account keeps its storage root and has no storage trie.
fn put_direct_pre_image(state: &HashedPartialTrie, tries: &HashMap<H256, SeparateTriePreImage>) -> Vec<u8> {
    let mut o = Vec::new();
    put_trie(&mut o, state);
    put_u32(&mut o, tries.len() as u32);
    for (h_addr, t) in tries {
        let SeparateTriePreImage::Direct(TrieDirect(t)) = t else { unimplemented!("processed_block_trace.rs:144") };
        o.extend_from_slice(h_addr.as_bytes());
        put_trie(&mut o, t);
    }
    o
}

/// BlockTrace + the code every `ContractCodeUsage::Read` resolves to + OtherBlockData -> FlatBlock.
/// `o` may be a Vec or a slice over a `ppd_alloc_pinned` buffer (then the GPU's copy engine reads it in place).
pub fn encode_block(bt: &BlockTrace, resolved: &[(H256, Vec<u8>)], other: &OtherBlockData) -> Vec<u8> {
    // kind 0: the TrieCompact bytes.  kind 2: Separate{Direct state trie, MultipleTries of Direct tries} as a
    // DirectPreImage payload (put_direct_pre_image below); the other Separate forms are todo!() in the reference too.
    let (kind, pre_image): (u32, std::borrow::Cow<[u8]>) = match &bt.trie_pre_images {
        BlockTraceTriePreImages::Combined(CombinedPreImages { compact: TrieCompact(bytes) }) => (0, bytes.into()),
        BlockTraceTriePreImages::Separate(SeparateTriePreImages {
            state: SeparateTriePreImage::Direct(TrieDirect(state)),
            storage: SeparateStorageTriesPreImage::MultipleTries(tries),
        }) => (2, put_direct_pre_image(state, tries).into()),
        _ => unimplemented!("processed_block_trace.rs:144,161 are todo!() in the reference too"),
    };
    let mut o = Vec::with_capacity(pre_image.len() + (1 << 20));
    put_u32(&mut o, FLAT_BLOCK_MAGIC); put_u32(&mut o, 1); put_u32(&mut o, kind);
    put_bytes(&mut o, &pre_image);
    put_u32(&mut o, bt.txn_info.len() as u32);
    for TxnInfo { traces, meta } in &bt.txn_info {
        put_u32(&mut o, traces.len() as u32);
        for (addr, tr) in traces {                       // HashMap order: any order is legal (the reference iterates it too)
            o.extend_from_slice(addr.as_bytes());
            let at = o.len(); o.push(0);
            let mut flags = 0u8;
            if let Some(b) = tr.balance { flags |= TR_BALANCE; put_u256(&mut o, b); }
            if let Some(n) = tr.nonce { flags |= TR_NONCE; put_u256(&mut o, n); }
            if let Some(r) = &tr.storage_read {
                flags |= TR_STORAGE_READ; put_u32(&mut o, r.len() as u32);
                for k in r { o.extend_from_slice(k.as_bytes()); }
            }
            if let Some(w) = &tr.storage_written {
                flags |= TR_STORAGE_WRITTEN; put_u32(&mut o, w.len() as u32);
                for (k, v) in w { o.extend_from_slice(k.as_bytes()); put_u256(&mut o, *v); }
            }
            match &tr.code_usage {
                Some(ContractCodeUsage::Read(h)) => { flags |= TR_CODE_READ; o.extend_from_slice(h.as_bytes()); }
                Some(ContractCodeUsage::Write(c)) => { flags |= TR_CODE_WRITE; put_bytes(&mut o, &c.0); }
                None => {}
            }
            if tr.self_destructed.unwrap_or(false) { flags |= TR_SELF_DESTRUCTED; }
            o[at] = flags;
        }
        put_bytes(&mut o, &meta.byte_code);
        put_bytes(&mut o, &meta.new_txn_trie_node_byte);
        put_bytes(&mut o, &meta.new_receipt_trie_node_byte);
        o.extend_from_slice(&meta.gas_used.to_le_bytes());
    }
    put_u32(&mut o, resolved.len() as u32);
    for (h, code) in resolved { o.extend_from_slice(h.as_bytes()); put_bytes(&mut o, code); }
    let wd = &other.b_data.withdrawals;
    put_u32(&mut o, wd.len() as u32);
    for (a, amt) in wd { o.extend_from_slice(a.as_bytes()); put_u256(&mut o, *amt); }
    o.extend_from_slice(other.checkpoint_state_trie_root.as_bytes());
    // b_meta / b_hashes are opaque to the library, which copies them into every IR; decode_ir_dump takes both from
    // `other` instead, so the shim sends them empty (no serialiser needed, and 8 KB less per IR on the way back)
    put_bytes(&mut o, &[]);
    put_bytes(&mut o, &[]);
    o
}

struct R<'a> { b: &'a [u8], p: usize }
impl<'a> R<'a> {
    fn take(&mut self, n: usize) -> &'a [u8] { let v = &self.b[self.p..self.p + n]; self.p += n; v }
    fn u8(&mut self) -> u8 { self.take(1)[0] }
    fn u32(&mut self) -> u32 { u32::from_le_bytes(self.take(4).try_into().unwrap()) }
    fn bytes(&mut self) -> &'a [u8] { let n = self.u32() as usize; self.take(n) }
    fn h256(&mut self) -> H256 { H256::from_slice(self.take(32)) }
    fn u256(&mut self) -> U256 { U256::from_big_endian(self.take(32)) }
    fn nibbles(&mut self) -> Nibbles {
        let n = self.u8() as usize;
        let mut nib = Nibbles::default();
        for &x in self.take(n) { nib.push_nibble_back(x); }
        nib
    }
    /// a Trie blob: nodes in pre-order (0 empty, 1 hash, 2 branch, 3 extension, 4 leaf)
    fn node(&mut self) -> HashedPartialTrie {
        match self.u8() {
            0 => HashedPartialTrie::new(Node::Empty),   // the constructor decoding.rs:577 uses
            1 => HashedPartialTrie::new(Node::Hash(self.h256())),
            2 => {
                let children = std::array::from_fn(|_| std::sync::Arc::new(Box::new(self.node())));
                let value = self.bytes().to_vec();
                HashedPartialTrie::new(Node::Branch { children, value })
            }
            3 => { let nibbles = self.nibbles(); let child = std::sync::Arc::new(Box::new(self.node())); HashedPartialTrie::new(Node::Extension { nibbles, child }) }
            4 => { let nibbles = self.nibbles(); let value = self.bytes().to_vec(); HashedPartialTrie::new(Node::Leaf { nibbles, value }) }
            k => panic!("bad node kind {k} in IrDump"),
        }
    }
}

/// IrDump -> Vec<GenerationInputs> (decoding.rs:131-145: the fields in the order the reference fills them)
pub fn decode_ir_dump(dump: &[u8], other: &OtherBlockData) -> Vec<GenerationInputs> {
    let mut r = R { b: dump, p: 0 };
    assert_eq!(r.u32(), IR_DUMP_MAGIC);
    (0..r.u32()).map(|_| {
        let txn_number_before = r.u256();
        let gas_used_before = r.u256();
        let gas_used_after = r.u256();
        let has_txn = r.u8() != 0;
        let txn = r.bytes().to_vec();
        let withdrawals: Vec<(Address, U256)> = (0..r.u32()).map(|_| (Address::from_slice(r.take(20)), r.u256())).collect();
        let state_trie = r.node();
        let transactions_trie = r.node();
        let receipts_trie = r.node();
        let storage_tries = (0..r.u32()).map(|_| (r.h256(), r.node())).collect();
        let trie_roots_after = TrieRoots { state_root: r.h256(), transactions_root: r.h256(), receipts_root: r.h256() };
        let checkpoint_state_trie_root = r.h256();
        let contract_code: HashMap<H256, Vec<u8>> = (0..r.u32()).map(|_| (r.h256(), r.bytes().to_vec())).collect();
        let (_meta, _hashes) = (r.bytes(), r.bytes());   // the caller's own bytes: taken from `other` instead of re-decoded
        GenerationInputs {
            txn_number_before, gas_used_before, gas_used_after,
            signed_txn: has_txn.then_some(txn),
            withdrawals,
            tries: TrieInputs { state_trie, transactions_trie, receipts_trie, storage_tries },
            trie_roots_after, checkpoint_state_trie_root, contract_code,
            block_metadata: other.b_data.b_meta.clone(),
            block_hashes: other.b_data.b_hashes.clone(),
        }
    }).collect()
}
