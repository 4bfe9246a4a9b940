// protocol_decoder/src/b200/mod.rs — the body `BlockTrace::into_txn_proof_gen_ir` (processed_block_trace.rs:38-50)
// gets when the crate is built with the B200 path.  Source only: this image has no rustc.
pub mod flat;
pub mod status;

use std::cell::RefCell;

use ethereum_types::H256;

use crate::decoding::TraceParsingResult;
use crate::gpu_ffi::GpuDecoder;
use crate::processed_block_trace::ProcessingMeta;
use crate::trace_protocol::{BlockTrace, ContractCodeUsage};
use crate::types::{CodeHashResolveFunc, OtherBlockData, TxnProofGenIR};

thread_local! {
    // one context per (thread, device); PPD_B200_DEVICE picks the device (default 0)
    static GPU: RefCell<GpuDecoder> = RefCell::new(
        GpuDecoder::new(std::env::var("PPD_B200_DEVICE").ok().and_then(|d| d.parse().ok()).unwrap_or(0))
            .unwrap_or_else(|rc| panic!("libppd_b200: no usable CUDA device (status {rc}); there is no CPU fallback")));
}

impl BlockTrace {
    pub fn into_txn_proof_gen_ir<F: CodeHashResolveFunc>(
        self, p_meta: &ProcessingMeta<F>, other_data: OtherBlockData,
    ) -> TraceParsingResult<Vec<TxnProofGenIR>> {
        // 1. Resolve every ContractCodeUsage::Read(hash) up front: all hashes are visible in the input
        //    (trace_protocol.rs:189-196), so the callback never has to cross the C boundary.
        let resolved: Vec<(H256, Vec<u8>)> = self.txn_info.iter()
            .flat_map(|t| t.traces.values())
            .filter_map(|tr| match &tr.code_usage { Some(ContractCodeUsage::Read(h)) => Some(*h), _ => None })
            .collect::<std::collections::BTreeSet<_>>().into_iter()
            .map(|h| (h, (p_meta.resolve_code_hash_fn)(&h))).collect();   // (the field becomes pub(crate): processed_block_trace.rs:188)
        // 2. BlockTrace + resolved code + OtherBlockData -> FlatBlock (include/ppd_flat.h, "input").
        let flat = flat::encode_block(&self, &resolved, &other_data);
        // 3. One call; kernels, copies and host threads are the library's business.
        let dump = GPU.with(|g| g.borrow_mut().block_decode(&flat)).map_err(status::status_to_error)?;
        // 4. IrDump -> Vec<GenerationInputs> (include/ppd_flat.h, "output"): every Trie is rebuilt as a
        //    HashedPartialTrie from its pre-order Node list; storage_tries / contract_code arrive sorted.
        Ok(flat::decode_ir_dump(&dump, &other_data))
    }
}
