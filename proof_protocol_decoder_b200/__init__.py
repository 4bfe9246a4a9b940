"""B200-native (sm_100a) decoding + root hashing for proof-protocol-decoder's hot path.

The product is the C-ABI shared library ``libppd_b200.so`` (include/ppd_b200.h).  This package
is the host-side mirror of the reference's Rust interface for the path
(`BlockTrace::into_txn_proof_gen_ir`, protocol_decoder/src/processed_block_trace.rs:38) on top
of that ABI.  There is no CPU fallback: without the CUDA extension every call raises.
"""
from .lib import PpdError, PpdLibrary, build_extension, load_library  # noqa: F401
from .trace_protocol import (  # noqa: F401
    BlockLevelData,
    BlockTrace,
    ContractCodeUsage,
    OtherBlockData,
    ProcessingMeta,
    TxnInfo,
    TxnMeta,
    TxnTrace,
)
