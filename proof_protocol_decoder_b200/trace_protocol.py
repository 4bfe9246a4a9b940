"""Host-side mirror of the reference's public interface for the path.

Same names, argument meaning and error behaviour as
  protocol_decoder/src/trace_protocol.rs:40-205        BlockTrace, TxnInfo, TxnMeta, TxnTrace, ContractCodeUsage
  protocol_decoder/src/types.rs:50-64                  OtherBlockData, BlockLevelData
  protocol_decoder/src/processed_block_trace.rs:37-50, 183-200   BlockTrace::into_txn_proof_gen_ir, ProcessingMeta
The reference's toolchain (Rust) is absent from this image, so the host side above the C ABI
is Python here; the Rust shim a maintainer would add is shown in INTEGRATION.md.
"""
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from . import flat
from .lib import Context, PpdError


@dataclass
class ContractCodeUsage:
    """trace_protocol.rs:189-196: exactly one of `read` (code hash) / `write` (new code bytes)."""

    read: Optional[bytes] = None
    write: Optional[bytes] = None


@dataclass
class TxnTrace:
    """trace_protocol.rs:152-183"""

    balance: Optional[int] = None
    nonce: Optional[int] = None
    storage_read: Optional[List[bytes]] = None
    storage_written: Optional[Dict[bytes, int]] = None
    code_usage: Optional[ContractCodeUsage] = None
    self_destructed: Optional[bool] = None


@dataclass
class TxnMeta:
    """trace_protocol.rs:126-145"""

    byte_code: bytes = b""
    new_txn_trie_node_byte: bytes = b""
    new_receipt_trie_node_byte: bytes = b""
    gas_used: int = 0


@dataclass
class TxnInfo:
    """trace_protocol.rs:112-122"""

    traces: Dict[bytes, TxnTrace] = field(default_factory=dict)
    meta: TxnMeta = field(default_factory=TxnMeta)


@dataclass
class BlockLevelData:
    """types.rs:58-64; b_meta / b_hashes are opaque to this path and copied into every IR."""

    b_meta: bytes = b""
    b_hashes: bytes = b""
    withdrawals: List[Tuple[bytes, int]] = field(default_factory=list)


@dataclass
class OtherBlockData:
    """types.rs:50-55"""

    b_data: BlockLevelData = field(default_factory=BlockLevelData)
    checkpoint_state_trie_root: bytes = bytes(32)


class ProcessingMeta:
    """processed_block_trace.rs:183-200: carries the CodeHashResolveFunc callback."""

    def __init__(self, resolve_code_hash_fn: Callable[[bytes], bytes]):
        self.resolve_code_hash_fn = resolve_code_hash_fn


@dataclass
class BlockTrace:
    """trace_protocol.rs:40-48.  `trie_pre_images` is {"combined": {"compact": bytes}} -- the only variant the
    reference implements end to end (processed_block_trace.rs:117-181) -- or
    {"separate": {"state": {"direct": trie}, "storage": {"multiple_tries": {hashed address: {"direct": trie}}}}}
    (trace_protocol.rs:58-108) with tries as node tuples (flat.encode_node) or encoded bytes: FlatBlock kind 2, this
    repo's completion of the reference's todo!() for a Direct trie per account (csrc/host_direct.cu).  Every other
    `separate` form is reported as unimplemented, as the reference's todo!() would."""

    trie_pre_images: dict
    txn_info: List[TxnInfo] = field(default_factory=list)

    @classmethod
    def from_json(cls, src) -> "BlockTrace":
        """The reference's serde JSON form (trace_protocol.rs:40-205, deserializers.rs:8-79): text or a parsed value."""
        from . import wire

        return wire.block_trace_from_json(src)

    def to_json(self) -> str:
        from . import wire

        return wire.block_trace_dumps(self)

    def to_flat(self, p_meta: ProcessingMeta, other_data: OtherBlockData) -> bytes:
        pre = self.trie_pre_images
        kind = flat.PRE_IMAGE_COMBINED
        if "combined" in pre:
            compact = pre["combined"]["compact"]
        else:
            sep = pre.get("separate") or {}
            state, storage = sep.get("state") or {}, sep.get("storage") or {}
            tries = storage.get("multiple_tries") if isinstance(storage, dict) else None
            if not (isinstance(state, dict) and "direct" in state and isinstance(tries, dict) and all(isinstance(t, dict) and "direct" in t for t in tries.values())):
                raise PpdError(45, "pre-image variant the reference leaves as todo!() (processed_block_trace.rs:144,161,167)")
            kind = flat.PRE_IMAGE_DIRECT
            compact = flat.encode_direct_pre_image(state["direct"], {h: t["direct"] for h, t in tries.items()})
        txns, wanted = [], []
        for ti in self.txn_info:
            traces = []
            for addr, tr in ti.traces.items():
                d = {"balance": tr.balance, "nonce": tr.nonce, "storage_read": tr.storage_read, "self_destructed": bool(tr.self_destructed)}
                if tr.storage_written is not None:
                    d["storage_written"] = list(tr.storage_written.items())
                if tr.code_usage is not None:
                    if tr.code_usage.read is not None:
                        d["code_read"] = tr.code_usage.read
                        wanted.append(tr.code_usage.read)
                    else:
                        d["code_write"] = tr.code_usage.write
                traces.append((addr, d))
            txns.append(
                {
                    "traces": traces,
                    "byte_code": ti.meta.byte_code,
                    "new_txn_trie_node_byte": ti.meta.new_txn_trie_node_byte,
                    "new_receipt_trie_node_byte": ti.meta.new_receipt_trie_node_byte,
                    "gas_used": ti.meta.gas_used,
                }
            )
        # The callback does not cross the ABI: every Read(code_hash) is visible in the input, so it is
        # resolved here, up front (the library consults code carried by the witness first, as the
        # reference does at processed_block_trace.rs:70-81).
        resolved, seen = [], set()
        for h in wanted:
            if h not in seen:
                seen.add(h)
                code = p_meta.resolve_code_hash_fn(h)
                if code is not None:  # None: the caller knows the witness carries this code
                    resolved.append((h, code))
        return flat.encode_flat_block(
            compact, txns, resolved, other_data.b_data.withdrawals, other_data.checkpoint_state_trie_root, other_data.b_data.b_meta, other_data.b_data.b_hashes,
            pre_image_kind=kind,
        )

    def into_txn_proof_gen_ir(self, p_meta: ProcessingMeta, other_data: OtherBlockData, ctx: Optional[Context] = None):
        """processed_block_trace.rs:38-50 -> Vec<TxnProofGenIR>; raises PpdError for every
        TraceParsingError variant and for the reference's panics."""
        ctx = ctx or default_context()
        return flat.parse_ir_dump(ctx.block_decode(self.to_flat(p_meta, other_data)))


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        import os

        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
