"""Multi-GPU sharding of the hot path: one process per GPU, no data-path collective.

The reference is single-threaded and has no notion of devices; what makes the path shardable is in
its data model:
  * blocks are independent — every BlockTrace carries its own pre-image
    (protocol_decoder/src/trace_protocol.rs:40-48), so block i goes to rank i mod G;
  * storage tries only meet at the account leaf's storage_root (decoding.rs:438-447,
    compact_prestate_processing.rs:617-621), so per-account tries are bin-packed over the ranks by
    size and only their 32-byte roots are exchanged.
The single exchange step is an all-gather of 32-byte roots (NCCL on the GPUs, gloo in the CPU tests).
"""
from typing import Callable, Dict, List, Sequence


def shard_blocks(n_blocks: int, rank: int, world: int) -> List[int]:
    """Indices of the blocks rank `rank` decodes: block i -> rank i mod world."""
    return list(range(rank, n_blocks, world))


def assign_tries(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Size-balanced assignment of independent tries to ranks (longest-processing-time first):
    result[r] = indices of the tries rank r hashes.  Deterministic (ties by index)."""
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i)):
        r = min(range(world), key=lambda r: (load[r], r))
        out[r].append(i)
        load[r] += int(sizes[i])
    for lst in out:
        lst.sort()
    return out


def all_gather_roots(local: Dict[int, bytes], n_total: int, dist=None, device="cpu") -> List[bytes]:
    """Every rank contributes {index: 32-byte root}; every rank gets the full list.  `dist` is
    torch.distributed (already initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert len(local) == n_total
        return [bytes(local[i]) for i in range(n_total)]
    import torch

    world = dist.get_world_size()
    buf = torch.zeros((n_total, 33), dtype=torch.uint8)
    for i, r in local.items():
        buf[i, 0] = 1
        buf[i, 1:] = torch.frombuffer(bytearray(r), dtype=torch.uint8)
    buf = buf.to(device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    parts = [p.cpu() for p in parts]
    out = []
    for i in range(n_total):
        owners = [p for p in parts if int(p[i, 0]) == 1]
        assert len(owners) == 1, f"root {i} has {len(owners)} owners"
        out.append(bytes(owners[0][i, 1:].numpy().tobytes()))
    return out


def sharded_trie_roots(root_fn: Callable[[int], bytes], sizes: Sequence[int], dist=None, device="cpu") -> List[bytes]:
    """Config 3 (storage-heavy block): per-account storage tries sharded over the ranks.
    root_fn(i) computes the root of trie i on this rank's GPU (Context.trie_root_sorted_leaves);
    returns the roots of all tries on every rank."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    mine = assign_tries(sizes, world)[rank]
    return all_gather_roots({i: root_fn(i) for i in mine}, len(sizes), dist, device)


def sharded_block_roots(decode_fn: Callable[[int], Sequence[bytes]], n_blocks: int, dist=None, device="cpu") -> List[List[bytes]]:
    """Config 4 (batch of blocks): block i is decoded on rank i mod world; decode_fn(i) returns the
    block's final (state, transactions, receipts) roots.  Every rank gets the roots of all blocks;
    the IR itself stays on the rank that produced it."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    local = {}
    for i in shard_blocks(n_blocks, rank, world):
        roots = decode_fn(i)
        assert len(roots) == 3
        for k in range(3):
            local[3 * i + k] = roots[k]
    flat = all_gather_roots(local, 3 * n_blocks, dist, device)
    return [flat[3 * i : 3 * i + 3] for i in range(n_blocks)]
