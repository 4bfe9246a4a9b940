"""Multi-GPU sharding of the hot path: one process per GPU, no data-path collective.

The reference is single-threaded and has no notion of devices; what makes the path shardable is in
its data model:
  * blocks are independent — every BlockTrace carries its own pre-image
    (protocol_decoder/src/trace_protocol.rs:40-48), so block i goes to rank i mod G;
  * storage tries only meet at the account leaf's storage_root (decoding.rs:438-447,
    compact_prestate_processing.rs:617-621), so per-account tries are bin-packed over the ranks by
    size and only their 32-byte roots are exchanged.
  * one huge trie (config 5, or a 1 M-slot storage trie of config 3) splits at its top nibble into 16 sub-tries
    whose keys share that nibble; each is hashed like a whole trie whose nodes sit one nibble below the root
    (ppd_trie_subroot_sorted_leaves_dev), 16 refs are gathered and the top branch is hashed last
    (ppd_trie_root_from_children).
The single exchange step is an all-gather of 32-byte roots / refs (NCCL on the GPUs, gloo in the CPU tests).
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


def top_nibble_ranges(torch, keys):
    """[lo, hi) of the leaves whose key starts with nibble i, for i in 0..15 (keys: sorted uint8 [n, 32] on any device)."""
    top = (keys[:, 0] >> 4).to(torch.int64)
    bounds = torch.searchsorted(top, torch.arange(17, dtype=torch.int64, device=keys.device)).tolist()
    return [(bounds[i], bounds[i + 1]) for i in range(16)]


def split_trie_refs(ctx, torch, keys, val_off, vals, nibbles=None):
    """The refs of the sub-tries below the top branch of the trie over the sorted leaves (device tensors), for the
    top nibbles in `nibbles` (default: all 16).  Returns (refs: 16 x 32 bytes, zeros where absent or not asked for;
    mask of the non-empty sub-tries among ALL 16; stats summed over the parts)."""
    ranges = top_nibble_ranges(torch, keys)
    refs = [bytes(32)] * 16
    mask = 0
    tot = {"nodes_hashed": 0, "node_permutations": 0, "node_bytes": 0, "gpu_ms": 0.0}
    for i, (lo, hi) in enumerate(ranges):
        if hi <= lo:
            continue
        mask |= 1 << i
        if nibbles is not None and i not in nibbles:
            continue
        v0 = int(val_off[lo].item())
        sub_off = (val_off[lo : hi + 1] - v0).contiguous()
        sub_keys = keys[lo:hi]
        refs[i] = ctx.trie_subroot_sorted_leaves_dev(sub_keys.data_ptr(), sub_off.data_ptr(), vals.data_ptr() + v0, hi - lo, int(sub_off[-1].item()), 1)
        st = ctx.stats()
        for k in tot:
            tot[k] += st[k]
    return refs, mask, tot


def all_gather_refs(torch, local_refs, owned, dist=None):
    """Every rank holds the refs of the top nibbles it owns; every rank gets all 16 (one NCCL all-reduce of 512 bytes:
    the entries a rank does not own are zero)."""
    buf = torch.zeros(512, dtype=torch.uint8, device="cuda" if torch.cuda.is_available() else "cpu")
    for i in owned:
        buf[32 * i : 32 * i + 32] = torch.frombuffer(bytearray(local_refs[i]), dtype=torch.uint8).to(buf.device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        wide = buf.to(torch.int32)
        dist.all_reduce(wide, op=dist.ReduceOp.SUM)
        buf = wide.to(torch.uint8)
    raw = bytes(buf.cpu().numpy().tobytes())
    return [raw[32 * i : 32 * i + 32] for i in range(16)]
