#!/usr/bin/env python3
"""One config-5 style call (sorted leaves resident in HBM -> root) for profiling under ncu.
   python profiles/run_c5.py <n_leaves> [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from proof_protocol_decoder_b200.lib import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = Context(0)
g = torch.Generator(device="cuda")
g.manual_seed(5)
keys = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
hi = keys[:, :8].to(torch.int64)
k64 = torch.zeros(n, dtype=torch.int64, device="cuda")
for b in range(8):
    k64 = (k64 << 8) | hi[:, b]
keys = keys[torch.argsort((k64 >> 1) & 0x7FFFFFFFFFFFFFFF)].contiguous()
lens = torch.randint(70, 81, (n,), dtype=torch.int64, device="cuda", generator=g)
val_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
val_off[1:] = torch.cumsum(lens, 0)
vb = int(val_off[-1].item())
vals = torch.randint(0, 256, (vb,), dtype=torch.uint8, device="cuda", generator=g)
torch.cuda.synchronize()
for _ in range(reps):
    root = ctx.trie_root_sorted_leaves_dev(keys.data_ptr(), val_off.data_ptr(), vals.data_ptr(), n, vb)
    st = ctx.stats()
print(root.hex(), st)
