"""Development helper (torchrun): every rank measures its PCIe link while all ranks copy at once, then alone in turn."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

def barrier():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

together = bench.measure_pcie(torch, barrier=barrier)
alone = None
for r in range(world):
    barrier()
    if r == rank:
        alone = bench.measure_pcie(torch)
    barrier()
t = torch.tensor([together, alone], dtype=torch.float64, device="cuda")
out = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(out, t)
if rank == 0:
    for r, o in enumerate(out):
        print(f"rank {r}: {o[0].item():6.1f} GB/s per direction with all ranks copying, {o[1].item():6.1f} alone")
    print(f"sum with all ranks copying: {sum(o[0].item() for o in out):.1f} GB/s per direction")
dist.destroy_process_group()
