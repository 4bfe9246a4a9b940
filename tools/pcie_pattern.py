"""Development helper: what the PCIe link gives for the pipeline's copy pattern — many streams, each alternating a 37 MB
pinned upload and a 53 MB pinned download — against one big copy per direction."""
import sys, time
import torch

def run(n_streams, up_mb, down_mb, secs=1.0):
    ups = [torch.empty(up_mb << 20, dtype=torch.uint8).pin_memory() for _ in range(n_streams)] if up_mb else []
    downs = [torch.empty(down_mb << 20, dtype=torch.uint8).pin_memory() for _ in range(n_streams)] if down_mb else []
    d_up = [torch.empty(max(1, up_mb) << 20, dtype=torch.uint8, device="cuda") for _ in range(n_streams)]
    d_down = [torch.empty(max(1, down_mb) << 20, dtype=torch.uint8, device="cuda") for _ in range(n_streams)]
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rounds = 0
    while time.perf_counter() - t0 < secs:
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                if up_mb:
                    d_up[i].copy_(ups[i], non_blocking=True)
                if down_mb:
                    downs[i].copy_(d_down[i], non_blocking=True)
        rounds += 1
        if rounds % 4 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return rounds * n_streams * up_mb * 1.048576e-3 / dt, rounds * n_streams * down_mb * 1.048576e-3 / dt

if __name__ == "__main__":
    for n, u, d in [(1, 256, 0), (1, 0, 256), (1, 256, 256), (1, 37, 53), (4, 37, 53), (16, 37, 53), (16, 0, 53), (16, 37, 0), (16, 4, 53)]:
        up, down = run(n, u, d)
        print(f"streams {n:2d}  up {u:3d} MB  down {d:3d} MB per copy:  H2D {up:6.1f} GB/s   D2H {down:6.1f} GB/s   blocks/s at 37/53 MB: {min(up / 0.0389 if u else 9e9, down / 0.0560 if d else 9e9):7.0f}", flush=True)
