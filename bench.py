#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path (BASELINE.json metric: MPT nodes keccak-hashed/sec
and blocks decoded/sec).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path, one rank per GPU
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU algorithm (oracle port)
                                                            on the host cores, same metric and config

Workload (config.workload): C2, the mainnet-shaped block of BASELINE.json configs[1] — ~20k touched
accounts embedded in a virtual 16^7-account state, 200 txns — one block per GPU per step (weak
scaling: rank r decodes the block of seed 2 + r).  A step = every Keccak of the block: all
addresses / slots / code plus every node of every version of the state, storage, txn and receipt
tries, hashed level-synchronously.
  value  nodes/s with the block's arena and key messages already resident in HBM (device time,
         CUDA events on the library's stream, L2 flushed between steps)
  e2e    the same metric through the C ABI (ppd_block_decode) with HOST buffers: FlatBlock in,
         IrDump out, every host<->device copy and all host work inside the timed region
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALU_OPS_PER_PERM = 4354  # 24 rounds x ~180 LOP3/SHF + 34 to absorb a rate block
ALU_LANES_PER_SM_CLK = 64
N_SM = 148


def load_traffic():
    """Mean DRAM bytes per hash_level_kernel launch from the latest committed ncu launch list (profiles/rNN_traffic.json)."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d.get("dram_bytes_per_launch_mean"), os.path.relpath(files[-1], ROOT)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", self.gpu],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


C2_PARAMS = dict(contract_frac=0.15, slots_lo=1, slots_hi=256, virtual_depth=7, virtual_accounts_log16=7, slot_reads=(0, 3), slot_writes=(0, 3),
                 allow_new_accounts=False, allow_self_destruct=False, inline_code_frac=0.02)


def c2_block(seed, scale):
    """One synthetic C2 block as a FlatBlock; cached under /tmp."""
    from proof_protocol_decoder_b200 import synth

    cache = f"/tmp/ppd_c2_seed{seed}_scale{scale}.flat"
    if os.path.exists(cache):
        return open(cache, "rb").read()
    blk = synth.gen_block(
        seed,
        n_accounts=max(10, int(20000 * scale)),
        n_txns=max(2, int(200 * scale)),
        accounts_per_txn=(80, 120) if scale >= 0.5 else (max(2, int(80 * scale * 4)), max(4, int(120 * scale * 4))),
        **C2_PARAMS,
    )
    f = blk.flat
    try:
        tmp = cache + f".{os.getpid()}.tmp"
        open(tmp, "wb").write(f)
        os.replace(tmp, cache)
    except OSError:
        pass
    return f


def _c2_block_job(a):
    c2_block(*a)
    return 0


def c2_blocks(seeds, scale, procs):
    """The blocks of the given seeds; missing ones are generated in parallel worker processes (the
    generator is pure Python: ~10 s per full-size block)."""
    missing = [s for s in seeds if not os.path.exists(f"/tmp/ppd_c2_seed{s}_scale{scale}.flat")]
    if len(missing) > 1 and procs > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(min(procs, len(missing))) as pool:
            pool.map(_c2_block_job, [(s, scale) for s in missing])
    return [c2_block(s, scale) for s in seeds]


def oracle_time_block(flat_bytes, repeats, threads):
    """Decode `flat_bytes` `repeats` times on each of `threads` host threads (ctypes releases the GIL).
    Returns (wall seconds, stats of one decode)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ppd_oracle_lib

    oracle = ppd_oracle_lib.load()
    stats = {}

    def work():
        for _ in range(repeats):
            oracle.block_decode(flat_bytes)
        stats.update(oracle.last_stats())

    ts = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, stats


def run_reference(args, rank, world):
    """The reference's CPU algorithm for the path (the oracle: a C++ port of the reference's Rust,
    which cannot be compiled here) on all host cores; each step decodes one bounded sample block per
    core.  Rank 0 alone works."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_scale = args.ref_scale
    flat_bytes = c2_block(2, sample_scale)
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_time_block(flat_bytes, 1, cores)
    t_total, st = 0.0, {}
    for _ in range(args.steps):
        t, st = oracle_time_block(flat_bytes, 1, cores)
        t_total += t
    nodes = st["nodes_hashed"] * cores * args.steps
    value = nodes / t_total
    line = {
        "impl": "reference",
        "metric": "mpt_nodes_keccak_hashed_per_sec",
        "value": value,
        "unit": "nodes/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": "C2 mainnet-shaped block (BASELINE.json configs[1])", "sample": f"{sample_scale:g}x-scale C2 block per core per step", "l2": "n/a (CPU)"},
        "blocks_per_sec": cores * args.steps / t_total,
        "cpu_baseline": {
            "value": value, "unit": "nodes/s", "cores": cores, "kind": "port",
            "sample": f"one {sample_scale:g}x-scale C2 block ({st['nodes_hashed']} node hashes) per core per step; C++ restatement of the reference algorithm (the Rust reference cannot be built in this image)",
        },
        "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def c5_sweep(ctx, sizes, peaks, sm_mhz):
    """Config 5: state-trie rehash over sorted leaves already resident in HBM (structure built and
    hashed on the GPU).  Reported beside the headline to show the hashing kernels at scale."""
    import torch

    out = []
    for n in sizes:
        try:
            g = torch.Generator(device="cuda")
            g.manual_seed(5)
            keys = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
            hi = keys[:, :8].to(torch.int64)
            k64 = torch.zeros(n, dtype=torch.int64, device="cuda")
            for b in range(8):
                k64 = (k64 << 8) | hi[:, b]
            order = torch.argsort((k64 >> 1) & 0x7FFFFFFFFFFFFFFF)  # top 63 bits, unsigned order
            del hi, k64
            keys = keys[order].contiguous()
            del order
            lens = torch.randint(70, 81, (n,), dtype=torch.int64, device="cuda", generator=g)
            val_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
            val_off[1:] = torch.cumsum(lens, 0)
            vb = int(val_off[-1].item())
            vals = torch.randint(0, 256, (vb,), dtype=torch.uint8, device="cuda", generator=g)
            torch.cuda.synchronize()
            best = None
            for _ in range(3):
                ctx.trie_root_sorted_leaves_dev(keys.data_ptr(), val_off.data_ptr(), vals.data_ptr(), n, vb)
                st = ctx.stats()
                if best is None or st["gpu_ms"] < best["gpu_ms"]:
                    best = st
            sec = best["gpu_ms"] / 1e3
            alu_peak = ALU_LANES_PER_SM_CLK * N_SM * (sm_mhz or peaks["sm_max_mhz"]) * 1e6
            out.append({
                "leaves": n, "nodes_hashed": best["nodes_hashed"], "permutations": best["node_permutations"], "gpu_ms": best["gpu_ms"],
                "nodes_per_sec": best["nodes_hashed"] / sec, "perms_per_sec": best["node_permutations"] / sec,
                "alu_frac": best["node_permutations"] * ALU_OPS_PER_PERM / sec / alu_peak,
                "hbm_frac": (best["node_bytes"] + 32 * best["nodes_hashed"]) / sec / 1e9 / peaks["hbm_gbs"],
                "note": "includes building the trie structure on the GPU (LCP, branch discovery, counting sort)",
            })
            del keys, vals, val_off, lens
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001 — the sweep is supplementary; never lose the headline line
            out.append({"leaves": n, "error": str(e)[:200]})
            break
    return out


def run_b200(args, rank, world, local_rank):
    # host threads: this rank's share of the box's cores; one block per host thread per step
    cores = os.cpu_count() or 1
    # two host threads per core of this rank's share: a thread sleeps while its lane's copies and kernels run
    threads = max(1, min(32, 2 * cores // world))
    os.environ.setdefault("PPD_HOST_THREADS", str(threads))
    threads = int(os.environ["PPD_HOST_THREADS"])
    n_blocks = min(64, args.blocks_per_step or 64)  # per GPU, whatever the world size (weak scaling); the library keeps up to 64 lanes resident
    seeds = [2 + rank * n_blocks + j for j in range(n_blocks)]
    # generated before CUDA is touched (worker processes are forked); ranks generate their own blocks
    flats = c2_blocks(seeds, args.scale, max(1, cores // world))
    flat_total = sum(len(f) for f in flats)

    import torch

    from proof_protocol_decoder_b200.lib import Context

    peaks = load_peaks()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = Context(local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    # the step's inputs live in page-locked host memory (ppd_alloc_pinned), as a caller that serialises its
    # BlockTrace for the library would place them; every step copies them to the device again
    flats = [ctx.pinned_copy(f) for f in flats]

    def decode_step():
        outs = ctx.blocks_decode_batch_view(flats)
        total = 0
        for o in outs:
            if isinstance(o, Exception):
                raise o
            total += o.nbytes + o.view[0] + o.view[o.nbytes - 1]  # read the result
            o.close()
        return total

    # ---- warm-up (also leaves the arenas resident for the device-resident measurement) ----
    for _ in range(max(3, args.warmup)):
        ir_len = decode_step()
    st = ctx.stats()
    for _ in range(max(3, args.warmup)):
        ctx.replay_last_hashing()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident: every kernel of the batch on the arenas already in HBM ----
    barrier()
    dev_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        dev_ms += ctx.replay_last_hashing()
    barrier()
    dev_ms = max_over_ranks(dev_ms)
    # ---- the same for the witness parse / arena kernels (ppd_parse.cu), witnesses resident in HBM ----
    parse_ms = 0.0
    if st["witnesses_on_gpu"]:
        for _ in range(max(3, args.warmup)):
            ctx.replay_last_parse()
        barrier()
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            parse_ms += ctx.replay_last_parse()
        barrier()
        parse_ms = max_over_ranks(parse_ms)
    # ---- end to end through the C ABI with host buffers ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ir_len = decode_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    clocks = sampler.stop()
    st = ctx.stats()
    # latency of one block alone (all host threads on its IR dump)
    single = []
    for _ in range(3):
        t1 = time.perf_counter()
        with ctx.block_decode_view(flats[0]) as v:
            _ = v.view[0]
        single.append(time.perf_counter() - t1)

    nodes_all = sum_over_ranks(float(st["nodes_hashed"]))
    perms_all = sum_over_ranks(float(st["node_permutations"] + st["key_permutations"]))
    keys_all = sum_over_ranks(float(st["key_hashes"]))
    dev_s_per_step = dev_ms / 1e3 / args.steps
    e2e_s_per_step = e2e_s / args.steps
    sm_mhz = clocks.get("sm_mhz")
    alu_peak = ALU_LANES_PER_SM_CLK * N_SM * (sm_mhz or peaks["sm_max_mhz"]) * 1e6  # instr/s on one GPU
    algo_bytes = st["node_bytes"] + 32 * st["nodes_hashed"]  # L + 32 per hashed node (this rank)
    achieved_gbs = algo_bytes / dev_s_per_step / 1e9
    traffic, traffic_src = load_traffic()
    level_launches = max(1, int(st["level_launches"]))
    parse_s_per_step = parse_ms / 1e3 / args.steps
    wit_bytes_all = sum_over_ranks(float(st["witness_bytes"]))
    wit_ins_all = sum_over_ranks(float(st["witness_instructions"]))
    line = {
        "metric": "mpt_nodes_keccak_hashed_per_sec",
        "value": nodes_all / dev_s_per_step,
        "unit": "nodes/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": 1e3 * dev_s_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {
            "workload": f"C2 mainnet-shaped blocks (BASELINE.json configs[1]): 20k touched accounts in a virtual 16^7-account state, 200 txns each; a batch of {n_blocks} independent blocks per GPU per step, one lane (stream + pools) per block, {threads} host thread(s) per GPU",
            "scale": args.scale,
            "blocks_per_step_per_gpu": n_blocks,
            "host_threads_per_gpu": threads,
            "flat_block_bytes_per_step": flat_total,
            "ir_dump_bytes_per_step": ir_len,
            "arena_nodes": st["arena_nodes"],
            "levels": st["levels"],
            "l2": "flushed between timed device-resident steps (256 MiB write)",
            "parallelism": f"blocks sharded over {world} GPU(s), no data-path collective",
        },
        "blocks_per_sec": world * n_blocks / (dev_s_per_step + parse_s_per_step),
        "blocks_per_sec_note": "device-resident: the blocks of a step / (device time of the witness parse + arena kernels, plus device time of the hashing kernels), each replayed with all lanes concurrent",
        "blocks_per_sec_hashing_only": world * n_blocks / dev_s_per_step,
        "permutations_per_sec": perms_all / dev_s_per_step,
        "nodes_hashed_per_step": nodes_all,
        "key_hashes_per_step": keys_all,
        "e2e": {
            "value": nodes_all / e2e_s_per_step,
            "unit": "nodes/s",
            "blocks_per_sec": world * n_blocks / e2e_s_per_step,
            "ms_per_step": 1e3 * e2e_s_per_step,
            "single_block_latency_ms": 1e3 * min(single),
            "h2d_bytes_per_step": sum_over_ranks(float(st["h2d_bytes"])),
            "d2h_bytes_per_step": sum_over_ranks(float(st["d2h_bytes"])),
            "note": "ppd_blocks_decode_batch: FlatBlocks (host) -> IrDumps (host); includes witness parse, trie shaping and IR serialisation on the host threads and every host<->device copy",
        },
        "gpu_launches": int(st["kernel_launches"]) * args.steps,
        "clocks": clocks,
        "roofline": {
            "bound": "hbm",
            "kernel": "hash_level_kernel (all level launches of the batch, lanes concurrent)",
            "achieved": achieved_gbs,
            "peak": peaks["hbm_gbs"],
            "unit": "GB/s",
            "frac": achieved_gbs / peaks["hbm_gbs"],
            "traffic": traffic,
            "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": algo_bytes / level_launches,
            "launches_per_step": level_launches,
            "peak_source": peaks["source"],
            "note": "achieved = sum over the step's level launches of (L + 32) bytes per hashed node / device time of the step (lanes run concurrently, so a per-launch duration does not exist); traffic = mean DRAM bytes per hash_level_kernel launch under ncu. Keccak is bound by the integer ALU pipe, not HBM (SURVEY.md 8d): see roofline_alu",
        },
        "parse": {
            "kernels": "ppd_parse.cu: tile_exit / group_exit / top_chain / tile_entry / tile_mark / scans / link / shape / emit / climb (all lanes concurrent)",
            "witnesses_on_gpu_per_step": int(sum_over_ranks(float(st["witnesses_on_gpu"]))),
            "witness_bytes_per_step": wit_bytes_all,
            "instructions_per_step": wit_ins_all,
            "ms_per_step": 1e3 * parse_s_per_step,
            "bound": "hbm",
            "achieved": (7.0 * wit_bytes_all / world / parse_s_per_step / 1e9) if parse_s_per_step else None,
            "peak": peaks["hbm_gbs"],
            "unit": "GB/s",
            "frac": (7.0 * wit_bytes_all / world / parse_s_per_step / 1e9 / peaks["hbm_gbs"]) if parse_s_per_step else None,
            "instructions_per_sec": (wit_ins_all / parse_s_per_step) if parse_s_per_step else None,
            "note": "algorithmic bytes = 7 per witness byte for the boundary search (read the byte, write and re-read its 4-byte exit link, write its 2-byte step link) -- the instruction-level arrays behind it are 30x smaller; the chain walks between tiles are latency-bound by construction (DESIGN.md 5)",
        },
        "roofline_alu": {
            "bound": "alu-pipe (LOP3/SHF)",
            "achieved": (st["node_permutations"] + st["key_permutations"]) * ALU_OPS_PER_PERM / dev_s_per_step / 1e12,
            "peak": alu_peak / 1e12,
            "unit": "Tinstr/s",
            "frac": (st["node_permutations"] + st["key_permutations"]) * ALU_OPS_PER_PERM / dev_s_per_step / alu_peak,
            "model": f"{ALU_OPS_PER_PERM} ALU instr per permutation; peak = {ALU_LANES_PER_SM_CLK} lanes/clk/SM x {N_SM} SMs x SM clock under load",
        },
    }
    if rank == 0 and world == 1:
        sample = c2_block(2, args.ref_scale)
        t, ost = oracle_time_block(sample, 1, 1)
        line["cpu_baseline"] = {
            "value": ost["nodes_hashed"] / t,
            "unit": "nodes/s",
            "cores": 1,
            "kind": "port",
            "blocks_per_sec_at_sample_scale": 1.0 / t,
            "sample": f"one {args.ref_scale:g}x-scale C2 block ({ost['nodes_hashed']} node hashes, {t:.2f} s) on 1 core; C++ restatement of the reference algorithm (the Rust reference cannot be built in this image)",
        }
        if not args.no_sweep:
            line["c5_sweep"] = c5_sweep(ctx, [int(x) for x in args.sweep.split(",") if x], peaks, sm_mhz)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the C2 block size (1.0 = the named config)")
    ap.add_argument("--blocks-per-step", type=int, default=0, help="blocks per GPU per step (default 64, at most 64)")
    ap.add_argument("--ref-scale", type=float, default=0.1, help="size of the bounded CPU sample block")
    ap.add_argument("--sweep", default="1000000,10000000", help="config-5 leaf counts measured beside the headline at N=1")
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
