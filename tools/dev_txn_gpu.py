"""Development check of the device txn loop on a GPU box: a few blocks decoded by the device path and by the host
path, both compared byte for byte with the CPU oracle, with the stats of each."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PPD_GPU_PARSE_MIN_BYTES", "0")
os.environ.setdefault("PPD_GPU_DUMP_MIN_TOUCHED", "0")
import ppd_oracle_lib
from proof_protocol_decoder_b200 import synth
from proof_protocol_decoder_b200.lib import Context

ctx = Context(0)
oracle = ppd_oracle_lib.load()
cases = {
 'c1_a': dict(seed=1, n_accounts=1000, n_txns=10),
 'c1_b': dict(seed=2, n_accounts=300, n_txns=12, contract_frac=0.5, slots_hi=8, accounts_per_txn=(5,20), slot_writes=(0,12), zero_write_frac=0.4),
 'c1_c': dict(seed=3, n_accounts=50, n_txns=30, contract_frac=0.6, slots_hi=4, accounts_per_txn=(5,25), slot_writes=(0,12), zero_write_frac=0.5),
 'c2s': dict(seed=4, n_accounts=2000, n_txns=20, contract_frac=0.15, slots_hi=256, virtual_depth=7, accounts_per_txn=(30,60), slot_reads=(0,3), slot_writes=(0,3), allow_new_accounts=False, allow_self_destruct=False, inline_code_frac=0.02),
 'tiny': dict(seed=5, n_accounts=3, n_txns=6, contract_frac=1.0, slots_hi=2, accounts_per_txn=(1,3), zero_write_frac=0.5),
}
bad = 0
for name, kw in cases.items():
    kw = dict(kw); seed = kw.pop('seed')
    blk = synth.gen_block(seed, **kw)
    want = oracle.block_decode(blk.flat)
    for mode in ("device", "host"):
        if mode == "host": os.environ["PPD_HOST_TXN"] = "1"
        else: os.environ.pop("PPD_HOST_TXN", None)
        t0 = time.time(); got = ctx.block_decode(blk.flat); dt = time.time() - t0
        st = ctx.stats()
        ok = got == want
        bad += not ok
        print(f"{name:6s} {mode:6s} {'OK ' if ok else 'DIFF'} {dt*1e3:8.2f} ms loops_on_gpu={st['txn_loops_on_gpu']} nodes_hashed={st['nodes_hashed']} arena={st['arena_nodes']} launches={st['kernel_launches']} txn_ms={st['txn_gpu_ms']:.3f} hash_ms={st['gpu_ms']:.3f} dump_ms={st['dump_gpu_ms']:.3f} h2d={st['h2d_bytes']:.0f} d2h={st['d2h_bytes']:.0f}", flush=True)
        if not ok:
            n = min(len(got), len(want)); i = next((k for k in range(n) if got[k] != want[k]), n)
            print(f"   lens {len(got)} {len(want)} first diff at {i}")
os.environ.pop("PPD_HOST_TXN", None)
print("bad", bad)
sys.exit(1 if bad else 0)
