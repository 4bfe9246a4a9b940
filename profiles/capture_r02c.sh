#!/bin/bash
# Round 2, second session, second call: the bench with every lane's replay queued by its own host thread, and the
# end-to-end knob sweep (tools/knob_sweep.py).  Run on the B200 box:  bash profiles/capture_r02c.sh
set -u
OUT=gpurun_out/profiles_r02b
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_device_txn_loop.py -m gpu -x -q > $OUT/tests_c.log 2>&1
echo "gpu parity tests: rc=$? $(tail -1 $OUT/tests_c.log)"
python bench.py --steps 3 --warmup 3 --no-sweep --no-split > $OUT/r02b_bench_mt.json 2> /tmp/bench_mt.err; echo "bench rc=$?"; tail -2 /tmp/bench_mt.err
PPD_REPLAY_THREADS=1 python bench.py --steps 3 --warmup 3 --no-sweep --no-split > $OUT/r02b_bench_st.json 2> /tmp/bench_st.err; echo "bench (single-threaded replay) rc=$?"
python tools/bench_summary.py $OUT/r02b_bench_mt.json $OUT/r02b_bench_st.json
python tools/knob_sweep.py > $OUT/r02b_knobs.txt 2>&1
cat $OUT/r02b_knobs.txt
