#!/bin/bash
# Round 2, second session: the reworked witness-parse kernels (tile_exit on lists, 8-wide pyramid level, emit /
# ins_info / shape with the keyed instructions on a list, thread-per-tile marking) against their round-1 forms
# (PPD_PARSE_V1=31).  Run on the B200 box:  bash profiles/capture_r02b.sh
# Every ncu pass runs AFTER the same command has exited 0 without ncu.
set -u
TAG=r02b
RAW=${RAW:-/tmp/ppd_cap}
OUT=gpurun_out/profiles_$TAG
mkdir -p $RAW $OUT
LIST="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"

# 1. parity: the whole GPU suite with the new kernels, then the parse tests with the host builder's tries compared node by node
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/tests_new.log 2>&1
echo "gpu tests (new kernels): rc=$?"; tail -3 $OUT/tests_new.log
PPD_VERIFY_GPU_PARSE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_noncanonical_witness.py -m gpu -x -q > $OUT/tests_verify.log 2>&1
echo "gpu parse tests under PPD_VERIFY_GPU_PARSE: rc=$?"; tail -3 $OUT/tests_verify.log
if ! tail -1 $OUT/tests_new.log | grep -q passed || tail -1 $OUT/tests_new.log | grep -q failed; then
  # which kernel?  each round-1 form back in turn
  for m in 1 2 4 8 16 31; do
    PPD_PARSE_V1=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/tests_v1mask$m.log 2>&1
    echo "mask $m: rc=$? $(tail -1 $OUT/tests_v1mask$m.log)"
  done
fi

# 2. one block alone: parse device time, new and old; per-kernel launch lists
PR="python profiles/run_parse.py"
$PR 3 > $OUT/${TAG}_block_plain_new.log 2> $RAW/block_new.err; echo "run_parse new rc=$?"
PPD_PARSE_V1=31 $PR 3 > $OUT/${TAG}_block_plain_v1.log 2> $RAW/block_v1.err; echo "run_parse v1 rc=$?"
ncu $LIST -c 3000 --log-file $RAW/${TAG}_launches_block.csv $PR 2 > /dev/null 2>&1
python profiles/summarize_block.py $RAW/${TAG}_launches_block.csv 2 "one config-2 block (python profiles/run_parse.py 2, the second decode), reworked parse kernels" > $OUT/${TAG}_launches_block_summary.txt
PPD_PARSE_V1=31 ncu $LIST -c 3000 --log-file $RAW/${TAG}_launches_block_v1.csv $PR 2 > /dev/null 2>&1
python profiles/summarize_block.py $RAW/${TAG}_launches_block_v1.csv 2 "one config-2 block (PPD_PARSE_V1=31 python profiles/run_parse.py 2, the second decode), round-1 parse kernels" > $OUT/${TAG}_launches_block_v1_summary.txt

# 3. the bench, new and old (replays + end to end)
B="python bench.py --steps 3 --warmup 3 --no-sweep --no-split"
$B > $OUT/${TAG}_bench_new.json 2> $RAW/bench_new.err; echo "bench new rc=$?"
PPD_PARSE_V1=31 $B > $OUT/${TAG}_bench_v1.json 2> $RAW/bench_v1.err; echo "bench v1 rc=$?"
python tools/bench_summary.py $OUT/${TAG}_bench_new.json $OUT/${TAG}_bench_v1.json 2>/dev/null | head -40

# 4. full capture of the parse kernels of the second decode of one block
ncu --set full --clock-control none --import-source on -k "regex:tile_exit_kernel|link_kernel16|emit_kernel|ins_info_kernel|shape_kernel|tile_mark|climb_kernel|group_exit|tile_entry" \
    --launch-skip 9 --launch-count 9 -o $RAW/${TAG}_parse_full $PR 2 > $RAW/${TAG}_ncu_parse.log 2>&1
SRC=$RAW DST=$OUT python profiles/summarize.py $TAG   # writes ${TAG}_parse_full.txt from the .ncu-rep
tail -3 $RAW/*.err 2>/dev/null | tail -20
ls -la $OUT
