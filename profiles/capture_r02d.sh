#!/bin/bash
# Round 2, second session, third call: emit split into emit_kernel + emit_keyed_kernel (list made by shape_kernel),
# ins_info back to one thread per instruction, climb without the no-op atomicMax; Direct pre-images (kind 2).
set -u
TAG=r02d
RAW=${RAW:-/tmp/ppd_cap}
OUT=gpurun_out/profiles_r02b
mkdir -p $RAW $OUT
LIST="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/tests_d.log 2>&1
echo "gpu tests: rc=$? $(tail -1 $OUT/tests_d.log)"; grep -E "FAILED|Error|assert" $OUT/tests_d.log | head -10
PPD_VERIFY_GPU_PARSE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_noncanonical_witness.py tests/test_direct_pre_image.py -m gpu -x -q > $OUT/tests_d_verify.log 2>&1
echo "verify: rc=$? $(tail -1 $OUT/tests_d_verify.log)"
PR="python profiles/run_parse.py"
$PR 3 > $OUT/${TAG}_block_plain_new.log 2> $RAW/block_new.err; echo "run_parse rc=$?"; cat $OUT/${TAG}_block_plain_new.log | cut -c1-300
ncu $LIST -c 3000 --log-file $RAW/${TAG}_launches_block.csv $PR 2 > /dev/null 2>&1
python profiles/summarize_block.py $RAW/${TAG}_launches_block.csv 2 "one config-2 block (python profiles/run_parse.py 2, the second decode), parse kernels of the round's last form" > $OUT/${TAG}_launches_block_summary.txt
head -28 $OUT/${TAG}_launches_block_summary.txt
python bench.py --steps 3 --warmup 3 --no-sweep --no-split > $OUT/${TAG}_bench.json 2> $RAW/bench.err; echo "bench rc=$?"; tail -2 $RAW/bench.err
python tools/bench_summary.py $OUT/${TAG}_bench.json
ncu --set full --clock-control none --import-source on -k "regex:tile_exit_kernel|link_kernel16|emit_kernel|emit_keyed_kernel|ins_info_kernel|shape_kernel|tile_mark|climb_kernel" \
    --launch-skip 9 --launch-count 9 -o $RAW/r02b_parse_full $PR 2 > $RAW/ncu_parse.log 2>&1
SRC=$RAW DST=$OUT python profiles/summarize.py r02b > /dev/null 2>&1
cp $OUT/r02b_parse_full.txt $OUT/${TAG}_parse_full.txt 2>/dev/null; cat $OUT/${TAG}_parse_full.txt | cut -c1-200
