#!/bin/bash
# Round-2 evidence, run on the B200 box (gpurun):  bash profiles/capture_r02.sh
# Every ncu pass runs AFTER the same command has exited 0 without ncu.  Raw captures go to $RAW (outside gpurun_out/);
# the small summaries to gpurun_out/profiles_r02/, from where they are copied into profiles/.
set -u
TAG=r02
RAW=${RAW:-/tmp/ppd_cap}
OUT=gpurun_out/profiles_$TAG
mkdir -p $RAW $OUT
LIST="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
# 1. launch list of the bench command at 4 blocks per step (all kernels; serialised by ncu: shares, not absolutes)
B="python bench.py --steps 2 --warmup 3 --no-sweep --no-split --blocks-per-step 4 --e2e-mult 1"
$B > $RAW/${TAG}_bench_plain.log 2> $RAW/${TAG}_bench_plain.err || { echo "plain bench failed"; tail -5 $RAW/${TAG}_bench_plain.err; exit 1; }
ncu $LIST -c 8000 --log-file $RAW/${TAG}_launches_c2.csv $B > $RAW/${TAG}_ncu_c2.log 2>&1
# 2. full capture of the level-hashing kernel: one block per step, the level launches of the first timed replay
#    (31 level launches per decode / replay; 3 warm-up decodes + 3 warm-up replays come first)
B1="python bench.py --steps 1 --warmup 3 --no-sweep --no-split --blocks-per-step 1 --e2e-mult 1"
NL=${NL:-31}  # level launches per block decode (profiles/r02_launches_block_summary.txt)
echo "level launches per block: $NL"
PPD_HOST_THREADS=1 $B1 > /dev/null 2>&1 && PPD_HOST_THREADS=1 ncu --set full --clock-control none --import-source on -k regex:hash_level_kernel \
    --launch-skip $((6 * NL)) --launch-count $NL -o $RAW/${TAG}_hash_level_full $B1 > $RAW/${TAG}_ncu_full.log 2>&1
# 3. full capture of two launches (16 txns each) of the txn loop kernel of one block
PR="python profiles/run_parse.py"
$PR 3 > $RAW/${TAG}_block_plain.log 2> $RAW/${TAG}_block_plain.err && \
  ncu --set full --clock-control none --import-source on -k regex:txn_loop_kernel --launch-skip 30 --launch-count 2 -o $RAW/${TAG}_txn_loop_full $PR 3 > $RAW/${TAG}_ncu_loop.log 2>&1
# 3b. full capture of the IR sizing / emit kernels of the third decode
ncu --set full --clock-control none --import-source on -k "regex:ir_size_kernel|ir_emit_kernel" --launch-skip 4 --launch-count 2 -o $RAW/${TAG}_dump_full $PR 3 > $RAW/${TAG}_ncu_dump.log 2>&1
# 4. per-kernel launch list of one block decode (warm)
ncu $LIST -c 3000 --log-file $RAW/${TAG}_launches_block.csv $PR 2 > /dev/null 2>&1
python profiles/summarize_block.py $RAW/${TAG}_launches_block.csv 2 "one config-2 block (python profiles/run_parse.py 2, the second decode)" > $OUT/${TAG}_launches_block_summary.txt
# 5. integer-pipe microbenchmark (the ALU roofline's denominator)
python profiles/microbench.py > $OUT/${TAG}_microbench.txt 2>&1
# 6. the pipeline's stage timeline under load (PPD_TRACE)
( cd tools && python dev_e2e_trace.py 128 $RAW/${TAG}_trace.csv > ../$OUT/${TAG}_pipeline_trace.txt 2>&1 )
SRC=$RAW DST=$OUT python profiles/summarize.py $TAG
cp $RAW/${TAG}_bench_plain.log $RAW/${TAG}_block_plain.log $OUT/ 2>/dev/null
ls -la $OUT
