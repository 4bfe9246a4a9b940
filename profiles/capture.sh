#!/bin/bash
# Run on the B200 box (gpurun): captures the ncu evidence bench.py's numbers are read against.
#   bash profiles/capture.sh <tag>        -> gpurun_out/<tag>_*.{csv,ncu-rep,log}
# Every ncu pass runs AFTER the same command has exited 0 without ncu.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-sweep"
$B > $OUT/${TAG}_bench_plain.log 2> $OUT/${TAG}_bench_plain.err || { echo "plain bench failed"; exit 1; }
# 1. launch list of the bench command (all kernels, serialised by ncu: shares, not absolutes)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/${TAG}_launches_c2.csv $B > $OUT/${TAG}_ncu_c2.log 2>&1
# 2. full capture of the dominant kernel: one block per step so that the launch order is known
#    (17 level launches per decode / replay; skip 3 warm-up decodes + 3 warm-up replays)
B1="python bench.py --steps 1 --warmup 3 --no-sweep --blocks-per-step 1"
PPD_HOST_THREADS=1 $B1 > /dev/null 2>&1 && PPD_HOST_THREADS=1 ncu --set full --clock-control none --import-source on -k regex:hash_level_kernel \
    --launch-skip 102 --launch-count 17 -o $OUT/${TAG}_hash_level_full $B1 > $OUT/${TAG}_ncu_full.log 2>&1
# 3. config 5 (sorted leaves, 10M): launch list + full capture of the two hashing kernels
C5="python profiles/run_c5.py 10000000 1"
$C5 > $OUT/${TAG}_c5_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_c5.csv $C5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hash_ -c 14 -o $OUT/${TAG}_c5_full $C5 > $OUT/${TAG}_ncu_c5_full.log 2>&1
ls -la $OUT | grep $TAG
