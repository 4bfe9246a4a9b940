#!/bin/bash
# Round 2, second session, final call: smoke, the whole GPU suite, the default bench line, the launch list of the bench
# command, the launch-rate probe and the pipeline timeline, all on the round's last code.
set -u
TAG=r02b
RAW=${RAW:-/tmp/ppd_cap}
OUT=gpurun_out/profiles_r02f
mkdir -p $RAW $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $OUT/smoke.log | cut -c1-120)"
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "gpu tests: rc=$? $(tail -1 $OUT/tests.log)"
t0=$(date +%s)
python bench.py > $OUT/${TAG}_bench_default.json 2> $RAW/bench.err; echo "default bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -2 $RAW/bench.err
python tools/bench_summary.py $OUT/${TAG}_bench_default.json
tools/launch_rate > $OUT/${TAG}_launch_rate.txt 2>&1; cat $OUT/${TAG}_launch_rate.txt
( cd tools && python dev_e2e_trace.py 128 $RAW/${TAG}_trace.csv > ../$OUT/${TAG}_pipeline_trace.txt 2>&1 ); head -30 $OUT/${TAG}_pipeline_trace.txt
LIST="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
B="python bench.py --steps 2 --warmup 3 --no-sweep --no-split --blocks-per-step 4 --e2e-mult 1"
# (in the round's last call this ncu pass did not finish within the 18 minutes that were left of the GPU budget -- the
#  earlier session's same pass took about two minutes; it is bounded now, and everything above had already been brought back)
$B > $OUT/${TAG}_bench_plain.log 2> $RAW/bench_plain.err && timeout 300 ncu $LIST -c 8000 --log-file $RAW/${TAG}_launches_c2.csv $B > $RAW/ncu_c2.log 2>&1
SRC=$RAW DST=$OUT python profiles/summarize.py $TAG > /dev/null 2>&1
head -50 $OUT/${TAG}_launches_c2_summary.txt
ls $OUT
