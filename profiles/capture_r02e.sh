#!/bin/bash
# tile_exit with four resident thread blocks per SM (PPD_TILE_EXIT_OCC=4) against three: one block alone, and parity.
set -u
OUT=gpurun_out/profiles_r02b
mkdir -p $OUT
PR="python profiles/run_parse.py"
$PR 4 > $OUT/r02e_block_occ3.log 2>/dev/null; echo "occ3: $(cut -c1-230 $OUT/r02e_block_occ3.log)"
PPD_TILE_EXIT_OCC=4 $PR 4 > $OUT/r02e_block_occ4.log 2>/dev/null; echo "occ4: $(cut -c1-230 $OUT/r02e_block_occ4.log)"
$PR 4 > $OUT/r02e_block_occ3b.log 2>/dev/null; echo "occ3: $(cut -c1-230 $OUT/r02e_block_occ3b.log)"
PPD_TILE_EXIT_OCC=4 $PR 4 > $OUT/r02e_block_occ4b.log 2>/dev/null; echo "occ4: $(cut -c1-230 $OUT/r02e_block_occ4b.log)"
PPD_TILE_EXIT_OCC=4 PPD_VERIFY_GPU_PARSE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_noncanonical_witness.py -m gpu -x -q > $OUT/tests_e.log 2>&1
echo "parity under occ4: rc=$? $(tail -1 $OUT/tests_e.log)"
LIST="--metrics gpu__time_duration.sum --clock-control none --csv"
PPD_TILE_EXIT_OCC=4 ncu $LIST -k regex:tile_exit -c 4 --log-file /tmp/te4.csv $PR 2 > /dev/null 2>&1; grep tile_exit /tmp/te4.csv | tail -2 | cut -c1-200
ncu $LIST -k regex:tile_exit -c 4 --log-file /tmp/te3.csv $PR 2 > /dev/null 2>&1; grep tile_exit /tmp/te3.csv | tail -2 | cut -c1-200
