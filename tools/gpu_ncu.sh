#!/bin/bash
# ncu launch list of a short bench run + full capture of selected kernels.  Usage: gpu_ncu.sh '<kernel regex>'
OUT=gpurun_out; mkdir -p $OUT
REGEX=${1:-'loss_kernel|score_filter|match_kernel|nms_kernel|image_topk'}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1 ; echo "ncu list rc=$?"
python tools/launch_summary.py $OUT/launches.csv | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" \
    -s 12 -c 6 -o $OUT/prof_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1 ; echo "ncu full rc=$?"
ncu -i $OUT/prof_top.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/prof_raw.csv | head -150
