#!/bin/bash
# ncu launch lists (the step alone, and the default bench command) + full capture of the step's kernels.
# Usage (GPU box, repo root): bash tools/gpu_ncu.sh ['<kernel regex>']
OUT=gpurun_out; mkdir -p $OUT
REGEX=${1:-'loss_kernel|score_filter|match_kernel|lazy_nms|loss_finalize|pack_targets'}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_step.csv \
    python bench.py --steps 4 --warmup 3 --only-step > $OUT/bench_under_ncu_step.log 2>&1 ; echo "ncu step list rc=$?"
python tools/launch_summary.py $OUT/launches_step.csv | head -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1 ; echo "ncu list rc=$?"
python tools/launch_summary.py $OUT/launches.csv | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" \
    -s 14 -c 7 -o $OUT/prof_step python bench.py --steps 1 --warmup 3 --only-step > $OUT/ncu_full.log 2>&1 ; echo "ncu full rc=$?"
ncu -i $OUT/prof_step.ncu-rep --page raw --csv > $OUT/prof_step_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/prof_step_raw.csv | head -150
