#!/bin/bash
# Round-2 GPU session C (1 GPU): tests, bench, loss-kernel tuning variants, post-processing timing + ncu.
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 $OUT/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 300 $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== loss timing variants"
for suf in "" _m6 _m4 _u8 _u2m8 _u2m6; do RN_LIB_SUFFIX=$suf python tools/loss_time.py 2>&1 | tail -1; done
echo "== pp timing" ; python tools/pp_time.py 2>&1 | tail -3
echo "== ncu launch list (graph step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_step.csv \
    python bench.py --steps 2 --warmup 3 --only-step > $OUT/bench_under_ncu.log 2>&1 ; echo "ncu list rc=$?"
echo "== ncu full: post-processing kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_filter|lazy' \
    -c 12 -o $OUT/prof_pp python tools/pp_time.py > $OUT/ncu_pp.log 2>&1 ; echo "ncu full rc=$?"
ncu -i $OUT/prof_pp.ncu-rep --page raw --csv > $OUT/prof_pp_raw.csv 2>/dev/null
ls -la $OUT | head -30
