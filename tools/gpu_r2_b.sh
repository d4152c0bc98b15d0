#!/bin/bash
# Round-2 GPU session B (1 GPU): GPU tests, bench, ncu launch list + metrics of the loss kernels.
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 $OUT/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 600 $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== loss timing" ; python tools/loss_time.py 2>&1 | tail -1
echo "== ncu launch list (graph step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_step.csv \
    python bench.py --steps 2 --warmup 3 --only-step > $OUT/bench_under_ncu.log 2>&1 ; echo "ncu list rc=$?"
echo "== ncu full: loss kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loss_kernel|match_kernel|loss_finalize' \
    -c 24 -o $OUT/prof_loss env RN_REPS=2 python tools/loss_time.py > $OUT/ncu_loss.log 2>&1 ; echo "ncu full rc=$?"
ncu -i $OUT/prof_loss.ncu-rep --page raw --csv > $OUT/prof_loss_raw.csv 2>/dev/null
ls -la $OUT | head -30
