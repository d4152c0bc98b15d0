#!/bin/bash
# Round-2 GPU session G (1 GPU): tests, fused step timing, bench, profiles (launch lists, ncu --set full), sanitizers.
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -4 $OUT/pytest_gpu.log
echo "== only-step fused (default)" ; timeout 300 python bench.py --steps 200 --warmup 10 --only-step 2> $OUT/step_fused.err | tee $OUT/step_fused.json
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 200 $OUT/bench.json ; tail -3 $OUT/bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ; echo "ref rc=$?" ; tail -c 300 $OUT/bench_ref.json
echo "== ncu launch list (graph step, fused)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_step.csv \
    python bench.py --steps 4 --warmup 3 --only-step > $OUT/bench_under_ncu.log 2>&1 ; echo "ncu list rc=$?"
echo "== ncu launch list (default bench, steps 2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_under_ncu2.log 2>&1 ; echo "ncu list2 rc=$?"
echo "== ncu full: graph step kernels (fused loss, lazy2, match)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loss_kernel|match_kernel|lazy2|prep_kernel' \
    -s 12 -c 8 -o $OUT/prof_step python bench.py --steps 2 --warmup 3 --only-step > $OUT/ncu_step.log 2>&1 ; echo "ncu step rc=$?"
ncu -i $OUT/prof_step.ncu-rep --page raw --csv > $OUT/prof_step_raw.csv 2>/dev/null
echo "== ncu full: loss kernels alone (fwd+grad, fwd-only)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loss_kernel|loss_finalize' \
    -c 16 -o $OUT/prof_loss env RN_REPS=2 python tools/loss_time.py > $OUT/ncu_loss.log 2>&1 ; echo "ncu loss rc=$?"
ncu -i $OUT/prof_loss.ncu-rep --page raw --csv > $OUT/prof_loss_raw.csv 2>/dev/null
echo "== ncu full: post-processing kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_filter|lazy' \
    -c 9 -o $OUT/prof_pp python tools/pp_time.py > $OUT/ncu_pp.log 2>&1 ; echo "ncu pp rc=$?"
ncu -i $OUT/prof_pp.ncu-rep --page raw --csv > $OUT/prof_pp_raw.csv 2>/dev/null
echo "== sanitizers" ; timeout 1500 bash tools/gpu_sanitize.sh
rm -f $OUT/*.ncu-rep
ls -la $OUT | head -40
