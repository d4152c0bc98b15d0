#!/bin/bash
# Round-2 GPU session F (1 GPU): tests, fused vs two-branch graph step, bench.
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -8 $OUT/pytest_gpu.log
echo "== only-step fused (default)" ; timeout 300 python bench.py --steps 200 --warmup 10 --only-step 2> $OUT/step_fused.err | tee $OUT/step_fused.json
echo "== only-step two branches" ; RN_BENCH_UNFUSED=1 timeout 300 python bench.py --steps 200 --warmup 10 --only-step 2> $OUT/step_unfused.err | tee $OUT/step_unfused.json
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 200 $OUT/bench.json ; tail -3 $OUT/bench.err
echo "== ncu launch list (graph step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_step.csv \
    python bench.py --steps 2 --warmup 3 --only-step > $OUT/bench_under_ncu.log 2>&1 ; echo "ncu list rc=$?"
ls -la $OUT | head
