#!/bin/bash
# Round-2 GPU session E (1 GPU): lazy2 phase timing, pp timing, tests, bench.
OUT=gpurun_out
mkdir -p $OUT
echo "== lazy2 phase timing" ; python tools/lazy_timing.py > $OUT/lazy2_timing.log 2>&1 ; grep "lazy2 img" $OUT/lazy2_timing.log | tail -16
echo "== pp timing" ; python tools/pp_time.py 2>&1 | tail -3
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $OUT/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 300 $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== loss timing" ; python tools/loss_time.py 2>&1 | tail -1
ls -la $OUT | head -30
