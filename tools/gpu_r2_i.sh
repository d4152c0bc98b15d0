#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest graph tests" ; timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -x > $OUT/pytest_graph.log 2>&1 ; echo "pytest rc=$?" ; tail -4 $OUT/pytest_graph.log
echo "== only-step pipeline (default)" ; timeout 300 python bench.py --steps 200 --warmup 10 --only-step 2> $OUT/step_pipe.err | tee $OUT/step_pipe.json
echo "== only-step no pipeline" ; RN_BENCH_NO_PIPELINE=1 timeout 300 python bench.py --steps 200 --warmup 10 --only-step 2> $OUT/step_nopipe.err | tee $OUT/step_nopipe.json
tail -3 $OUT/step_pipe.err
