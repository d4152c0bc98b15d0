#!/bin/bash
# tests + host profile + bench (no ncu)
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/host_profile.py > $OUT/host_profile.log 2>&1 ; cat $OUT/host_profile.log | tail -14
timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 2500 $OUT/bench.json ; tail -5 $OUT/bench.err
