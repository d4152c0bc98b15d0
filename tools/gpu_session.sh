#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list + full captures of the top kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh [quick|full]
MODE=${1:-full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 $OUT/smoke.log
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 $OUT/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 3000 $OUT/bench.json ; tail -5 $OUT/bench.err
if [ "$MODE" = "full" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1 ; echo "ncu list rc=$?"
  echo "== ncu full: loss / score_filter / match / nms"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loss_kernel|score_filter|match_kernel|nms_kernel|image_topk' \
      -s 12 -c 8 -o $OUT/prof_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1 ; echo "ncu full rc=$?"
fi
ls -la $OUT
