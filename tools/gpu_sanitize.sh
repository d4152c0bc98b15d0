#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the small-shape GPU tests; summaries into gpurun_out/.
OUT=gpurun_out; mkdir -p $OUT
SEL='known_answers or random_small or edge_cases or lazy_fallback or box_coding or dense_focal or nms_segments or anchor_generator_api or pre_nms_topk_extension and 1-50'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_levels.py tests/test_gpu_graph.py -m gpu -x -q \
      -k "$SEL or 200 or (graph_step and 1-3) or train_only or (train_loss_call and (cfg1 or odd_A))" > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_$tool.log | tail -3
done
