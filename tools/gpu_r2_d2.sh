#!/bin/bash
# Round-2 multi-GPU session (N GPUs): multi-GPU tests, bench at N, exchange variants (only-step).
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topology_n$N.txt 2>&1
echo "== pytest multi" ; timeout 700 python -m pytest tests/test_gpu_multi.py -q -rA > $OUT/pytest_multi_n$N.log 2>&1 ; echo "pytest rc=$?" ; tail -12 $OUT/pytest_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "== bench N=$N" ; timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-levels > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err ; echo "bench rc=$?" ; tail -c 200 $OUT/bench_n$N.json ; tail -3 $OUT/bench_n$N.err
for mode in auto nccl; do
  echo "== only-step exchange=$mode"
  RN_BENCH_EXCHANGE=$mode timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 --only-step 2> $OUT/step_${mode}_n$N.err | tee $OUT/step_${mode}_n$N.json
done
echo "== only-step NO exchange (diagnostic)"
RN_BENCH_NO_EXCHANGE=1 timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 --only-step 2> $OUT/step_none_n$N.err | tee $OUT/step_none_n$N.json
ls $OUT | head -40
