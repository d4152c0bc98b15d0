#!/bin/bash
# Round-2 8-GPU session: multi-GPU tests, bench at N=8 (+ e2e at N=4), exchange variants.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topology_n8.txt 2>&1
nproc >> $OUT/topology_n8.txt; free -g | head -2 >> $OUT/topology_n8.txt
echo "== pytest multi (8 GPUs)" ; timeout 600 python -m pytest tests/test_gpu_multi.py -q -rA > $OUT/pytest_multi_n8.log 2>&1 ; echo "pytest rc=$?" ; tail -8 $OUT/pytest_multi_n8.log
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612"
echo "== bench N=8" ; timeout 420 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --no-levels > $OUT/bench_n8.json 2> $OUT/bench_n8.err ; echo "bench rc=$?" ; tail -c 200 $OUT/bench_n8.json ; tail -3 $OUT/bench_n8.err
for mode in auto nccl; do
  echo "== only-step N=8 exchange=$mode"
  RN_BENCH_EXCHANGE=$mode timeout 200 $TR8 bench.py --gpus 8 --steps 200 --warmup 10 --only-step 2> $OUT/step_${mode}_n8.err | tee $OUT/step_${mode}_n8.json
done
echo "== only-step N=8 NO exchange (diagnostic)"
RN_BENCH_NO_EXCHANGE=1 timeout 200 $TR8 bench.py --gpus 8 --steps 200 --warmup 10 --only-step 2> $OUT/step_none_n8.err | tee $OUT/step_none_n8.json
echo "== bench N=4 (e2e)" ; timeout 300 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 --no-levels --no-extra --no-cpu-baseline > $OUT/bench_n4.json 2> $OUT/bench_n4.err ; echo "bench rc=$?" ; tail -c 200 $OUT/bench_n4.json
ls $OUT | head -40
