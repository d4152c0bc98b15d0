#!/bin/bash
# Round-2 GPU session D (N GPUs, default 2): full GPU suite (multi-GPU tests included), bench at 1 and N, exchange variants.
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topology_n$N.txt 2>&1
echo "== pytest -m gpu (all, $N GPUs visible)" ; timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_n$N.log 2>&1 ; echo "pytest rc=$?" ; tail -25 $OUT/pytest_gpu_n$N.log
echo "== bench N=1" ; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 200 $OUT/bench.json ; tail -3 $OUT/bench.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "== bench N=$N" ; timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err ; echo "bench rc=$?" ; tail -c 200 $OUT/bench_n$N.json ; tail -5 $OUT/bench_n$N.err
for mode in auto nccl; do
  echo "== only-step exchange=$mode"
  RN_BENCH_EXCHANGE=$mode timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --only-step 2> $OUT/step_${mode}_n$N.err | tee $OUT/step_${mode}_n$N.json
done
echo "== only-step NO exchange (diagnostic)"
RN_BENCH_NO_EXCHANGE=1 timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --only-step 2> $OUT/step_none_n$N.err | tee $OUT/step_none_n$N.json
echo "== only-step N=1 reference point"
timeout 600 python bench.py --steps 200 --warmup 10 --only-step 2> $OUT/step_n1.err | tee $OUT/step_n1.json
ls -la $OUT | head -40
