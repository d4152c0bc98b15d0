#!/bin/bash
# Round-2 GPU session A (1 GPU): topology, smoke, full GPU test suite, both bench arms.
OUT=gpurun_out
mkdir -p $OUT
{
  nvidia-smi --query-gpu=index,name,pci.bus_id,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
  nvidia-smi topo -m
  lscpu | head -30
  ls /sys/devices/system/node/
  for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ]; then echo "$d numa_node=$(cat $d/numa_node)"; fi; done
  grep -i allowed /proc/self/status
  nproc
  free -g | head -3
} > $OUT/topology.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 $OUT/smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --durations=15 > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -45 $OUT/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -c 1500 $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ; echo "ref rc=$?" ; tail -c 800 $OUT/bench_ref.json
ls -la $OUT | head -30
