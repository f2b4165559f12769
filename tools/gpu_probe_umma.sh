#!/bin/bash
# tcgen05 / TMEM probe on one GPU: every test in its own process, with a timeout.
O=gpurun_out/${1:-umma}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/gpu.txt 2>&1
for t in ${2:-1 2 3 4 5 7 6}; do
  echo "== test $t" | tee -a $O/umma_probe.txt
  timeout 120 tools/probes/umma_probe.bin $t >> $O/umma_probe.txt 2>&1
  echo "exit $?" >> $O/umma_probe.txt
done
cat $O/umma_probe.txt
