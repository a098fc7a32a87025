#!/bin/bash
mkdir -p gpurun_out
for cfgv in "3 30" "5 20" "2 50"; do set -- $cfgv
  for n in 160 250 296; do
    echo -n "views=$2 n=$n auto: "; python tools/prof_run.py --config $1 --objects $n --iters 100 --launches 3 | grep "launch 2" | cut -d' ' -f3-6
    for s in 8 12 16; do echo -n "views=$2 n=$n T=512 slices=$s: "; python tools/prof_run.py --config $1 --objects $n --iters 100 --launches 3 --threads 512 --cluster 1 --max-slices $s | grep "launch 2" | cut -d' ' -f3-6; done
    echo -n "views=$2 n=$n T=256 (old): "; python tools/prof_run.py --config $1 --objects $n --iters 100 --launches 3 --threads 256 --cluster 1 | grep "launch 2" | cut -d' ' -f3-6
  done
done > gpurun_out/s12_mid.log 2>&1
cat gpurun_out/s12_mid.log
