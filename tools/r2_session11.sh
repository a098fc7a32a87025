#!/bin/bash
mkdir -p gpurun_out
for n in 150 200 250 300 400 500 700; do
  for alt in "" "--cluster 2 --threads 256 --layout 2" "--cluster 2 --threads 256 --layout 1" "--cluster 2 --threads 512" "--cluster 1 --threads 512"; do
    echo -n "n=$n views=30 [$alt]: "; python tools/prof_run.py --config 3 --objects $n --iters 100 --launches 3 $alt | grep "launch 2" | cut -d' ' -f3-6
  done
done > gpurun_out/s11_mid.log 2>&1
for n in 150 250 400; do
  for alt in "" "--cluster 2 --threads 256 --layout 2" "--cluster 2 --threads 512"; do
    echo -n "n=$n views=20 [$alt]: "; python tools/prof_run.py --config 5 --objects $n --iters 100 --launches 3 $alt | grep "launch 2" | cut -d' ' -f3-6
  done
done >> gpurun_out/s11_mid.log 2>&1
cat gpurun_out/s11_mid.log
