#!/bin/bash
# 4-CTA clusters close to the SM count: is there a cliff like the one of 3-CTA clusters at 46 objects?
mkdir -p gpurun_out
{
for n in 30 32 34 35 36 37; do
  for c in 4 3; do
    echo -n "${n}x50 cluster $c: "; python tools/prof_run.py --config 2 --objects $n --launches 4 --cluster $c | grep "launch 3"
  done
done
for n in 70 72 74; do
  for c in 2 1; do
    echo -n "${n}x50 cluster $c: "; python tools/prof_run.py --config 2 --objects $n --launches 4 --cluster $c | grep "launch 3"
  done
done
} > gpurun_out/s31_cluster4.log 2>&1
cat gpurun_out/s31_cluster4.log
