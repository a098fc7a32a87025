#!/bin/bash
# slice count / CTA size of the latency regime once more, now with the 128-register build
mkdir -p gpurun_out
{
for args in "" "--max-slices 16" "--max-slices 20" "--max-slices 12" "--threads 448" "--threads 480" "--threads 384 --max-slices 15"; do
  echo -n "config 2 [$args]: "; python tools/prof_run.py --config 2 --launches 4 $args | grep "launch 3"
done
for args in "" "--max-slices 16" "--threads 384"; do
  echo -n "100x50 [$args]: "; python tools/prof_run.py --config 2 --objects 100 --launches 4 $args | grep "launch 3"
  echo -n "50x20 [$args]: "; python tools/prof_run.py --config 2 --views 20 --launches 4 $args | grep "launch 3"
done
} > gpurun_out/s41_solo_sweep.log 2>&1
cat gpurun_out/s41_solo_sweep.log
