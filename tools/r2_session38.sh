#!/bin/bash
# long tracks (config 4): 3 CTAs x 256 threads with 80 registers instead of 64?
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in rl10 s256b3; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v config 4 auto: "; python tools/prof_run.py --config 4 --iters 40 --launches 3 | grep "launch 2"
  echo -n "$v config 4 threads 256: "; python tools/prof_run.py --config 4 --iters 40 --launches 3 --threads 256 | grep "launch 2"
  echo -n "$v 600x50 threads 256 layout 1: "; python tools/prof_run.py --config 2 --objects 600 --launches 3 --threads 256 --layout 1 | grep "launch 2"
done
done
} > gpurun_out/s38_s256b3.log 2>&1
cat gpurun_out/s38_s256b3.log
