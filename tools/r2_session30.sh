#!/bin/bash
# 3-CTA clusters in the launch heuristics: GPU tests + where is the cliff (objects x 3 close to the SM count)?
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s30_test.log 2>&1
grep -E "passed|failed" gpurun_out/s30_test.log
{
for n in 38 42 44 45 46 47 48; do
  echo -n "${n}x50 auto: "; python tools/prof_run.py --config 2 --objects $n --launches 4 | grep "launch 3"
  echo -n "${n}x50 cluster 2: "; python tools/prof_run.py --config 2 --objects $n --launches 4 --cluster 2 | grep "launch 3"
  echo -n "${n}x50 cluster 3: "; python tools/prof_run.py --config 2 --objects $n --launches 4 --cluster 3 | grep "launch 3"
done
for n in 20 30 37; do
  for v in 24 30; do
    echo -n "${n}x${v} auto: "; python tools/prof_run.py --config 2 --objects $n --views $v --launches 4 | grep "launch 3"
    echo -n "${n}x${v} cluster 2: "; python tools/prof_run.py --config 2 --objects $n --views $v --launches 4 --cluster 2 | grep "launch 3"
  done
done
echo -n "config 2 auto: "; python tools/prof_run.py --config 2 --launches 4 | grep "launch 3"
} > gpurun_out/s30_cluster3.log 2>&1
cat gpurun_out/s30_cluster3.log
