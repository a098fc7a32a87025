#!/bin/bash
# cluster sizes from the occupancy query (one CTA per SM): capacities + automatic choice around the cliffs + tests
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s32_test.log 2>&1
grep -E "passed|failed" gpurun_out/s32_test.log
{
python -c "
from odam_b200 import api
print('cluster capacity (objects with exclusive SMs):', api.cluster_capacity(0))
"
for n in 31 32 33 34 37 45 46 50 74 75; do
  echo -n "${n}x50 auto: "; python tools/prof_run.py --config 2 --objects $n --launches 4 | grep "launch 3"
done
for n in 33; do
  for c in 4 3; do echo -n "${n}x50 cluster $c: "; python tools/prof_run.py --config 2 --objects $n --launches 4 --cluster $c | grep "launch 3"; done
done
} > gpurun_out/s32_occq.log 2>&1
cat gpurun_out/s32_occq.log
