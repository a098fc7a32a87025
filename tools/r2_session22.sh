#!/bin/bash
# F fan-out A/B: latency regime (config 2, 20-view and 30-view variants) + dense sanity; parity tests with the new default
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s22_test.log 2>&1
grep -E "passed|failed" gpurun_out/s22_test.log
{
for rep in 1 2; do
for v in base fan1 fan2 fan4; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v config 2: "; python tools/prof_run.py --config 2 --launches 4 | grep "launch 3"
done
done
for v in base fan4; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v 50x20: "; python tools/prof_run.py --config 2 --views 20 --launches 4 | grep "launch 3"
  echo -n "$v 74x50: "; python tools/prof_run.py --config 2 --objects 74 --launches 4 | grep "launch 3"
  echo -n "$v 120x30: "; python tools/prof_run.py --config 2 --objects 120 --views 30 --launches 4 | grep "launch 3"
  echo -n "$v config 5: "; python tools/prof_run.py --config 5 --objects 9472 --iters 40 --launches 3 | grep "launch 2"
  echo -n "$v config 3: "; python tools/prof_run.py --config 3 --iters 40 --launches 3 | grep "launch 2"
  echo -n "$v config 4: "; python tools/prof_run.py --config 4 --iters 40 --launches 3 | grep "launch 2"
done
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_fan4.so
python tools/prof_run.py --config 2 --cycles
} > gpurun_out/s22_ab.log 2>&1
cat gpurun_out/s22_ab.log
