#!/bin/bash
mkdir -p gpurun_out
for v in occ4 occ3 occ2 occ1; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v config 5: "; python tools/prof_run.py --config 5 --objects 9472 --iters 100 --launches 3 | grep "launch 2"
  echo -n "$v config 3: "; python tools/prof_run.py --config 3 --iters 100 --launches 3 | grep "launch 2"
done > gpurun_out/s19_occ.log 2>&1
cat gpurun_out/s19_occ.log
