#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s -x ) > gpurun_out/s5_test.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s5_test.log
grep -E "passed|failed" gpurun_out/s5_test.log
tools/ab_run.sh f2 sh1 > gpurun_out/s5_ab.log 2>&1
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_sh1.so
for c in "5 --objects 9472" "3" "4" "2"; do
  for m in 0 1; do echo "config $c combine $m"; python tools/prof_run.py --config $c --iters 100 --launches 3 --combine $m | grep "launch 2"; done
done > gpurun_out/s5_combine.log 2>&1
python tools/prof_run.py --config 2 --launches 3 --cycles > gpurun_out/s5_cyc2.log 2>&1
python tools/prof_run.py --config 5 --objects 9472 --iters 200 --launches 2 --cycles > gpurun_out/s5_cyc5.log 2>&1
cat gpurun_out/s5_ab.log gpurun_out/s5_combine.log; tail -22 gpurun_out/s5_cyc2.log
