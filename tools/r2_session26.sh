#!/bin/bash
# compact build with immediate barrier ids (3 barriers instead of 16 reserved): 4 vs 5 resident CTAs per SM
mkdir -p gpurun_out
( time python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) > gpurun_out/s26_test.log 2>&1
grep -E "passed|failed" gpurun_out/s26_test.log
{
for rep in 1 2; do
for v in spec bar4 bar5; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v config 5: "; python tools/prof_run.py --config 5 --objects 9472 --iters 100 --launches 3 | grep "launch 2"
  echo -n "$v config 3: "; python tools/prof_run.py --config 3 --iters 100 --launches 3 | grep "launch 2"
done
done
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_bar5.so
echo -n "bar5 config 5 (all 50k, 200 it): "; python tools/prof_run.py --config 5 --launches 2 | grep "launch 1"
echo -n "bar5 600x50: "; python tools/prof_run.py --config 2 --objects 600 --launches 3 | grep "launch 2"
echo -n "bar5 2000x40: "; python tools/prof_run.py --config 3 --views 40 --iters 100 --launches 3 | grep "launch 2"
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_bar4.so
echo -n "bar4 config 5 (all 50k, 200 it): "; python tools/prof_run.py --config 5 --launches 2 | grep "launch 1"
echo -n "bar4 600x50: "; python tools/prof_run.py --config 2 --objects 600 --launches 3 | grep "launch 2"
echo -n "bar4 2000x40: "; python tools/prof_run.py --config 3 --views 40 --iters 100 --launches 3 | grep "launch 2"
} > gpurun_out/s26_bar.log 2>&1
cat gpurun_out/s26_bar.log
