#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s8_test.log 2>&1
grep -E "passed|failed" gpurun_out/s8_test.log
tools/ab_run.sh f2 r1 r2 > gpurun_out/s8_ab.log 2>&1
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_r2.so
python tools/prof_run.py --config 2 --launches 3 --cycles > gpurun_out/s8_cyc2.log 2>&1
python tools/prof_run.py --config 5 --objects 9472 --iters 200 --launches 2 --cycles > gpurun_out/s8_cyc5.log 2>&1
cat gpurun_out/s8_ab.log; tail -18 gpurun_out/s8_cyc2.log; tail -18 gpurun_out/s8_cyc5.log
