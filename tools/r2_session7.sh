#!/bin/bash
mkdir -p gpurun_out
tools/ab_run.sh f2 b5 > gpurun_out/s7_ab.log 2>&1
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_b5.so
python tools/prof_run.py --config 5 --objects 9472 --iters 200 --launches 2 >> gpurun_out/s7_ab.log 2>&1
python tools/prof_run.py --config 3 --iters 200 --launches 2 >> gpurun_out/s7_ab.log 2>&1
cat gpurun_out/s7_ab.log
