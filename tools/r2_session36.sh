#!/bin/bash
# 128-register build: more ILP in the serial phases (D two samples interleaved, full unroll of F's arg search, 8-way combine)
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in solo s_d2 s_res s_comb s_all; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v config 2: "; python tools/prof_run.py --config 2 --launches 4 | grep "launch 3"
  echo -n "$v 100x50: "; python tools/prof_run.py --config 2 --objects 100 --launches 4 | grep "launch 3"
done
done
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_s_all.so
python tools/prof_run.py --config 2 --cycles
} > gpurun_out/s36_solo_ilp.log 2>&1
cat gpurun_out/s36_solo_ilp.log
