#!/bin/bash
# latency regime: one 512-thread CTA per SM could use 128 registers per thread (no spills, less rematerialisation)
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in head regs128; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v config 2: "; python tools/prof_run.py --config 2 --launches 4 | grep "launch 3"
  echo -n "$v 45x50: "; python tools/prof_run.py --config 2 --objects 45 --launches 4 | grep "launch 3"
  echo -n "$v 33x50: "; python tools/prof_run.py --config 2 --objects 33 --launches 4 | grep "launch 3"
  echo -n "$v 100x50: "; python tools/prof_run.py --config 2 --objects 100 --launches 4 | grep "launch 3"
  echo -n "$v 50x20: "; python tools/prof_run.py --config 2 --views 20 --launches 4 | grep "launch 3"
done
done
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_regs128.so
python tools/prof_run.py --config 2 --cycles
} > gpurun_out/s34_regs128.log 2>&1
cat gpurun_out/s34_regs128.log
