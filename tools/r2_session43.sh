#!/bin/bash
# cycle marks as a separate instantiation (kProf): tests, timing of the production kernels, the cycle table still works
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s43_test.log 2>&1
grep -E "passed|failed" gpurun_out/s43_test.log
{
tools/ab_run.sh tprof
python tools/prof_run.py --config 2 --cycles | tail -17
} > gpurun_out/s43_tprof.log 2>&1
cat gpurun_out/s43_tprof.log
