#!/bin/bash
# phase D two samples interleaved; resolve on packed pairs: tests + A/B
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s25_test.log 2>&1
grep -E "passed|failed" gpurun_out/s25_test.log
tools/ab_run.sh spec d2 res d2res spec d2res > gpurun_out/s25_ab.log 2>&1
cat gpurun_out/s25_ab.log
