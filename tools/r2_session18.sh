#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s18_test.log 2>&1
grep -E "passed|failed" gpurun_out/s18_test.log
tools/ab_run.sh st0 st1 st0 st1 > gpurun_out/s18_ab.log 2>&1
cat gpurun_out/s18_ab.log
