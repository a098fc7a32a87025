#!/bin/bash
# speculative slot placement in B0.1 (pool_place only in iterations where a split count changed): tests + A/B
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s24_test.log 2>&1
grep -E "passed|failed" gpurun_out/s24_test.log
tools/ab_run.sh base spec base spec > gpurun_out/s24_ab.log 2>&1
cat gpurun_out/s24_ab.log
