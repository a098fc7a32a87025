#!/bin/bash
# oriented boxes fused behind the optimiser inside the host call: tests + call-site timing
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s33_test.log 2>&1
grep -E "passed|failed" gpurun_out/s33_test.log; tail -5 gpurun_out/s33_test.log
python tools/time_optim_process.py > gpurun_out/s33_callsite.log 2>&1
head -12 gpurun_out/s33_callsite.log
