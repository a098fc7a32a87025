#!/bin/bash
# the 128-register build for one-CTA-per-SM launches wired into the launch path: tests + A/B
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s35_test.log 2>&1
grep -E "passed|failed" gpurun_out/s35_test.log
tools/ab_run.sh head solo head solo > gpurun_out/s35_ab.log 2>&1
cat gpurun_out/s35_ab.log
