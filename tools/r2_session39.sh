#!/bin/bash
# warp-uniform validity variant in phase E: tests + A/B (synthetic scenes are all-valid: expect no change)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/s39_test.log 2>&1
grep -E "passed|failed" gpurun_out/s39_test.log
tools/ab_run.sh rl10 wuni rl10 wuni > gpurun_out/s39_ab.log 2>&1
cat gpurun_out/s39_ab.log
