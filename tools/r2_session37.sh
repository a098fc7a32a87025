#!/bin/bash
# ptxas --register-usage-level 0 / 10 against the default 5
mkdir -p gpurun_out
tools/ab_run.sh solo rul0 rul10 solo rul0 rul10 > gpurun_out/s37_rul.log 2>&1
cat gpurun_out/s37_rul.log
