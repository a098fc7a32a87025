#!/bin/bash
# production kernels without the per-phase cycle marks (the marks as a separate instantiation)?
mkdir -p gpurun_out
tools/ab_run.sh final nomarks final nomarks > gpurun_out/s42_nomarks.log 2>&1
cat gpurun_out/s42_nomarks.log
