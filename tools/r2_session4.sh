#!/bin/bash
mkdir -p gpurun_out
./build/bp2 > gpurun_out/s4_bp2.log 2>&1
tools/ab_run.sh f2 g8 g16 p224 p212 > gpurun_out/s4_ab.log 2>&1
cat gpurun_out/s4_bp2.log gpurun_out/s4_ab.log
