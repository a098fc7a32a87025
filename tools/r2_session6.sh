#!/bin/bash
# slice-count sweep in the dense regime (launch heuristics) with the FFMA2 build
mkdir -p gpurun_out
for s in 4 6 8 10 12 16; do echo -n "config 3 max_slices $s: "; python tools/prof_run.py --config 3 --iters 100 --launches 3 --max-slices $s | grep "launch 2"; done > gpurun_out/s6_slices.log 2>&1
for s in 4 6 8 10 12 16; do echo -n "config 5 max_slices $s: "; python tools/prof_run.py --config 5 --objects 9472 --iters 100 --launches 3 --max-slices $s | grep "launch 2"; done >> gpurun_out/s6_slices.log 2>&1
for t in 256 288 320 384; do for s in 2 3 4; do echo -n "config 4 threads $t max_slices $s: "; python tools/prof_run.py --config 4 --iters 40 --launches 3 --threads $t --max-slices $s | grep "launch 2"; done; done >> gpurun_out/s6_slices.log 2>&1
for s in 16 20 25; do for t in 512 640 768; do echo -n "config 2 threads $t max_slices $s: "; python tools/prof_run.py --config 2 --launches 3 --threads $t --max-slices $s | grep "launch 2"; done; done >> gpurun_out/s6_slices.log 2>&1
cat gpurun_out/s6_slices.log
