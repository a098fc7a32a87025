#!/bin/bash
# Round-2 GPU session 1 (runs ON THE GPU BOX under gpurun): parity suite with the new reference-pinned cases, FFMA2 pipe
# micro-benchmark, A/B of the scalar and the packed phase E, CTA-size sweep, per-phase cycles, one source-level ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.log 2>&1
( time python -m pytest tests -m gpu -q -s -x ) > gpurun_out/s1_test.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_test.log
./build/mb2 > gpurun_out/s1_mb2.log 2>&1
tools/ab_run.sh f0 f2 > gpurun_out/s1_ab.log 2>&1
tools/ab_sweep.sh "f2" "5 3" "128 192 256" > gpurun_out/s1_sweep.log 2>&1
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_f2.so
python tools/prof_run.py --config 5 --objects 9472 --iters 200 --launches 2 --cycles > gpurun_out/s1_cyc5.log 2>&1
python tools/prof_run.py --config 3 --iters 200 --launches 2 --cycles > gpurun_out/s1_cyc3.log 2>&1
python tools/prof_run.py --config 2 --launches 3 --cycles > gpurun_out/s1_cyc2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sq_optimize -s 1 -c 1 -o gpurun_out/s1_c5 \
    python tools/prof_run.py --config 5 --objects 2368 --iters 200 > gpurun_out/s1_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sq_optimize -s 1 -c 1 -o gpurun_out/s1_c2 \
    python tools/prof_run.py --config 2 > gpurun_out/s1_ncu_c2.log 2>&1
ls -la gpurun_out | tail -20
