#!/bin/bash
# Round-2 GPU session 2: full -m gpu suite (new reference-pinned, post-processing, drop-in tests), pipe micro-benchmark,
# the default bench line, call-site timing.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/s2_test.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s2_test.log
./build/mbp > gpurun_out/s2_mbp.log 2>&1
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
python tools/time_optim_process.py > gpurun_out/s2_callsite.log 2>&1
tail -3 gpurun_out/s2_test.log; head -c 1500 gpurun_out/s2_bench.json; tail -5 gpurun_out/s2_bench.err
