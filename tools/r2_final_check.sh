#!/bin/bash
# what the driver does at round end, in one go: build check, GPU tests, smoke, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/final_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_test.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
( time python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -2 gpurun_out/final_smoke.log; tail -4 gpurun_out/final_test.log; head -c 300 gpurun_out/final_bench.json; echo; head -c 300 gpurun_out/final_bench_ref.json; echo; tail -3 gpurun_out/final_bench.err
