#!/bin/bash
# 2-GPU session: multi-GPU tests (one process / many devices, torchrun sharded) and the --gpus 2 bench line
mkdir -p gpurun_out
( time python -m pytest tests/test_multi_gpu.py tests/test_parity_gpu.py::test_ragged_launch_of_reference_cases -m gpu -q -s ) > gpurun_out/s3_test.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s3_test.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/s3_bench_n2.json 2> gpurun_out/s3_bench_n2.err
tail -15 gpurun_out/s3_test.log; head -c 600 gpurun_out/s3_bench_n2.json; tail -5 gpurun_out/s3_bench_n2.err
