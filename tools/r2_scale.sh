#!/bin/bash
# bench.py at N GPUs as the driver launches it (N = number of visible GPUs)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 3 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/multi_gpu_n$N.log 2>&1
tail -3 gpurun_out/multi_gpu_n$N.log; head -c 400 gpurun_out/bench_n$N.json; tail -4 gpurun_out/bench_n$N.err
