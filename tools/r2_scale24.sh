#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 3 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
head -c 300 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
