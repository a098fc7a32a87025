#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): ncu launch list of the bench command + one full-set capture of the optimiser
# kernel per config.  Outputs under gpurun_out/; summarise here with tools/summarise_profiles.py.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
for cfg in 2 4 5; do
  extra=""; [ "$cfg" = "5" ] && extra="--objects 9472"
  ncu --set full --clock-control none --import-source on -k regex:sq_optimize -s 1 -c 1 -o gpurun_out/full_c$cfg \
      python tools/prof_run.py --config $cfg --iters 40 $extra > gpurun_out/full_c$cfg.log 2>&1
done
python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
ls -la gpurun_out
