#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): the round-2 evidence set.  Outputs under gpurun_out/; summarise in the build
# container with `python tools/summarise_profiles.py r02`.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/r02_parity.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_parity.txt
# launch list of the bench command (headline step only: the sweep / strong-scaling / call-site legs generate their
# scenes with thousands of torch kernels that would bury the list; they are timed by the un-profiled run below)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep --no-strong --no-call-site > gpurun_out/bench_under_ncu.log 2>&1
declare -A EXTRA=( [2]="" [3]="" [4]="--iters 40" [5]="--objects 2368" )
for cfg in 2 3 4 5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sq_optimize -s 1 -c 1 -o gpurun_out/full_c$cfg \
      python tools/prof_run.py --config $cfg ${EXTRA[$cfg]} > gpurun_out/full_c$cfg.log 2>&1
  echo "python tools/prof_run.py --config $cfg ${EXTRA[$cfg]}" > gpurun_out/full_c$cfg.cmd
done
for cfg in 2 3 5; do
  e=""; [ "$cfg" = 5 ] && e="--objects 9472"
  python tools/prof_run.py --config $cfg $e --launches 2 --cycles > gpurun_out/cycles_c$cfg.log 2>&1
done
python tools/prof_run.py --config 4 --iters 40 --launches 2 --cycles > gpurun_out/cycles_c4.log 2>&1
( time python bench.py --steps 30 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/time_optim_process.py > gpurun_out/callsite.log 2>&1
ls -la gpurun_out | tail -30
