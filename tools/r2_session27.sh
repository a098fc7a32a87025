#!/bin/bash
# latency regime: is the instruction footprint (71 KB hot per iteration in the straight-line build) the hidden cost?
# compact build (one sampler copy for both grids, 44 KB hot) with 512-thread CTAs, scan blocks of 4 / 8 / 16 points
mkdir -p gpurun_out
{
for v in c512 c512g8 c512g16; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo -n "$v config 2 straight: "; python tools/prof_run.py --config 2 --launches 4 | grep "launch 3"
  echo -n "$v config 2 compact 512: "; python tools/prof_run.py --config 2 --launches 4 --layout 2 --threads 512 | grep "launch 3"
  echo -n "$v 74x50 compact 512: "; python tools/prof_run.py --config 2 --objects 74 --launches 4 --layout 2 --threads 512 | grep "launch 3"
  echo -n "$v 50x20 compact 512: "; python tools/prof_run.py --config 2 --views 20 --launches 4 --layout 2 --threads 512 | grep "launch 3"
  echo -n "$v config 4 compact 256: "; python tools/prof_run.py --config 4 --iters 40 --launches 3 --layout 2 | grep "launch 2"
done
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_c512g16.so
python tools/prof_run.py --config 2 --layout 2 --threads 512 --cycles
} > gpurun_out/s27_compact512.log 2>&1
cat gpurun_out/s27_compact512.log
