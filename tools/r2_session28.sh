#!/bin/bash
# 3-CTA clusters: 38..49 objects (3n <= 148), and 50 objects with 150 CTAs on 148 SMs (two SMs carry two CTAs)
mkdir -p gpurun_out
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_cl3.so
{
for n in 40 45 49; do
  for args in "" "--cluster 3" "--cluster 3 --threads 512 --max-slices 25 --layout 1"; do
    echo -n "${n}x50 [$args]: "; python tools/prof_run.py --config 2 --objects $n --launches 4 $args | grep "launch 3"
  done
done
for args in "" "--cluster 3" "--cluster 3 --threads 512 --max-slices 25 --layout 1" "--cluster 3 --threads 512 --max-slices 20 --layout 1" "--cluster 3 --threads 384 --max-slices 25 --layout 1"; do
  echo -n "config 2 [$args]: "; python tools/prof_run.py --config 2 --launches 4 $args | grep "launch 3"
done
for args in "" "--cluster 3 --threads 512 --max-slices 25 --layout 1"; do
  echo -n "45x30 [$args]: "; python tools/prof_run.py --config 2 --objects 45 --views 30 --launches 4 $args | grep "launch 3"
  echo -n "45x20 [$args]: "; python tools/prof_run.py --config 2 --objects 45 --views 20 --launches 4 $args | grep "launch 3"
done
python tools/prof_run.py --config 2 --cluster 3 --threads 512 --max-slices 25 --layout 1 --cycles
} > gpurun_out/s28_cluster3.log 2>&1
cat gpurun_out/s28_cluster3.log
