#!/bin/bash
# config 2 launch variants: more CTAs than SMs on purpose (4-CTA clusters, two CTAs on some SMs)
mkdir -p gpurun_out
{
for args in "" "--cluster 4" "--cluster 4 --max-slices 25" "--cluster 4 --max-slices 16" "--cluster 4 --threads 256" "--cluster 4 --threads 384" "--cluster 2 --threads 256" "--cluster 4 --threads 512 --layout 1 --max-slices 20"; do
  echo -n "config 2 [$args]: "; python tools/prof_run.py --config 2 --launches 4 $args | grep "launch 3"
done
for args in "" "--cluster 4" "--cluster 2"; do
  echo -n "74x50 [$args]: "; python tools/prof_run.py --config 2 --objects 74 --launches 4 $args | grep "launch 3"
  echo -n "100x50 [$args]: "; python tools/prof_run.py --config 2 --objects 100 --launches 4 $args | grep "launch 3"
  echo -n "50x20 [$args]: "; python tools/prof_run.py --config 2 --views 20 --launches 4 $args | grep "launch 3"
done
python tools/prof_run.py --config 2 --cluster 4 --cycles
} > gpurun_out/s23_cluster4.log 2>&1
cat gpurun_out/s23_cluster4.log
