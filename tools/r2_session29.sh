#!/bin/bash
# how does phase E's time (thread 0's view) scale with the views per CTA?  50 objects, 2-CTA clusters, 512 threads
mkdir -p gpurun_out
{
for v in 8 16 24 32 50 64 100; do
  echo "== views $v"; python tools/prof_run.py --config 2 --views $v --cluster 2 --threads 512 --cycles | grep -E "launch 1|E project|combine|total"
done
echo "== views 50 cluster 1 (one CTA per object)"; python tools/prof_run.py --config 2 --cluster 1 --threads 512 --cycles | grep -E "launch 1|E project|combine|total"
echo "== views 50 max-slices 10"; python tools/prof_run.py --config 2 --cluster 2 --threads 512 --max-slices 10 --cycles | grep -E "launch 1|E project|combine|total"
echo "== views 50 max-slices 5"; python tools/prof_run.py --config 2 --cluster 2 --threads 512 --max-slices 5 --cycles | grep -E "launch 1|E project|combine|total"
} > gpurun_out/s29_escale.log 2>&1
cat gpurun_out/s29_escale.log
