#!/bin/bash
# where the launch heuristics switch regimes: objects x {auto, alternatives} at 50 views  ->  M unit/s
cd "$(dirname "$0")/.."
for n in $1; do
  echo -n "n=$n auto: "; python tools/prof_run.py --config 2 --objects $n --iters 100 --launches 3 | grep "launch 2" | cut -d' ' -f3-6
  for alt in "--cluster 1 --threads 512" "--cluster 2 --threads 512" "--cluster 1 --threads 256 --layout 1" "--cluster 1 --threads 256 --layout 2" "--cluster 2 --threads 256 --layout 1"; do
    echo -n "n=$n $alt: "; python tools/prof_run.py --config 2 --objects $n --iters 100 --launches 3 $alt | grep "launch 2" | cut -d' ' -f3-6
  done
done
