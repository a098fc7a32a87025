#!/bin/bash
# dense-regime sweep: views x code layout x CTA size  ->  M unit/s   (tools/layout_sweep.sh "8 20 30 50 100")
cd "$(dirname "$0")/.."
for v in $1; do
  n=$(( 120000 / v )); [ $n -gt 6000 ] && n=6000
  for lay in 1 2; do
    for t in 128 192 256 320; do
      [ $lay = 2 ] && [ $t -gt 256 ] && continue
      echo -n "V=$v n=$n layout=$lay T=$t: "
      python tools/prof_run.py --config 3 --objects $n --views $v --iters 30 --launches 3 --threads $t --layout $lay | grep "launch 2" | cut -d' ' -f3-8
    done
  done
done
