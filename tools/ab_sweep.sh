#!/bin/bash
# thread-count sweep of A/B library variants:  tools/ab_sweep.sh "u1b u2" "5 3" "128 192 256"
cd "$(dirname "$0")/.."
for v in $1; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  for c in $2; do
    extra=""; [ "$c" = 5 ] && extra="--objects 9472"
    for t in $3; do
      echo -n "$v config $c T=$t: "
      python tools/prof_run.py --config $c $extra --iters 40 --launches 3 --threads $t | grep "launch 2"
    done
  done
done
